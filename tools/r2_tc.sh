# gen-2 tensor-core kernel iteration: dense tests (both accumulator modes), cfg4 timing, ncu capture
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dense.py tests/test_gpu_parity.py -m gpu -x -q -k "dense or wide64 or golden or tone_energies or multi" > gpurun_out/r2_pytest_tc.log 2>&1; echo "dense rc=$?" >> gpurun_out/r2_pytest_tc.log
tail -8 gpurun_out/r2_pytest_tc.log
ANM_TC_NO_BIAS=1 timeout 300 python -m pytest tests/test_dense.py -m gpu -x -q > gpurun_out/r2_pytest_tc_nobias.log 2>&1; echo "dense (unbiased accumulators) rc=$?" >> gpurun_out/r2_pytest_tc_nobias.log
tail -4 gpurun_out/r2_pytest_tc_nobias.log
timeout 600 python bench.py --preset wide64 --channels 4736 --steps 5 --warmup 2 --e2e-steps 0 --no-cpu-baseline --no-sustain --no-cfg4 > gpurun_out/r2_bench_tc.json 2> gpurun_out/r2_bench_tc.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_tc.json').read().strip().splitlines()[-1])
print('wide64 value',d['value'],'ms/step',d['ms_per_step'],'roofline',d['roofline']['frac'],'kernel ms',d['roofline']['avg_kernel_ms'],'frames',d['frames_ok'])
PY
ANM_BENCH_CHUNKS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_demod_tc --launch-skip 2 -c 1 -f -o gpurun_out/prof_r2_tc_${TAG:-x} python bench.py --preset wide64 --channels 4736 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sustain --no-cfg4 > gpurun_out/r2_ncu_tc.log 2>&1
echo done
