# round 2, GPU run 2: gen-2 tensor-core kernel: dense tests, full suite, cfg4 numbers, ncu
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dense.py tests/test_gpu_parity.py -m gpu -x -q -k "dense or wide64 or golden or tone_energies" > gpurun_out/r2_pytest2a.log 2>&1; echo "dense rc=$?" >> gpurun_out/r2_pytest2a.log
tail -15 gpurun_out/r2_pytest2a.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -5 gpurun_out/r2_pytest2.log
ANM_BENCH_CHUNKS=6 timeout 600 python bench.py --steps 10 --warmup 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench2.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'roofline',d['roofline']['frac'],d['roofline']['frac_of_step_time'],'kernel ms',d['roofline']['avg_kernel_ms'])
print('sustained',d.get('sustained'))
print('e2e',d['e2e']['value'],d['e2e']['frac_of_h2d_ceiling'])
print('cfg4',{k:v for k,v in d['cfg4'].items() if k!='ncu'})
PY
ANM_BENCH_CHUNKS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_demod_tc --launch-skip 4 -c 1 -f -o gpurun_out/prof_r2_tc_a python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sustain > gpurun_out/r2_ncu_tc_a.log 2>&1
echo done
