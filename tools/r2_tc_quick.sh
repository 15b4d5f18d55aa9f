# quick look at the dense kernel: dense tests + configs[3] timing, no ncu
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dense.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --preset wide64 --channels 4736 --steps 5 --warmup 2 --e2e-steps 0 --no-cpu-baseline --no-sustain --no-cfg4 > gpurun_out/r2_bench_tc.json 2> gpurun_out/r2_bench_tc.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_tc.json').read().strip().splitlines()[-1])
print('wide64 value',d['value'],'roofline',d['roofline']['frac'],'kernel ms',d['roofline']['avg_kernel_ms'],'frames',d['frames_ok'])
PY
