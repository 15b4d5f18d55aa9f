# phase-major ring + candidate layout: the whole GPU suite, then both kernels' times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_ring.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_ring.log
bash tools/r2_tc_quick.sh | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-sustain --no-cfg4 > gpurun_out/r2_bench_ring.json 2> gpurun_out/r2_bench_ring.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_ring.json').read().strip().splitlines()[-1])
print('north star value',d['value'],'roofline',d['roofline']['frac'],'kernel ms',d['roofline']['avg_kernel_ms'])
PY
