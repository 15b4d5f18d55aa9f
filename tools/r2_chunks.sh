# multi-chunk launches of k_demod: whole step per launch against one launch per chunk, same box
mkdir -p gpurun_out
for lc in 1 0 1 0; do
  timeout 600 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-cfg4 --launch-chunks $lc > gpurun_out/r2_bench_lc$lc.json 2> gpurun_out/r2_bench_lc$lc.err; echo "lc=$lc rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_lc$lc.json').read().strip().splitlines()[-1])
print('launch-chunks $lc: value', d['value'], 'ms/step', d['ms_per_step'], 'roofline frac', d['roofline']['frac'], 'of step', d['roofline']['frac_of_step_time'], 'launches', d['gpu_launches'], 'frames_ok', d['frames_ok'], 'sustained', d.get('sustained',{}).get('hbm_frac_of_step_time'), d.get('sustained',{}).get('clocks',{}).get('sm_mhz'))
PY
done
