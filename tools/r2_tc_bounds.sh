# timing experiment: the dense kernel without its epilogue work, and without its MMAs (results are garbage, only the times count);
# the variant libraries are built by hand with -DANM_TC_DEBUG_SKIP_EPI / -DANM_TC_DEBUG_SKIP_MMA into tools/_variants/
mkdir -p gpurun_out
for v in ${VARIANTS:-SKIP_EPI SKIP_MMA}; do
  ANM_LIB_PATH=$PWD/tools/_variants/libanmodem_$v.so timeout 600 python bench.py --preset wide64 --channels 4736 --steps 5 --warmup 2 --e2e-steps 0 --no-cpu-baseline --no-sustain --no-cfg4 > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err; echo "$v bench rc=$?"; tail -2 gpurun_out/r2_bench_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_$v.json').read().strip().splitlines()[-1])
print('$v kernel ms',d['roofline']['avg_kernel_ms'])
PY
done
bash tools/r2_tc_quick.sh | tail -1
