# final dense kernel of round 2: whole GPU suite, time (3x), the two bounds (no epilogue work / no MMAs), ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_tc_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_tc_final.log
for i in 1 2 3; do bash tools/r2_tc_quick.sh | tail -1; done
cp gpurun_out/r2_bench_tc.json gpurun_out/r2_bench_tc_final.json
VARIANTS="skip_epi skip_mma" bash tools/r2_tc_bounds.sh | grep "kernel ms"
ANM_BENCH_CHUNKS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_demod_tc --launch-skip 2 -c 1 -f -o gpurun_out/prof_r2_tc_${TAG:-x} python bench.py --preset wide64 --channels 4736 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sustain --no-cfg4 > gpurun_out/r2_ncu_tc.log 2>&1
echo done
