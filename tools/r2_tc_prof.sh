set -x
ANM_BENCH_CHUNKS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_demod_tc --launch-skip 2 -c 1 -f -o gpurun_out/prof_r2_tc_${TAG:-x} python bench.py --preset wide64 --channels 4736 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sustain --no-cfg4 > gpurun_out/r2_ncu_tc.log 2>&1
tail -3 gpurun_out/r2_ncu_tc.log
