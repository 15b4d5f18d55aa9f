# ncu evidence for the committed kernels (run AFTER the same commands exited 0 without ncu)
python bench.py --steps 3 --warmup 2 --e2e-steps 0 --no-cpu-baseline --no-cfg4 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_demod --launch-skip 3 -c 1 -f -o gpurun_out/prof_r1j python bench.py --steps 3 --warmup 2 --e2e-steps 0 --no-cpu-baseline --no-cfg4 > gpurun_out/ncu_j.log 2>&1
python bench.py --preset wide64 --channels 4096 --steps 3 --warmup 2 --e2e-steps 0 --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_demod_tc --launch-skip 3 -c 1 -f -o gpurun_out/prof_tc_e python bench.py --preset wide64 --channels 4096 --steps 3 --warmup 2 --e2e-steps 0 --no-cpu-baseline > gpurun_out/ncu_tc_e.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1z.csv python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo profiled
