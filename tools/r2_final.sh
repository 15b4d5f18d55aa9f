# round 2 evidence run on one B200: full GPU suite, smoke, both bench arms, CELT decode throughput, launch lists, ncu --set full of the demodulator
# kernels and the CELT kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log
tail -4 gpurun_out/r2_pytest_final.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke_final.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref_final.json 2> gpurun_out/r2_ref_final.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_final.err
timeout 600 python tools/celt_bench.py --streams 4096 --cpu-baseline > gpurun_out/r2_celt_bench.json 2> gpurun_out/r2_celt_bench.err; echo "celt bench rc=$?"; cat gpurun_out/r2_celt_bench.json
ANM_BENCH_CHUNKS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sustain > gpurun_out/r2_ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_celt_launches.csv python tools/celt_bench.py --streams 4096 --reps 1 > gpurun_out/r2_celt_ncu_launch.log 2>&1
ANM_BENCH_CHUNKS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_demod --launch-skip 1 -c 1 -f -o gpurun_out/prof_r2d python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-cfg4 --no-sustain > gpurun_out/r2_ncu_d.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_celt_spectrum|k_celt_blocks|k_celt_overlap|k_celt_deemphasis" --launch-skip 2 -c 4 -f -o gpurun_out/prof_r2_celt python tools/celt_bench.py --streams 1024 --reps 1 > gpurun_out/r2_ncu_celt.log 2>&1
echo done
