# round 2 evidence run on one B200: full GPU suite, both bench arms, launch list, ncu --set full of both demodulator kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log
tail -4 gpurun_out/r2_pytest_final.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref_final.json 2> gpurun_out/r2_ref_final.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_final.err
ANM_BENCH_CHUNKS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sustain > gpurun_out/r2_ncu_launch.log 2>&1
ANM_BENCH_CHUNKS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_demod --launch-skip 6 -c 1 -f -o gpurun_out/prof_r2b python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-cfg4 --no-sustain > gpurun_out/r2_ncu_b.log 2>&1
ANM_BENCH_CHUNKS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_demod_tc --launch-skip 2 -c 1 -f -o gpurun_out/prof_r2_tc_g python bench.py --preset wide64 --channels 4736 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sustain --no-cfg4 > gpurun_out/r2_ncu_tc_g.log 2>&1
ANM_BENCH_CHUNKS=4 timeout 600 ncu --set full --clock-control none -k regex:"k_celt_entropy|k_pb_deframe|k_opus_parse" --launch-skip 9 -c 3 -f -o gpurun_out/prof_r2_chain python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-cfg4 --no-sustain > gpurun_out/r2_ncu_chain.log 2>&1
echo done
