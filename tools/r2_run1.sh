# round 2, GPU run 1: tests, bench (both arms), launch list, ncu full capture of k_demod
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
nproc >> gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench1.json
tail -5 gpurun_out/r2_bench1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_ref1.json 2> gpurun_out/r2_ref1.err; echo "ref rc=$?"
cat gpurun_out/r2_ref1.json
ANM_BENCH_CHUNKS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches1.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sustain > gpurun_out/r2_ncu_launch.log 2>&1
ANM_BENCH_CHUNKS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_demod --launch-skip 6 -c 1 -f -o gpurun_out/prof_r2a python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-cfg4 --no-sustain > gpurun_out/r2_ncu_a.log 2>&1
echo done
