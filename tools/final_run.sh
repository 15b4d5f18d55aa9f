set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; tail -1 gpurun_out/bench_final.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1z.csv python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/launches_r1z.csv
