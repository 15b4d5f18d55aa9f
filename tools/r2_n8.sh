set -x
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
nproc >> gpurun_out/r2_topo_n$N.txt; numactl -H >> gpurun_out/r2_topo_n$N.txt 2>&1; lscpu | grep -i numa >> gpurun_out/r2_topo_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/stream_cfg3.py --out gpurun_out/r2_cfg3_as_written_n$N.json > gpurun_out/r2_cfg3_n$N.log 2>&1; echo "cfg3 rc=$?"
tail -2 gpurun_out/r2_cfg3_n$N.log | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/stream_cfg3.py --realtime-factor 40 --out gpurun_out/r2_cfg3_paced_n$N.json > gpurun_out/r2_cfg3_paced_n$N.log 2>&1; echo "cfg3 paced rc=$?"
tail -1 gpurun_out/r2_cfg3_paced_n$N.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 2 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n$N.json') if l.startswith('{')][-1])
print('N',d['n_gpus'],'value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],d['roofline']['frac_of_step_time'])
print('sustained',d.get('sustained',{}).get('value'),d.get('sustained',{}).get('clocks'))
print('e2e',json.dumps(d['e2e']))
PY
