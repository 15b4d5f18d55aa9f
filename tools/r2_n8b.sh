# scaling bench only (the configs[2]-as-written runs of tools/r2_n8.sh are not repeated): torchrun bench at N GPUs
N=${N:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 2 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n$N.json') if l.startswith('{')][-1])
print('N',d['n_gpus'],'value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],d['roofline']['frac_of_step_time'])
print('sustained',d.get('sustained',{}).get('value'),d.get('sustained',{}).get('clocks'))
print('e2e',json.dumps(d['e2e'])[:600])
PY
