# multi-GPU check: torchrun bench at N GPUs (gloo frame gather inside the e2e leg) + the multi-device C handle on real devices
set -x
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi" > gpurun_out/r2_pytest_multi_n$N.log 2>&1; tail -3 gpurun_out/r2_pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --e2e-steps ${E2E:-2} > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n$N.json') if l.startswith('{')][-1])
print('N',d['n_gpus'],'value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],d['roofline']['frac_of_step_time'])
print('sustained',d.get('sustained',{}).get('value'),d.get('sustained',{}).get('clocks'))
print('e2e',json.dumps(d['e2e']))
PY
