for m in 0xF 0x5 0x0; do
export ANM_LIB_PATH=$PWD/audio-network_b200/libanmodem_$m.so
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "bit_exact_clean or golden or noisy" 2>&1 | tail -1
for i in 1 2; do python bench.py --no-cpu-baseline --e2e-steps 0 --no-cfg4 --steps 60 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', d['ms_per_step'], d['roofline']['frac'], d['frames_ok'])"; done
done
