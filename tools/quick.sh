timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --e2e-steps 0 --no-cfg4 --steps 60 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ref4', d['ms_per_step'], d['roofline']['frac'], d['frames_ok'])"
python bench.py --preset wide64 --channels 4096 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('wide64', d['ms_per_step'], d['value'], d['roofline']['frac'], d['frames_ok'])"
