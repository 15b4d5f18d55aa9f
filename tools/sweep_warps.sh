python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in ${WLIST:-17 18 19}; do ANM_WARPS=$w python bench.py --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('W',d['config']['launch']['warps_per_cta'], d['ms_per_step'], d['roofline']['frac'], d['frames_ok'])"; done
