# ref4 leg of bench.py for chunk lengths / warps per CTA (experiment knobs: ANM_BENCH_CHUNK_SYMS, ANM_WARPS)
for cs in ${CSLIST:-344 352}; do for w in ${WLIST:-20 19}; do
ANM_BENCH_CHUNK_SYMS=$cs ANM_WARPS=$w python bench.py --no-cpu-baseline --e2e-steps 0 --no-cfg4 --steps 60 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk_syms', $cs, 'W', d['config']['launch']['warps_per_cta'], d['ms_per_step'], d['value'], d['roofline']['frac'], d['frames_ok'])"
done; done
