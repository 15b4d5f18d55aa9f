# bench the ref4 leg with variant builds of the library (ANM_LIB_PATH); build a variant as audio-network_b200/libanmodem_<name>.so, then VARIANTS="default <name>" bash tools/variants.sh
for m in ${VARIANTS:-default}; do
if [ "$m" = default ]; then unset ANM_LIB_PATH; else export ANM_LIB_PATH=$PWD/audio-network_b200/libanmodem_$m.so; fi
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "bit_exact_clean or golden or noisy" 2>&1 | tail -1
for i in 1 2; do python bench.py --no-cpu-baseline --e2e-steps 0 --no-cfg4 --steps 60 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', d['ms_per_step'], d['roofline']['frac'], d['frames_ok'])"; done
done
