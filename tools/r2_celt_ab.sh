# A/B of two builds of the CELT kernels on one box (tools/_variants/libanmodem_*.so)
for i in 1 2; do for f in tools/_variants/libanmodem_celt*.so; do echo "$f: $(ANM_LIB_PATH=$PWD/$f timeout 600 python tools/celt_bench.py --streams 4096 --reps 3 2>&1 | tail -1 | cut -c150-280)"; done; done
