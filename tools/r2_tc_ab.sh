# A/B on one box: the committed dense kernel against compile-time variants in tools/_variants/ (tools/build_variant.sh)
mkdir -p gpurun_out
run() { # name, lib
  for i in 1 2; do
    env ${2:+ANM_LIB_PATH=$2} timeout 600 python bench.py --preset wide64 --channels 4736 --steps 5 --warmup 2 --e2e-steps 0 --no-cpu-baseline --no-sustain --no-cfg4 2> gpurun_out/r2_bench_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$1] kernel ms',d['roofline']['avg_kernel_ms'],'frames',d['frames_ok'])"
  done
}
run default ""
for f in tools/_variants/libanmodem_*.so; do n=${f##*libanmodem_}; run ${n%.so} $PWD/$f; done
run default ""
