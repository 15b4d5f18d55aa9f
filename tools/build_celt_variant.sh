#!/bin/bash
# build_celt_variant.sh NAME "-DFLAG=.. ..." : libanmodem with other compile-time choices in the CELT kernels -> tools/_variants/libanmodem_celt_NAME.so
# (experiments only: ANM_LIB_PATH points the Python binding at it; tools/r2_celt_ab.sh runs every such build on one box)
set -e
cd "$(dirname "$0")/../audio-network_b200/csrc"
mkdir -p ../../tools/_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include $2 -c anm_celt_gpu.cu -o /tmp/anm_celt_$1.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/_variants/libanmodem_celt_$1.so /tmp/anm_celt_$1.o $(ls *_cu.o *_c.o | grep -v anm_celt_gpu_cu.o) -lpthread -lm
cuobjdump -res-usage ../../tools/_variants/libanmodem_celt_$1.so 2>/dev/null | grep -A1 "k_celt_spectrum" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | paste - - | sed "s/^/$1: /"
