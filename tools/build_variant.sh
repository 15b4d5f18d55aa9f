#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. ..." : libanmodem with other compile-time choices in the dense kernel -> tools/_variants/libanmodem_NAME.so
# (experiments only: ANM_LIB_PATH points the Python binding at it)
set -e
cd "$(dirname "$0")/../audio-network_b200/csrc"
mkdir -p ../../tools/_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $2 -c anm_cuda.cu -o /tmp/anm_cuda_$1.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/_variants/libanmodem_$1.so /tmp/anm_cuda_$1.o $(ls *_cu.o *_c.o | grep -v anm_cuda_cu.o) -lpthread -lm
cuobjdump -res-usage ../../tools/_variants/libanmodem_$1.so 2>/dev/null | grep -A1 "k_demod_tcILi64ELi256ELi4ELi0" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | paste - - | sed "s/^/$1: /"
