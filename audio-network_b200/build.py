"""Builds libanmodem.so (sm_100a CUDA kernels + C-ABI host code) in-tree.

Usage: python audio-network_b200/build.py [--force]
The shared object is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libanmodem.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

CU = ["anm_cuda.cu", "anm_multi.cu", "anm_tx.cu", "anm_pb_gpu.cu", "anm_opus_gpu.cu", "anm_celt_gpu.cu"]
C = ["anm_config.c", "anm_tx.c", "anm_pb.c", "anm_pb_msgs.c", "anm_pacer.c", "anm_celt_tables.c"]
DEPS = ["anm_kernels.cuh", "anm_kernels_tc.cuh", "anm_internal.h", "anm_host_queue.h", "anm_pb_wire.h", "anm_celt_entropy.h", "anm_celt_vec.h", "anm_celt_synth.h", "../../include/anmodem.h", "../../include/anmodem_pb.h", "../../include/anmodem_opus.h"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=True):
    deps = [os.path.join(CSRC, d) for d in DEPS]
    objs = []
    for src in CU + C:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ("_cu.o" if src.endswith(".cu") else "_c.o"))
        objs.append(obj)
        if force or _newer(obj, [path] + deps):
            if src.endswith(".cu"):
                cmd = [NVCC] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-c", path, "-o", obj]
            else:
                cmd = ["gcc", "-O2", "-fPIC", "-std=gnu11", "-Wall", "-Wextra", "-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    if force or _newer(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lpthread", "-lm"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
