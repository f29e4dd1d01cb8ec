"""Build the native libraries of wabbit_b200 in-tree (so the .so files travel with the repo snapshot).

  libwabbit_gpu.so   CUDA kernels + C ABI (include/wabbit_gpu.h), nvcc, sm_100a only
  libwabbit_host.so  host forest metadata (include/wabbit_host.h), g++
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")

GPU_LIB = os.path.join(HERE, "libwabbit_gpu.so")
HOST_LIB = os.path.join(HERE, "libwabbit_host.so")

GPU_SRCS = ["capi.cu", "kernels.cu", "wavelet.cu", "jump.cu", "topology.cu", "multigpu.cu", "statistics.cu", "krylov.cu"]
HOST_SRCS = ["host_forest.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--threads", "6", "-ldl"]


def _stale(out: str, srcs) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = list(srcs) + [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)] + \
        [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build_host(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in HOST_SRCS]
    if force or _stale(HOST_LIB, srcs):
        tmp = f"{HOST_LIB}.{os.getpid()}.tmp"      # several ranks may get here at once: build aside, then rename atomically
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-I", INCLUDE, "-o", tmp, *srcs])
        os.replace(tmp, HOST_LIB)
    return HOST_LIB


def build_gpu(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in GPU_SRCS]
    if force or _stale(GPU_LIB, srcs):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        tmp = f"{GPU_LIB}.{os.getpid()}.tmp"
        subprocess.check_call([nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-o", tmp, *srcs])
        os.replace(tmp, GPU_LIB)
    return GPU_LIB


def build_all(force: bool = False) -> None:
    build_host(force)
    build_gpu(force)


if __name__ == "__main__":
    build_all(force=True)
    print(GPU_LIB)
    print(HOST_LIB)
