"""In-tree build of the native code (no JIT cache: the .so files travel with the repo snapshot to the GPU box).

  lib/libbwa_b200.so     the product: CUDA engine (csrc/engine.cu, sm_100a) + index builder, C-ABI of include/bwamem_b200.h
  lib/libbwaidx_b200.so  the index builder alone (host g++), used by the data tools on boxes without nvcc
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
ENGINE_SO = os.path.join(LIB, "libbwa_b200.so")
IDX_SO = os.path.join(LIB, "libbwaidx_b200.so")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-Wno-deprecated-declarations", "-shared", "-Xlinker", "-Bsymbolic"]


def _newest_src() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC))


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the engine cannot be built (there is no CPU fallback)")
    return p


def build_engine(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    if not force and os.path.exists(ENGINE_SO) and os.path.getmtime(ENGINE_SO) >= _newest_src():
        return ENGINE_SO
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", ENGINE_SO, os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "sa_build.cu"), os.path.join(CSRC, "index_build.cpp"), os.path.join(CSRC, "host_scan.cpp"), "-lz"]
    subprocess.check_call(cmd, cwd=HERE)
    return ENGINE_SO


def build_index_tool(force: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    src = os.path.join(CSRC, "index_build.cpp")
    newest = max(os.path.getmtime(src), os.path.getmtime(os.path.join(CSRC, "bwa_image.h")))
    if not force and os.path.exists(IDX_SO) and os.path.getmtime(IDX_SO) >= newest:
        return IDX_SO
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", IDX_SO, src, "-lz", "-lpthread"])
    return IDX_SO


if __name__ == "__main__":
    build_index_tool(force="--force" in sys.argv)
    print(build_engine(force="--force" in sys.argv, verbose="-v" in sys.argv))
