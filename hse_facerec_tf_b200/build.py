"""Builds hse_facerec_tf_b200/libhfr.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhfr.so")
CU_SOURCES = ["launch.cu", "api.cu", "mtcnn.cu"]
CC_SOURCES = ["graphdef.cc", "compiler.cc", "h5keras.cc"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-Xcompiler", "-fPIC,-Wall", "--use_fast_math=false"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "hfr.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in CU_SOURCES + CC_SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src + ".o")
        objs.append(obj)
        if not force and _newer(obj, [path] + headers):
            continue
        if src.endswith(".cu"):
            cmd = [nvcc] + [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + ["-I", CSRC, "-c", path, "-o", obj]
        else:
            cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-Wall", "-I", CSRC, "-c", path, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    if force or not _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-lpthread", "-ldl", "-lrt"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
