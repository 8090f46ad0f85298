"""Build libsgv3d_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so must travel with
the repository snapshot to the GPU box).  ``python -m sgv3d_b200.build [--force] [--verbose]``."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libsgv3d_b200.so")
SOURCES = ["common.cu", "geometry.cu", "voxel_pooling.cu", "lift_splat.cu", "lift_splat_block.cu"]
HEADERS = ["common.cuh", "geometry.cuh", "sort.cuh", "transpose.cuh", "ls_shared.cuh", "ls_block.cuh",
           "../../include/sgv3d_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-cudart", "static",
              "--fmad=false"]
# --fmad=false: every multiply-add that may fuse is written as an explicit __fmaf_rn; nothing in
# this library relies on implicit contraction, and the geometry must never be contracted.


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, "_build", src.replace(".cu", ".o"))
        os.makedirs(os.path.dirname(o), exist_ok=True)
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
               "-o", LIB, *objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
