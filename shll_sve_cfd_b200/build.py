"""Build libshll_b200.so in-tree with nvcc for sm_100a (no torch involved: plain C ABI, static cudart)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libshll_b200.so")
SOURCES = ["shll_capi.cu", "shll_group.cu", "step1d.cu", "step2d_o1.cu", "step2d_o2_strict.cu", "step2d_o2_fast.cu", "step2d_acc.cu", "step2d_acc_o2.cu", "selftest.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: contraction is never left to the compiler (STRICT must not fuse; FAST fuses explicitly with __fmaf_rn).
# -ftz=false -prec-div=true -prec-sqrt=true are the defaults, spelled out because STRICT mode depends on them.
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-fmad=false", "-ftz=false", "-prec-div=true", "-prec-sqrt=true",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _deps_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(os.path.abspath(__file__)))


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    dep = _deps_mtime()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= dep:
        return LIB

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= dep:
            return obj
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print("[build]", " ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-lpthread"]
    if verbose:
        print("[build]", " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
