"""Builds wgmath_b200/libwgebra_b200.so in-tree with nvcc for sm_100a (no torch involved).

    python -m wgmath_b200.build [--force] [--verbose]

One object per .cu (compiled in parallel), linked into one shared library with a static CUDA
runtime, so the .so that travels to the GPU box depends only on libcuda / libdl / libstdc++.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libwgebra_b200.so")
SOURCES = ["abi.cu", "level1.cu", "scan_sort.cu", "geometry.cu", "gemv.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm.cu", "comm.cu"]
# the tcgen05 kernel variants, one operand family per translation unit so they compile in parallel (gemm_tc_kernel.cuh)
SOURCES += sorted(f for f in os.listdir(CSRC) if f.startswith("gemm_tc_inst_") and f.endswith(".cu"))
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# /usr/bin/g++ explicitly: this image exports CXX=/opt/gcc/bin/g++ (an incomplete toolchain).
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", HOST_CXX,
                     "--expt-relaxed-constexpr", "-Xptxas", "-v"]


# per-file flags.  geometry.cu: no FMA contraction, so the factorizations evaluate exactly the operation sequence of the WGSL
# (and of oracle/geometry_oracle.c, built with -ffp-contract=off); the only fused operations are the WGSL's own fma() calls.
EXTRA_FLAGS = {"geometry.cu": ["-fmad=false"]}


def nvcc() -> str:
    for cand in (os.environ.get("WGB_NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libwgebra_b200.so cannot be built")


def _deps(src: str):
    yield os.path.join(CSRC, src)
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h")):
            yield os.path.join(CSRC, f)
    yield os.path.join(HERE, "..", "include", "wgb200.h")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    cmd = [nvcc()] + NVCC_FLAGS + EXTRA_FLAGS.get(src, []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    if verbose:
        print(log)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    todo = [s for s in SOURCES if force or _stale(os.path.join(OBJ_DIR, s.replace(".cu", ".o")), _deps(s))]
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in SOURCES]
    if todo or _stale(LIB, objs):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-ccbin", HOST_CXX, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
