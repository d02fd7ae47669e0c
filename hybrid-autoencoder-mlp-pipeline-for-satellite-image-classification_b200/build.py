"""Build libae_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python build.py [--force] [--verbose]

The library is the product's only compute path; the Python layer loads it with ctypes and
fails loudly if it is missing.  NCCL is linked dynamically (libnccl.so.2, the copy torch
already loaded wins at run time).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# AE_B200_BUILD_VARIANT=trace: developer build with per-CTA phase timestamps in the row GEMM (-DAE_TRACE), written to
# libae_b200_trace.so; the product library never contains that code
VARIANT = os.environ.get("AE_B200_BUILD_VARIANT", "")
OBJ = os.path.join(HERE, "build" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "libae_b200" + ("_" + VARIANT if VARIANT else "") + ".so")
SOURCES = ["api.cu", "simt_gemm.cu", "thin.cu", "elementwise.cu", "head.cu", "mlp.cu", "tma_gemm.cu", "rowgemm2.cu", "dense_tc.cu", "engine.cu", "dp.cu", "augment.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"] + (["-DAE_TRACE"] if VARIANT == "trace" else [])


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    srcp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(srcp), _deps_mtime()):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return obj, r.stderr if verbose else ""


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [o for o, _ in results]
    for _, log in results:
        if log:
            print(log)
    need_link = force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if need_link:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lnccl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
