"""In-tree build of libhsenet_sm100a.so with nvcc (sm_100a only; cross-compiles without a GPU).

    python -m hsenet_b200.build [--force] [--verbose]

Objects go to hsenet_b200/csrc/_build/, the shared library to hsenet_b200/libhsenet_sm100a.so (git-ignored,
but shipped to the GPU box by gpurun).  The library links the static CUDA runtime only; cuTensorMapEncodeTiled is
resolved at run time through cudaGetDriverEntryPoint, so no libcuda is needed at build time.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# HSENET_BUILD_TAG=<tag> builds a variant (e.g. the trace build) beside the product library: objects in _build_<tag>/,
# library libhsenet_sm100a_<tag>.so; select it at run time with HSENET_LIB_PATH.
_TAG = os.environ.get("HSENET_BUILD_TAG", "")
BUILD = os.path.join(CSRC, "_build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(HERE, "libhsenet_sm100a" + ("_" + _TAG if _TAG else "") + ".so")
SOURCES = ["api.cu", "gemm_tcgen05.cu", "gemm_tcgen05_2cta.cu", "attention_tcgen05.cu", "attention_bwd_tcgen05.cu", "patch_embed_tcgen05.cu", "backward.cu", "train.cu", "slice_trunk.cu", "rowops.cu", "ingest.cu", "verify_fp32.cu"]
HEADERS = ["common.cuh", "kernels.h", "gemm_epilogue.cuh", "composite.cuh", os.path.join("..", "..", "include", "hsenet_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
] + os.environ.get("HSENET_NVCC_EXTRA", "").split()     # e.g. -DHSENET_ATT_TRACE for tools/attn_trace.py


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _stale(target, deps):
    t = _mtime(target)
    return t == 0.0 or any(_mtime(d) > t for d in deps)


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    return r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(BUILD, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append([NVCC] + FLAGS + ["-c", src, "-o", obj])
    logs = []
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            logs = list(ex.map(lambda c: _run(c, verbose), jobs))
    if force or jobs or _stale(LIB, objs):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"], verbose)
    if logs:
        with open(os.path.join(BUILD, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
