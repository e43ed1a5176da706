"""Builds libeorb_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libeorb_b200.so")
SOURCES = ["capi.cu", "capi_lk.cu", "orb_kernels.cu", "orb_fast.cu", "orb_tiles.cu", "match_kernels.cu", "hamming_tc.cu", "event_kernels.cu", "lk_kernels.cu", "guided_kernels.cu", "capi_guided.cu", "bow_kernels.cu", "capi_bow.cu"]
HEADERS = ["eorb_math.cuh", "tma_utils.cuh", "fast_score.cuh", "octree_core.cuh", "orb_plan.h", "orb_kernels.h", "match_kernels.h", "event_kernels.h", "lk_kernels.h", "guided_kernels.h", "bow_kernels.h",
           "brief_pattern_31.inc", os.path.join("..", "..", "include", "eorb_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=true",
              "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=default", "-shared", "-cudart", "static", "-ldl"]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _obj_stale(src: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(CSRC, f) for f in HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """one object per translation unit (compiled in parallel, rebuilt only when it or a header changed), then one link"""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f not in ("-shared", "-ldl")]
    cflags = [f for i, f in enumerate(cflags) if not (f == "static" or f == "-cudart")]

    def compile_one(s):
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s + ".o")
        if not force and not _obj_stale(src, obj):
            return obj, 0, ""
        r = subprocess.run([nvcc] + cflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src], capture_output=True, text=True)
        return obj, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for obj, rc, out in results:
        if verbose or rc != 0:
            sys.stderr.write(out)
    if any(rc != 0 for _, rc, _ in results):
        raise RuntimeError("nvcc failed building libeorb_b200.so")
    r = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", LIB] + [o for o, _, _ in results], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libeorb_b200.so")
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
