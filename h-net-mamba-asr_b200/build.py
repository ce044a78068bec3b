"""Build libhnet_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python h-net-mamba-asr_b200/build.py [--force]

One object per .cu (compiled in parallel, rebuilt only when a source/header is newer), then one
shared library next to the Python package so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "dcasr_b200", "libhnet_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", INCLUDE]


def _newest_header() -> float:
    hs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, force: bool) -> tuple[str, bool]:
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), _newest_header())
    if stale:
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, stale


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in res]
    if any(st for _, st in res) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-lcudart", "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"built {LIB} from {len(objs)} objects ({sum(st for _, st in res)} recompiled)")
    elif verbose:
        print(f"{LIB} is up to date")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
