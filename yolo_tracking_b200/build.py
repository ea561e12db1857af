"""Build libb200track.so in-tree with nvcc for sm_100a (no torch dependency).

`python -m yolo_tracking_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles
without a GPU.  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libb200track.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = glob.glob(os.path.join(CSRC, "*")) + [os.path.join(os.path.dirname(HERE), "include", "b200track.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def _compile_one(args):
    nvcc, src, obj, verbose = args
    cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", obj, src]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=CSRC)


def build(force: bool = False, verbose: bool = True):
    """One object per .cu (compiled in parallel, rebuilt when the source or any header is newer), then one link."""
    if not force and not needs_build():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(HERE, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [p for p in glob.glob(os.path.join(CSRC, "*")) if not p.endswith(".cu")]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "b200track.h"))
    hdr_t = max(os.path.getmtime(p) for p in headers)
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((nvcc, src, obj, verbose))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(_compile_one, jobs))
    cmd = [nvcc, "-shared", "-o", OUT] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=CSRC)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
