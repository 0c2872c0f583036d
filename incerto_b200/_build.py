"""In-tree build of the CUDA engine: nvcc -> incerto_b200/lib/libincerto_b200.so.

Every translation unit is compiled for sm_100a only
(`-gencode arch=compute_100a,code=sm_100a -lineinfo`); nvcc cross-compiles
without a GPU.  Objects are cached by source mtime so rebuilds are incremental.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libincerto_b200.so")

SOURCES = [
    "engine.cu",
    "sampling.cu",
    "module_forest.cu",
    "module_pandemic.cu",
    "kernels_pandemic.cu",
    "module_spatial.cu",
    "kernels_spatial.cu",
    "module_traders.cu",
    "kernels_traders.cu",
    "module_air.cu",
    "kernels_air.cu",
    "comm.cu",
    "kernels_generic.cu",
    "kernels_coin.cu",
    "kernels_forest.cu",
    "kernels_grid.cu",
    "debug_lottery.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the engine has no non-CUDA build")


def _deps_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".hpp", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(verbose: bool = False, force: bool = False, extra_flags: list[str] | None = None,
          variant: str | None = None) -> str:
    """`variant` builds a tuning variant into lib/libincerto_b200.<variant>.so (selected with INCERTO_B200_LIB).
    Objects are cached per flag set: `extra_flags` without a `variant` get an object directory (and a library) named
    after a hash of the flags, so a sweep can never relink stale objects into the default library."""
    flags = NVCC_FLAGS + (extra_flags or [])
    if extra_flags and not variant:
        import hashlib
        variant = "f" + hashlib.sha1(" ".join(extra_flags).encode()).hexdigest()[:10]
    objdir = os.path.join(HERE, "build", variant) if variant else OBJDIR
    lib = os.path.join(LIBDIR, f"libincerto_b200.{variant}.so") if variant else LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    hdr_m = _deps_mtime()
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            jobs.append([nvcc, *flags, "-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose and (r.stdout or r.stderr):
            print(r.stdout + r.stderr, flush=True)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(lib):
        # libcudart is linked statically (nvcc default); NCCL is dlopen-ed at attach time
        run([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-lpthread"])
    return lib


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(path)
