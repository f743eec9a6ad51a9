"""Build libcgrt.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "cgrt_api.cu")
OUT = os.path.join(HERE, "libcgrt.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # fp64 results must replay the reference's un-contracted arithmetic bit for bit
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    srcdir = os.path.join(HERE, "csrc")
    deps = [os.path.join(srcdir, f) for f in os.listdir(srcdir)] + [os.path.join(HERE, "..", "include", "cgrt.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    extra = os.environ.get("CGRT_NVCC_EXTRA", "").split()  # dev: e.g. -DCGRT_TRACE_MINB=6 for launch-bound experiments
    out = os.environ.get("CGRT_BUILD_OUT", OUT)
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC]
    env = dict(os.environ)
    env.pop("CXX", None)  # the image's CXX wrapper lacks an OpenMP spec; nvcc should use g++ from PATH
    env.pop("CC", None)
    subprocess.check_call(cmd, env=env)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
