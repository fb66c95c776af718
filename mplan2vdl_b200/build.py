"""Builds libvdl_cuda.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libvdl_cuda.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def sources() -> list:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(HERE, "..", "include", "vdl_cuda.h")]
    return any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("VDL_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + sources()
    env = dict(os.environ)
    env["PATH"] = "/usr/bin:" + env.get("PATH", "")      # host compiler: the distro gcc
    subprocess.check_call(cmd, env=env)
    return SO


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
