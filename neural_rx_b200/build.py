"""Build ``libnrx_b200.so`` (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

``python -m neural_rx_b200.build`` or ``build_library()``; the shared object lands next to this
file so that it travels with the repository snapshot to the GPU box (no JIT cache involved).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnrx_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "nrx_engine.cu")]
HEADERS = [os.path.join(HERE, "csrc", "nrx_kernels.cuh"), os.path.join(HERE, "csrc", "sm100_prims.cuh"),
           os.path.join(HERE, "csrc", "nrx_stack.cuh"), os.path.join(HERE, "csrc", "nrx_stack_pair.cuh"),
           os.path.join(HERE, "csrc", "nrx_stack_tm.cuh"), os.path.join(HERE, "csrc", "nrx_stack_ws.cuh"), os.path.join(HERE, "csrc", "nrx_agg_ws.cuh"),
           os.path.join(HERE, "..", "include", "nrx_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libnrx_b200.so")
    return nvcc


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
