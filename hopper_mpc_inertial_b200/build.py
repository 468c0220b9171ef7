"""In-tree build of libhmpc_b200.so with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(_HERE, "csrc", "hmpc_api.cu")]
DEPS = [os.path.join(_HERE, "csrc", f) for f in ("hmpc_api.cu", "hmpc_qp.cuh", "hmpc_mpc.cuh", "hmpc_sim.cuh")] + \
       [os.path.join(_HERE, "..", "include", "hmpc.h")]
OUT = os.path.join(_HERE, "libhmpc_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_lib(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + SRC + ["-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
