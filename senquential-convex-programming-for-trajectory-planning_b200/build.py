"""Build libscpb200.so in-tree with nvcc for sm_100a (the only target)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libscpb200.so")
SOURCES = ["scpb200.cu"]
DEPS = ["scp_common.cuh", "ipm_core.cuh", "ops_pair.cuh", "ops_dense.cuh", "scp_kernels.cuh", "scpb200.cu",
        os.path.join("..", "..", "include", "scpb200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-shared", "--use_fast_math=false"]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
        cmd = ["nvcc"] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + \
              [os.path.join(CSRC, s) for s in SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
