"""Build libscpb200.so in-tree with nvcc for sm_100a (the only target).

The library is ten translation units — the host API with the small kernels (scpb200.cu), the SCP kernel with run-time
dimensions (scp_solve_generic.cu, once for CTAs of up to 256 threads and once for up to 512) and with the literal
dimensions of the shapes BASELINE.json names (scp_solve_fixed.cu: 8 vehicles x Hp 10 at two CTA widths, 8 vehicles x
Hp 20 at two) — compiled in parallel and linked into one shared object."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libscpb200.so")
# object files are scratch: gpurun_out/ is neither tracked by git nor shipped to the GPU box (the linked .so is)
OBJ = os.path.join(os.path.dirname(HERE), "gpurun_out", "build_obj")
# (object name, source, extra flags)
UNITS = [("scpb200.o", "scpb200.cu", []),
         ("scp_solve_generic.o", "scp_solve_generic.cu", []),
         ("scp_solve_v8h10_256.o", "scp_solve_fixed.cu", ["-DSCP_FIXED_HP=10", "-DSCP_FIXED_NT=256"]),
         ("scp_solve_v8h10_128.o", "scp_solve_fixed.cu", ["-DSCP_FIXED_HP=10", "-DSCP_FIXED_NT=128"]),
         ("scp_solve_v8h20_256.o", "scp_solve_fixed.cu", ["-DSCP_FIXED_HP=20", "-DSCP_FIXED_NT=256"]),
         ("scp_solve_v8h20_512.o", "scp_solve_fixed.cu", ["-DSCP_FIXED_HP=20", "-DSCP_FIXED_NT=512"]),
         ("scp_solve_generic_wide.o", "scp_solve_generic.cu", ["-DSCP_GENERIC_WIDE"]),
         ("scp_rollout_generic.o", "scp_rollout_generic.cu", []),
         ("scp_rollout_generic_wide.o", "scp_rollout_generic.cu", ["-DSCP_GENERIC_WIDE"]),
         ("scp_rollout_v8h10_128.o", "scp_rollout_fixed.cu", [])]
DEPS = ["scp_common.cuh", "ipm_core.cuh", "ops_pair.cuh", "ops_dense.cuh", "scp_kernels.cuh", "scp_solve_kernel.cuh",
        "scpb200.cu", "scp_solve_generic.cu", "scp_solve_fixed.cu", "scp_rollout_generic.cu", "scp_rollout_fixed.cu", os.path.join("..", "..", "include", "scpb200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]


def needs_build(out: str = OUT) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """Compile (if stale) and return the path of the shared object.  `defines` (e.g. ["-DSCP_PHASE_TIMERS"]) with a
    different `out` builds a tuning variant next to the product library."""
    if not (force or needs_build(out)):
        return out
    tag = os.path.splitext(os.path.basename(out))[0]
    objdir = os.path.join(OBJ, tag)
    os.makedirs(objdir, exist_ok=True)
    procs = []
    headers = [os.path.join(CSRC, d) for d in DEPS if d.endswith((".cuh", ".h"))]
    for obj, src, extra in UNITS:
        # per-unit staleness: an object is rebuilt when its own source or any header is newer (objects of other variants,
        # i.e. other `defines`, live in their own directory)
        op = os.path.join(objdir, obj)
        if not force and os.path.exists(op) and not verbose:
            t = os.path.getmtime(op)
            if all(os.path.getmtime(f) <= t for f in headers + [os.path.join(CSRC, src)]):
                continue
        cmd = ["nvcc"] + NVCC_FLAGS + list(defines) + extra + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", "-o", os.path.join(objdir, obj), os.path.join(CSRC, src)]
        procs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    fail = []
    for obj, pr in procs:
        so, se = pr.communicate()
        if verbose:
            sys.stderr.write(f"---- {obj}\n{se}")
        if pr.returncode != 0:
            fail.append(f"{obj}:\n{so}{se}")
    if fail:
        raise RuntimeError("nvcc failed:\n" + "\n".join(fail))
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + \
          [os.path.join(objdir, obj) for obj, _, _ in UNITS]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a for a in sys.argv[1:] if a.endswith(".so")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else OUT, defines=defs))
