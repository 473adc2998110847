"""Small driver for profiling: a few controller steps of the bench workload at a chosen batch size.

    python tools/run_scp_once.py --batch 148 --steps 2 [--max-scp-iter 1] [--hp 10]
Prints per-step CUDA-event times of K1 / K4 and the iteration statistics."""
import argparse
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "senquential-convex-programming-for-trajectory-planning_b200"
capi = importlib.import_module(PKG + "._capi")
batch = importlib.import_module(PKG + ".batch")
scen = importlib.import_module(PKG + ".scenarios")

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=148)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--hp", type=int, default=10)
ap.add_argument("--max-scp-iter", dest="msi", type=int, default=20)
ap.add_argument("--step-lo", type=int, default=6)
ap.add_argument("--step-hi", type=int, default=7)
ap.add_argument("--assemble", action="store_true")
ap.add_argument("--warm", type=int, default=-1)
ap.add_argument("--warm-relgap", dest="wrg", type=float, default=-1.0)
ap.add_argument("--snap-min-iter", dest="smi", type=int, default=0)
ap.add_argument("--carry", type=int, default=0)
ap.add_argument("--warm-max-iter", dest="wmi", type=int, default=0)
ap.add_argument("--rollout", type=int, default=0, help="run the steps as ONE scpb200_mpc_rollout launch")
args = ap.parse_args()

cb = scen.circle_batch(args.batch, Hp=args.hp, step_lo=args.step_lo, step_hi=args.step_hi)
p = capi.Params()
capi.load().scpb200_default_params(C.byref(p))
p.max_scp_iter = args.msi
if args.warm >= 0:
    p.qp_warm_start = args.warm
p.qp_warm_min_iter = args.smi if args.smi else p.qp_warm_min_iter
p.qp_warm_carry = args.carry
if args.wmi > 0:
    p.qp_warm_max_iter = args.wmi
if args.wrg > 0:
    p.qp_warm_relgap = args.wrg
bs = batch.BatchSCP(args.batch, 8, args.hp, params=p)
bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((args.batch, 8 * args.hp)))
print("plan", bs.plan())
TOT_IPM = TOT_QP = 0
if args.rollout:
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record(); R = bs.rollout(args.steps, scen.MECH_LIMIT, scen.DU_LIM); e[1].record()
    torch.cuda.synchronize()
    TOT_QP, TOT_IPM = int(R["qp_total"].sum()), int(R["ipm_total"].sum())
    print(f"rollout of {args.steps} steps: {e[0].elapsed_time(e[1]):.3f} ms ({e[0].elapsed_time(e[1]) / args.steps:.3f} per step), QPs {TOT_QP}, IPM its {TOT_IPM}")
for s in range(0 if args.rollout else args.steps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); bs.setup(); e[1].record(); bs.solve(); e[2].record()
    torch.cuda.synchronize()
    qp, ipm = int(bs.scp_iters.sum()), int(bs.ipm_iters.sum())
    print(f"step {s}: setup {e[0].elapsed_time(e[1]):.3f} ms, solve {e[1].elapsed_time(e[2]):.3f} ms, QPs {qp}, IPM its {ipm}, "
          f"max ipm/instance {int(bs.ipm_iters.max())}, us per ipm-iteration-on-critical-path "
          f"{1e3 * e[1].elapsed_time(e[2]) / max(1, int(bs.ipm_iters.max())):.1f}")
    TOT_IPM += ipm; TOT_QP += qp
    bs.advance_linear(scen.MECH_LIMIT, scen.DU_LIM)
if args.assemble:
    out = bs.assemble_dense()
    torch.cuda.synchronize()
    print("assembled", out["A"].shape)
if os.environ.get("SCPB200_LIB", "").endswith("timers.so"):
    import ctypes
    arr = (ctypes.c_ulonglong * 32)()
    lib = capi.load()
    lib.scpb200_debug_read_timers(arr)
    names = {0: "misc/outside", 1: "form_normal", 2: "chol potrf", 3: "chol trsm", 4: "chol gemm", 5: "diag inverses", 6: "inv W", 7: "inv product",
             8: "solve", 9: "residual phase", 10: "pass rhs (w1, A'w1)", 11: "pass dz/ds/step",
             12: "form: forces", 13: "form: M_v(k)", 14: "form: omega+diag"}
    tot = sum(arr)
    nfac = int(bs.ipm_iters.sum()) + int(bs.scp_iters.sum())
    nfac = TOT_IPM + TOT_QP
    print(f"phase timers, all steps (cycles of thread 0 summed over CTAs); IPM iterations + QPs over all steps: {nfac}")
    for i in range(15):
        print(f"  {names[i]:24s} {100.0 * arr[i] / tot:5.1f}%   {arr[i] / 1e6:10.2f} Mcyc   {arr[i] / max(1, nfac):10.0f} cyc/factorisation")
