"""Tuning aid (needs a library built with -DSCP_DIAG_LOG): where the interior-point method stops on the QPs that reach
the iteration cap.  python tools/diag_ipm_stop.py --hp 50 --batch 64 [--trust 0.2]"""
import argparse, ctypes as C, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
PKG = "senquential-convex-programming-for-trajectory-planning_b200"
capi = importlib.import_module(PKG + "._capi"); batch = importlib.import_module(PKG + ".batch"); scen = importlib.import_module(PKG + ".scenarios")
ap = argparse.ArgumentParser()
ap.add_argument("--hp", type=int, default=50); ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--trust", type=float, default=0.0); ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--msi", type=int, default=20); ap.add_argument("--dual-reg", dest="dreg", type=float, default=0.0)
a = ap.parse_args()
cb = scen.circle_batch(a.batch, Hp=a.hp, step_lo=4, step_hi=7)
p = capi.Params(); capi.load().scpb200_default_params(C.byref(p)); p.max_scp_iter = a.msi
if a.trust > 0: p.trust_radius = a.trust * p.uLim
if a.dreg > 0: p.qp_dual_reg = a.dreg
bs = batch.BatchSCP(a.batch, 8, a.hp, params=p)
bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((a.batch, 8 * a.hp)))
for s in range(a.steps):
    bs.controller_step(); torch.cuda.synchronize()
    log = bs.log.cpu().numpy(); scp = bs.scp_iters.cpu().numpy()
    rows = np.concatenate([log[b, :scp[b]] for b in range(a.batch)])
    bad = rows[rows[:, 9].astype(int) & 1 == 1]
    print(f"step {s}: QPs {len(rows)}, accepted at the dual-residual floor {(rows[:, 9].astype(int) & 32 > 0).sum()}, at the cap {len(bad)}, pivot repairs {(rows[:, 9].astype(int) & 2 > 0).sum()}, ipm/QP {rows[:, 8].mean():.1f}")
    for r in bad[:8]:
        print(f"   iters {int(r[8])} gap {r[3]:.2e} relgap {r[7]:.2e} pres {r[6]:.2e} dres {r[5]:.2e} slack {r[0]:.2e} status {int(r[9])}")
    ok = rows[rows[:, 9].astype(int) & 1 == 0]
    if len(ok): print(f"   converged QPs: relgap median {np.median(ok[:, 7]):.1e} max {ok[:, 7].max():.1e}; iters max {int(ok[:, 8].max())}")
    bs.advance_linear(scen.MECH_LIMIT, scen.DU_LIM)
