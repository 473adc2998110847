"""Tuning aid: |u - u*|_inf of K4 (teacher-forced, one QP per SCP iteration of every golden step) against the
extended-precision minimisers in tests/golden, for combinations of qp_dual_reg and qp_dres_floor_factor."""
import ctypes as C, glob, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
PKG = "senquential-convex-programming-for-trajectory-planning_b200"
capi = importlib.import_module(PKG + "._capi"); batch = importlib.import_module(PKG + ".batch")
files = sorted(f for f in glob.glob(os.path.join(ROOT, "tests", "golden", "circle*_step*.npz")) if "hp50" not in f)
for dreg in (1e-12, 1e-11, 1e-10, 1e-9):
    for fac in (0, 10, 100):
        worst, its, flo = 0.0, 0, 0
        for f in files:
            G = dict(np.load(f)); nit = int(G["scp_iters"]); nVeh, Hp = int(G["sc_nVeh"]), int(G["sc_Hp"])
            p = capi.Params(); capi.load().scpb200_default_params(C.byref(p))
            p.dt, p.uLim, p.dsafeExtra, p.max_scp_iter = float(G["sc_dt"]), float(G["sc_uLim"]), float(G["sc_dsafeExtra"]), 1
            p.qp_dual_reg, p.qp_dres_floor_factor = dreg, fac
            bs = batch.BatchSCP(nit, nVeh, Hp, params=p)
            veh = np.stack([G["sc_Lf"], G["sc_Lr"], G["sc_Q"], G["sc_Q_final"], G["sc_R"]], axis=1)
            rep = lambda a: np.repeat(a[None], nit, axis=0)
            bs.load_inputs(x0=rep(G["x0"]), u0=rep(G["u0"].reshape(nVeh)), veh=rep(veh), poly=rep(G["sc_poly"]),
                           dsafe=rep(G["sc_dsafeVehicles"]), u=G["prev_u"][:nit])
            bs.controller_step(); torch.cuda.synchronize()
            u = bs.u.cpu().numpy()
            worst = max(worst, max(np.abs(u[i] - G["x"][i][:-1]).max() for i in range(nit)))
            its += int(bs.ipm_iters.sum()); flo += int(((bs.status.cpu().numpy() & 32) > 0).sum())
        print(f"dual_reg {dreg:.0e} floor factor {fac:3d}: worst |u-u*| {worst:.2e}, ipm iterations {its}, floor acceptances {flo}")
