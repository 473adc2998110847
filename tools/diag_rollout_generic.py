"""Rollout entry vs the per-step call sequence on shapes served by the run-time-dimension kernels (with and without
steering-rate rows): prints the largest difference per output (the fixed-shape kernel is covered bit for bit by
tests/test_gpu_workloads.py)."""
import ctypes as C, importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "senquential-convex-programming-for-trajectory-planning_b200"
capi = importlib.import_module(PKG + "._capi"); batch = importlib.import_module(PKG + ".batch"); scen = importlib.import_module(PKG + ".scenarios")
for nVeh, Hp, rate in ((8, 10, 1), (3, 10, 0), (3, 10, 1), (8, 12, 0)):
    B, nsteps = 24, 4
    cb = scen.circle_batch(B, nVeh=nVeh, Hp=Hp, instance0=0, step_lo=4, step_hi=7)
    def fresh():
        p = capi.Params(); capi.load().scpb200_default_params(C.byref(p))
        p.enable_rate_rows, p.duLim = rate, 0.15 * scen.DU_LIM
        bs = batch.BatchSCP(B, nVeh, Hp, params=p)
        bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((B, nVeh * Hp)))
        return bs
    a = fresh()
    for s in range(nsteps):
        a.params.noise_counter = s
        a.setup(); a.solve(); a.advance_linear(scen.MECH_LIMIT, scen.DU_LIM)
        if s == 0:
            first = {k: getattr(a, k).clone() for k in ("g", "H", "qv", "gamma0", "u", "U", "scp_iters")}
    r = fresh(); r.params.noise_counter = 0
    r1 = fresh(); r1.params.noise_counter = 0; r1.rollout(1, scen.MECH_LIMIT, scen.DU_LIM)
    r.rollout(nsteps, scen.MECH_LIMIT, scen.DU_LIM)
    torch.cuda.synchronize()
    print(f"nVeh {nVeh} Hp {Hp} rate {rate} plan {a.plan()}")
    print("   after 1 step :", {k: float((first[k].double() - getattr(r1, k).double()).abs().max()) for k in first})
    print("   after", nsteps, "steps:", {k: float((getattr(a, k).double() - getattr(r, k).double()).abs().max()) for k in ("u", "U", "x0", "u0", "scp_iters", "ipm_iters", "status")})
