"""Per-instance results must not depend on the work-queue policy (which CTA solves which QP, in which order):
solve the same 1024-instance batch under several policies / launch shapes and compare bit for bit."""
import ctypes as C, importlib, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
PKG = "senquential-convex-programming-for-trajectory-planning_b200"
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    capi = importlib.import_module(PKG + "._capi"); batch = importlib.import_module(PKG + ".batch"); scen = importlib.import_module(PKG + ".scenarios")
    B = 1024
    cb = scen.circle_batch(B, Hp=10, step_lo=5, step_hi=7)
    p = capi.Params(); capi.load().scpb200_default_params(C.byref(p))
    bs = batch.BatchSCP(B, 8, 10, params=p)
    bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((B, 80)))
    outs = []
    for s in range(3):
        bs.controller_step(); torch.cuda.synchronize()
        outs += [bs.u.cpu().numpy().copy(), bs.ipm_iters.cpu().numpy().copy(), bs.scp_iters.cpu().numpy().copy()]
        bs.advance_linear(scen.MECH_LIMIT, scen.DU_LIM)
    np.savez(sys.argv[2], *outs)
    sys.exit(0)
ref = None
for name, env in [("default", {}), ("pinned 0", {"SCPB200_PINNED": "0"}), ("pinned all", {"SCPB200_PINNED": "444"}),
                  ("quantum 3", {"SCPB200_QUANTUM": "3"}), ("1 CTA per SM", {"SCPB200_CTAS_PER_SM": "1"}),
                  ("2 CTAs per SM", {"SCPB200_CTAS_PER_SM": "2"}),
                  ("256 threads (another reduction order: differences at rounding level are expected)", {"SCPB200_THREADS": "256"})]:
    out = f"/tmp/det_{len(name)}_{abs(hash(name)) % 1000}.npz"
    subprocess.run([sys.executable, __file__, "child", out], env={**os.environ, **env}, check=True)
    D = np.load(out); arrs = [D[k] for k in D.files]
    if ref is None:
        ref = arrs; print(f"{name}: reference, ipm total {sum(int(a.sum()) for a in arrs[1::3])}"); continue
    same = all(np.array_equal(a, b) for a, b in zip(arrs, ref))
    worst = max(float(np.abs(a - b).max()) for a, b in zip(arrs[0::3], ref[0::3]))
    nd = [int((a != b).any(axis=-1).sum()) if a.ndim > 1 else int((a != b).sum()) for a, b in zip(arrs, ref)]
    print(f"{name}: identical={same}, max |du| {worst:.3e}, differing instances per array {nd}")
