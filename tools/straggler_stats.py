"""Who sets the step time at 1024 instances: per MPC step of the bench workload, the instances with the longest chain of
interior-point iterations, with their per-QP iteration counts (log column 8) and status.

    python tools/straggler_stats.py [--batch 1024] [--steps 12]"""
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
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--step-lo", type=int, default=4)
ap.add_argument("--step-hi", type=int, default=7)
ap.add_argument("--top", type=int, default=4)
args = ap.parse_args()

cb = scen.circle_batch(args.batch, Hp=10, step_lo=args.step_lo, step_hi=args.step_hi)
p = capi.Params()
capi.load().scpb200_default_params(C.byref(p))
p.noise_sigma, p.seed, p.instance0 = 3e-6, 20261018, 0
bs = batch.BatchSCP(args.batch, 8, 10, params=p)
bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((args.batch, 80)))
dump = {}
for s in range(args.steps):
    bs.params.noise_counter = s
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    bs.setup(); e[0].record(); bs.solve(); e[1].record()
    torch.cuda.synchronize()
    ipm = bs.ipm_iters.cpu().numpy(); scp = bs.scp_iters.cpu().numpy(); st = bs.status.cpu().numpy()
    log = bs.log.cpu().numpy()
    ms = e[0].elapsed_time(e[1])
    print(f"step {s}: solve {ms:.2f} ms, QPs {scp.sum()}, ipm {ipm.sum()} ({ipm.sum() / scp.sum():.2f}/QP), max chain {ipm.max()}, "
          f"throughput-bound share {ipm.sum() / 296 / max(1, ipm.max()):.2f}, us/ipm on the longest chain {1e3 * ms / ipm.max():.1f}")
    hist = np.bincount(np.minimum(ipm // 50, 12))
    print("   chain-length histogram (bins of 50 ipm iterations):", hist.tolist())
    for b in np.argsort(-ipm)[: args.top]:
        per = log[b, : scp[b], 8].astype(int).tolist()
        print(f"   instance {b}: scp {scp[b]} ipm {ipm[b]} status {st[b]:#x} per-QP {per} slack {log[b, scp[b] - 1, 0]:.2e}")
    dump[f'perqp_{s}'] = log[:, :, 8].astype(np.int16); dump[f'log_{s}'] = log.astype(np.float32); dump[f'scp_{s}'] = scp; dump[f'ms_{s}'] = ms
    bs.advance_linear(scen.MECH_LIMIT, scen.DU_LIM)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'straggler_dump.npz'), **dump)
