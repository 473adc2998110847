"""Time K2 (scpb200_assemble_dense) alone: B instances, CUDA events around each launch, L2 flushed in between.
    python tools/time_assemble.py [--batch 1024] [--hp 10] [--reps 20]"""
import argparse, importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "senquential-convex-programming-for-trajectory-planning_b200"
batch = importlib.import_module(PKG + ".batch")
scen = importlib.import_module(PKG + ".scenarios")
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--hp", type=int, default=10)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
cb = scen.circle_batch(a.batch, Hp=a.hp, step_lo=6, step_hi=7)
bs = batch.BatchSCP(a.batch, 8, a.hp)
bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((a.batch, 8 * a.hp)))
bs.setup()
out = bs.assemble_dense()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(a.reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); bs.assemble_dense_into(bs.u, out); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
n1, mc, n = bs.n1, bs.mc, bs.n
byt = a.batch * (8 * (n1 * n1 + n1 + mc * n1 + mc + 2 * n1) + 8 * (16 * 8 + 64 + n))
ts = np.array(ts)
print(f"K2 B={a.batch} Hp={a.hp}: median {np.median(ts):.1f} us  min {ts.min():.1f} us  bytes {byt}  "
      f"{byt / np.median(ts) / 1e3:.1f} GB/s median  ({100 * byt / np.median(ts) / 1e3 / 6536.7:.1f}% of 6536.7)")
