"""Scenario sharding over the GPUs of one box.

Every scenario / noise sample is an independent unit and the SCP loop has no cross-instance dependence, so the
solve needs NO collective: rank r of G owns the contiguous slice [r*B/G, (r+1)*B/G) of the global batch, one process
per GPU (torch.distributed, backend "nccl" on the GPUs, "gloo" in the CPU tests).  Instance-keyed inputs and noise
streams (scenarios.circle_batch(instance0=...), params.instance0) make an instance's result independent of G.

The only communication is optional and off the critical path: an all-gather of trajectories / controls / status
for whoever wants the whole batch in one place, and an all-reduce of a handful of Monte-Carlo statistics.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(B_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of the global batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(B_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_batch(local: torch.Tensor, B_global: int) -> torch.Tensor:
    """Concatenate the per-rank slices [b_lo:b_hi, ...] of a batch-leading tensor in global instance order.

    Slices may differ in length by one; they are padded to a common length for the collective and trimmed after."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(B_global, r, world) for r in range(world)]
    maxlen = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxlen,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def gather_results(bs, B_global: int) -> Dict[str, torch.Tensor]:
    """All-gather U, traj and the per-instance bookkeeping of a BatchSCP (one call per MPC step at most)."""
    return {k: all_gather_batch(getattr(bs, k), B_global) for k in ("U", "traj", "scp_iters", "ipm_iters", "status", "obj",
                                                                  "max_violation")}


def monte_carlo_stats(bs) -> Dict[str, float]:
    """All-reduced statistics of the current step: QPs solved, infeasible instances, mean / variance of the cost."""
    dev = bs.obj.device
    obj = bs.obj.double()
    v = torch.stack([bs.scp_iters.sum().double(), ((bs.status & 8) != 0).sum().double(), obj.sum(), (obj * obj).sum(),
                     torch.tensor(float(bs.B), dtype=torch.float64, device=dev), bs.ipm_iters.sum().double()])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    n = float(v[4])
    mean = float(v[2]) / n
    return {"qps": float(v[0]), "infeasible": float(v[1]), "cost_mean": mean, "cost_var": float(v[3]) / n - mean * mean,
            "instances": n, "ipm_iterations": float(v[5])}
