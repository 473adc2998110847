"""World-size-2 test of the scenario sharding logic on CPU (gloo): the solve itself needs no collective; what is
checked here is that shards tile the global batch, that instance-keyed inputs do not depend on the sharding, and
that the optional all-gather / all-reduce reassemble results in global instance order."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = "senquential-convex-programming-for-trajectory-planning_b200"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeBatch:
    """Stands in for BatchSCP's result tensors (the solve is covered by the GPU tests)."""

    def __init__(self, lo, hi, Hp=10, nVeh=8):
        B = hi - lo
        ids = torch.arange(lo, hi, dtype=torch.float64)
        self.B = B
        self.U = ids[:, None, None].expand(B, Hp, nVeh).contiguous()
        self.traj = ids[:, None, None, None].expand(B, Hp, 2, nVeh).contiguous() * 2
        self.scp_iters = (torch.arange(lo, hi) % 5 + 1).to(torch.int32)
        self.ipm_iters = self.scp_iters * 13
        self.status = ((torch.arange(lo, hi) % 7) == 0).to(torch.int32) * 8
        self.obj = ids * 0.5
        self.max_violation = torch.zeros(B, dtype=torch.float64)


def _worker(rank, world, port, B_global, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = importlib.import_module(PKG + ".parallel")
    scen = importlib.import_module(PKG + ".scenarios")
    lo, hi = par.shard_range(B_global, rank, world)
    cb = scen.circle_batch(hi - lo, instance0=lo)
    ref = scen.circle_batch(B_global)
    ok_inputs = bool(np.array_equal(cb.x0, ref.x0[lo:hi]) and np.array_equal(cb.dsafe, ref.dsafe[lo:hi]))
    fb = _FakeBatch(lo, hi)
    g = par.gather_results(fb, B_global)
    full = _FakeBatch(0, B_global)
    ok_gather = all(torch.equal(g[k], getattr(full, k)) for k in g)
    st = par.monte_carlo_stats(fb)
    ok_stats = (st["qps"] == float(full.scp_iters.sum()) and st["instances"] == B_global
                and abs(st["cost_mean"] - float(full.obj.mean())) < 1e-12 and st["infeasible"] == float((full.status != 0).sum()))
    q.put((rank, lo, hi, ok_inputs, ok_gather, ok_stats))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B_global", [10, 13])
def test_two_rank_sharding(B_global):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B_global, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == B_global       # shards tile the batch
    for _, _, _, ok_inputs, ok_gather, ok_stats in res:
        assert ok_inputs and ok_gather and ok_stats


def test_shard_range_properties():
    par = importlib.import_module(PKG + ".parallel")
    for B in (0, 1, 7, 1024, 4096, 65537):
        for G in (1, 2, 4, 8):
            r = [par.shard_range(B, k, G) for k in range(G)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
