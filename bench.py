#!/usr/bin/env python
"""bench.py — batched SCP QP solves/s on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--hp HP]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): the default 8-vehicle circle scenario (Hp = 10), 1024 perturbed instances per
GPU with process noise (Monte-Carlo), synthetic inputs from scenarios.circle_batch.  A "step" is the
controller stage of one closed-loop MPC step over the whole batch: K1 set-up -> K4 fused SCP solve (-> linear
advance).  `value` = QPs solved (sum over instances of SCP iterations) per second, inputs resident in HBM, timed
with CUDA events per step (L2 flushed between steps), max over ranks.  `e2e` = the same through the public API
with HOST buffers (pinned H2D of every input, D2H of every result, per step).  `cpu_baseline` / `--impl
reference` time the oracle port of the reference's CPU path on the box's host cores (bounded sample).

The product arm never touches oracle/: only the cpu_baseline leg and `--impl reference` import it.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "senquential-convex-programming-for-trajectory-planning_b200"

METRIC = "batched SCP QP solves/sec"
UNIT = "QP/s"


# ---------------------------------------------------------------------------------------------------- helpers
def algorithmic_flops_per_ipm_iteration(nVeh, Hp):
    """n1^3/3 (Cholesky) + 2 sum_rows nnz_r^2 (A'DA) + 8 n1^2 (two RHS) + 8 nnz(G) (mat-vecs)."""
    n1 = nVeh * Hp + 1
    npair = nVeh * (nVeh - 1) // 2
    sum_nnz2 = npair * sum((2 * (k + 1) + 1) ** 2 for k in range(Hp))
    nnzG = npair * sum(2 * (k + 1) + 1 for k in range(Hp)) + 2 * n1
    return n1 ** 3 / 3.0 + 2.0 * sum_nnz2 + 8.0 * n1 ** 2 + 8.0 * nnzG


def assembly_bytes_per_qp(nVeh, Hp):
    """out = 8 (n1^2 + n1 + mc n1 + mc + 2 n1); in = 8 (16 nVeh + nVeh^2 + nVeh Hp)."""
    n1 = nVeh * Hp + 1
    mc = Hp * nVeh * (nVeh - 1) // 2
    return 8 * (n1 * n1 + n1 + mc * n1 + mc + 2 * n1) + 8 * (16 * nVeh + nVeh * nVeh + nVeh * Hp)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def fp64_peak():
    """FP64 peak of this pool's B200 from the committed measurement (profiles/fp64_peak.json, written from
    tools/microbench_dmma.cu / tools/microbench.cu runs): MEASURED_PEAKS.json carries no FP64 entry."""
    with open(os.path.join(ROOT, "profiles", "fp64_peak.json")) as f:
        d = json.load(f)
    return float(d["dmma_tflops"]), d["source"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def rank_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------------- reference arm
def cpu_controller_run(B, nVeh, Hp, steps, warmup, instance0, threads, opts, noise_sigma, seed, uMax, duLim, step_lo, step_hi):
    """The oracle port of the reference's CPU path on the same closed-loop workload: per MPC step
    MPCclass set-up (orc_mpc_setup) + SCP_controller (orc_scp_controller_batch, dense assembly + coneqp), advance
    on the linear model.  Returns (QPs solved in the timed steps, seconds, ipm iterations)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle  # noqa: E402  (cpu_baseline / --impl reference only)
    scen = importlib.import_module(PKG + ".scenarios")
    cb = scen.circle_batch(B, nVeh=nVeh, Hp=Hp, instance0=instance0, step_lo=step_lo, step_hi=step_hi)
    x0, u0, u = cb.x0.copy(), cb.u0.copy(), np.zeros((B, nVeh * Hp))
    qps = ipm = 0
    t_total = 0.0
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        S = oracle.mpc_setup(x0, u0, cb.veh, cb.poly, Hp=Hp, dt=scen.DT, noise_sigma=noise_sigma, seed=seed,
                             instance0=instance0, noise_counter=s)
        R = oracle.scp_controller_batch(S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], cb.dsafe, u, opts=opts, threads=threads)
        dt_ = time.perf_counter() - t0
        u = R["u"]
        ua = np.clip(R["U"][:, 0, :], -uMax, uMax)
        ua = np.clip(ua, u0 - duLim, u0 + duLim)
        abe = S["abe"]
        Ad, Bd, Ed = abe[..., :36].reshape(B, nVeh, 6, 6), abe[..., 36:42], abe[..., 42:48]
        x0 = np.einsum("bvij,bvj->bvi", Ad, x0) + Bd * ua[..., None] + Ed
        u0 = ua
        if s >= warmup:
            qps += int(R["scp_iters"].sum()); ipm += int(R["ipm_iters"].sum()); t_total += dt_
    return qps, t_total, ipm


def run_reference(args):
    rank, _, world = rank_env()
    if rank != 0:
        return
    scen = importlib.import_module(PKG + ".scenarios")
    cores = os.cpu_count() or 1
    # the product arm's own configuration: the same `batch` instances (global ids 0 .. batch-1, i.e. rank 0's shard), the same
    # closed-loop MPC steps; --cpu-sample B overrides (the product arm's cpu_baseline leg uses a smaller sample)
    Bs = args.cpu_sample if args.cpu_sample > 0 else args.batch
    opts = dict(abstol=1e-7, reltol=1e-6, feastol=1e-7, maxiters=100)        # CVXOPT's documented defaults
    qps, sec, ipm = cpu_controller_run(Bs, args.nveh, args.hp, args.steps, args.warmup, 0, cores, opts, args.noise_sigma,
                                       args.seed, scen.MECH_LIMIT, scen.DU_LIM, args.step_lo, args.step_hi)
    val = qps / sec if sec > 0 else 0.0
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, Bs),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{Bs} instances x {args.steps} closed-loop MPC steps ({qps} QPs, {ipm} IPM iterations), "
                                   f"oracle coneqp restatement (C) at CVXOPT default tolerances 1e-7/1e-6/1e-7 (looser than the "
                                   f"product arm's), {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, B_per_gpu, note=None):
    cfg = {"workload": f"default 8-vehicle circle scenario (Scenarios.py:109-125), Hp={args.hp}, batched x{B_per_gpu} per GPU, "
                       f"perturbed instances with process noise sigma={args.noise_sigma} (Monte-Carlo), closed-loop MPC steps "
                       f"starting at step U{{{args.step_lo}..{args.step_hi}}} (BASELINE.json configs[1])",
           "nVeh": args.nveh, "Hp": args.hp, "batch_per_gpu": B_per_gpu, "n1": args.nveh * args.hp + 1,
           "mc": args.hp * args.nveh * (args.nveh - 1) // 2, "l2": "flushed between timed steps (256 MiB write)"}
    if getattr(args, "trust_frac", 0) > 0:
        cfg["trust_radius"] = f"{args.trust_frac} * uLim"
    if getattr(args, "max_scp_iter", 0) > 0:
        cfg["max_scp_iter"] = args.max_scp_iter
    if note:
        cfg["note"] = note
    return cfg


# ---------------------------------------------------------------------------------------------------- product arm
def ncu_traffic(kernel, args=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed `ncu --set full` capture
    summary (profiles/ncu_traffic.json, written by tools/ncu_traffic.py); None if that kernel was not captured, or if
    this run is not the default workload the captures were taken on (8 vehicles, Hp = 10, 1024 instances per GPU)."""
    if args is not None and (args.nveh, args.hp, args.batch) != (8, 10, 1024):
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def strong_leg(args, mods, dev, rank, world, flush, barrier, total=4096):
    """The north-star configuration (BASELINE.json: a 4096-scenario batch over the GPUs of the box): `total` instances in
    all, rank r owns the contiguous shard parallel.shard_range(total, r, world); the same closed-loop MPC steps, timed per
    step with CUDA events (L2 flushed), summed per rank.  Returns (seconds on this rank, QPs, IPM iterations, shard size)."""
    import torch
    capi, batch, scen, par = mods
    lo, hi = par.shard_range(total, rank, world)
    B = hi - lo
    cb = scen.circle_batch(B, nVeh=args.nveh, Hp=args.hp, instance0=lo, step_lo=args.step_lo, step_hi=args.step_hi)
    p = capi.Params()
    capi.load().scpb200_default_params(C.byref(p))
    p.noise_sigma, p.seed, p.instance0 = args.noise_sigma, args.seed, lo
    if args.trust_frac > 0:
        p.trust_radius = args.trust_frac * p.uLim
    if args.max_scp_iter > 0:
        p.max_scp_iter = args.max_scp_iter
    bs = batch.BatchSCP(B, args.nveh, args.hp, params=p, device=dev, keep_log=False)
    bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((B, args.nveh * args.hp)))
    uMax, duLim = scen.MECH_LIMIT, scen.DU_LIM
    for s in range(args.warmup):
        bs.params.noise_counter = s
        bs.setup(); bs.solve(); bs.advance_linear(uMax, duLim)
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    barrier()
    for k in range(args.steps):
        flush.zero_()
        e0[k].record()
        bs.params.noise_counter = args.warmup + k
        bs.setup(); bs.solve(); bs.advance_linear(uMax, duLim)
        e1[k].record()
        cnt += torch.stack([bs.scp_iters.sum(), bs.ipm_iters.sum()])
    barrier()
    ms = [e0[k].elapsed_time(e1[k]) for k in range(args.steps)]
    asm_frac = None
    if rank == 0 and not args.skip_assembly:                # the assembly kernel at this shard size (north star: >= 60 % of HBM)
        out = bs.assemble_dense()
        a0 = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
        a1 = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
        for i in range(3):
            bs.assemble_dense_into(bs.u, out)
        for i in range(10):
            flush.zero_()
            a0[i].record(); bs.assemble_dense_into(bs.u, out); a1[i].record()
        torch.cuda.synchronize(dev)
        ams = float(np.mean([a0[i].elapsed_time(a1[i]) for i in range(10)]))
        peaks, _ = measured_peaks()
        asm_frac = assembly_bytes_per_qp(args.nveh, args.hp) * B / (ams * 1e-3) / 1e9 / peaks["hbm_gbs"]
        del out
    c = cnt.cpu()
    # the same shard and steps through the rollout entry (all timed MPC steps of every instance in ONE launch: no step of the
    # shard waits for its longest chain of QPs, which is what bounds the per-step line at 4096 / N instances per GPU)
    bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((B, args.nveh * args.hp)))
    bs.params.noise_counter = 0
    if args.warmup > 0:
        bs.rollout(args.warmup, uMax, duLim)
    torch.cuda.synchronize(dev)
    barrier()
    bs.params.noise_counter = args.warmup
    flush.zero_()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    ro = bs.rollout(args.steps, uMax, duLim)
    r1.record()
    barrier()
    ro_t = r0.elapsed_time(r1) * 1e-3
    ro_qps = int(ro["qp_total"].sum().item())
    return float(np.sum(ms)) * 1e-3, int(c[0]), int(c[1]), B, float(np.median(ms)), asm_frac, ro_t, ro_qps


def sharding_check(args, mods, dev, rank, world, nsample=64, nsteps=2):
    """Bit-identity of per-instance results across shards, on the hardware: rank r solves `nsample` instances of its OWN shard
    and the first `nsample` of rank (r+1) % world's shard (as a batch of its own, at other batch positions, on another
    GPU); SHA-256 of (U, u, traj, scp_iters, ipm_iters, status) after `nsteps` closed-loop MPC steps must agree.  At
    world = 1 the second batch is the same instances at shifted batch positions (instances 32 .. 32+nsample)."""
    import hashlib
    import torch
    import torch.distributed as dist
    capi, batch, scen, par = mods
    B = args.batch

    def digest(i0, n):
        cb = scen.circle_batch(n, nVeh=args.nveh, Hp=args.hp, instance0=i0, step_lo=args.step_lo, step_hi=args.step_hi)
        p = capi.Params()
        capi.load().scpb200_default_params(C.byref(p))
        p.noise_sigma, p.seed, p.instance0 = args.noise_sigma, args.seed, i0
        bs = batch.BatchSCP(n, args.nveh, args.hp, params=p, device=dev, keep_log=False)
        bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((n, args.nveh * args.hp)))
        for s in range(nsteps):
            bs.params.noise_counter = s
            bs.setup(); bs.solve(); bs.advance_linear(scen.MECH_LIMIT, scen.DU_LIM)
        torch.cuda.synchronize(dev)
        return [hashlib.sha256(b"".join(getattr(bs, k)[i].cpu().numpy().tobytes() for k in
                                        ("U", "u", "traj", "scp_iters", "ipm_iters", "status"))).hexdigest() for i in range(n)]

    if world == 1:
        a, b = digest(0, nsample), digest(nsample // 2, nsample)
        same = a[nsample // 2:] == b[: nsample - nsample // 2]
        return {"ok": bool(same), "instances_compared": nsample - nsample // 2,
                "how": "one GPU: the same global instances solved at two different batch positions, SHA-256 per instance"}
    own = digest(rank * B, nsample)
    nxt = digest(((rank + 1) % world) * B, nsample)
    gathered = [None] * world
    dist.all_gather_object(gathered, (own, nxt))
    ok = all(gathered[r][1] == gathered[(r + 1) % world][0] for r in range(world))
    return {"ok": bool(ok), "instances_compared": nsample * world,
            "how": f"rank r re-solved the first {nsample} instances of rank (r+1) % N's shard on its own GPU; SHA-256 per instance of "
                   f"U, u, traj, iteration counts and status after {nsteps} closed-loop MPC steps, all-gathered and compared"}



def run_product(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = rank_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    capi = importlib.import_module(PKG + "._capi")
    batch = importlib.import_module(PKG + ".batch")
    scen = importlib.import_module(PKG + ".scenarios")
    par = importlib.import_module(PKG + ".parallel")
    mods = (capi, batch, scen, par)
    B, nVeh, Hp = args.batch, args.nveh, args.hp
    inst0 = rank * B                                        # weak scaling: B instances per rank, disjoint global ids
    cb = scen.circle_batch(B, nVeh=nVeh, Hp=Hp, instance0=inst0, step_lo=args.step_lo, step_hi=args.step_hi)
    p = capi.Params()
    capi.load().scpb200_default_params(C.byref(p))
    p.noise_sigma, p.seed, p.instance0 = args.noise_sigma, args.seed, inst0
    if args.trust_frac > 0:
        p.trust_radius = args.trust_frac * p.uLim
    if args.max_scp_iter > 0:
        p.max_scp_iter = args.max_scp_iter
    bs = batch.BatchSCP(B, nVeh, Hp, params=p, device=dev, keep_log=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    uMax, duLim = scen.MECH_LIMIT, scen.DU_LIM

    def reset():
        bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((B, nVeh * Hp)))
        torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step(s):
        bs.params.noise_counter = s
        bs.setup()
        bs.solve()
        bs.advance_linear(uMax, duLim)

    # ---------------- device-resident timing (value) ----------------
    reset()
    for s in range(args.warmup):
        device_step(s)
    torch.cuda.synchronize(dev)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    evs0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    evs1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    qp_counts = torch.zeros(args.steps, dtype=torch.int64, device=dev)
    ipm_counts = torch.zeros(args.steps, dtype=torch.int64, device=dev)
    stat_counts = torch.zeros(6, dtype=torch.int64, device=dev)   # instances x steps with status bits 1,2,4,8,16,32
    launches0 = bs.kernel_launches
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                       # L2 flush, outside the event pair
        ev0[k].record()
        bs.params.noise_counter = args.warmup + k
        bs.setup()
        evs0[k].record()
        bs.solve()
        evs1[k].record()
        bs.advance_linear(uMax, duLim)
        ev1[k].record()
        qp_counts[k] = bs.scp_iters.sum()
        ipm_counts[k] = bs.ipm_iters.sum()
        stat_counts += torch.stack([((bs.status >> i) & 1).sum() for i in range(6)])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = bs.kernel_launches - launches0
    step_ms = np.array([ev0[k].elapsed_time(ev1[k]) for k in range(args.steps)])
    solve_ms = np.array([evs0[k].elapsed_time(evs1[k]) for k in range(args.steps)])
    setup_ms = np.array([ev0[k].elapsed_time(evs0[k]) for k in range(args.steps)])
    t_dev = float(step_ms.sum()) * 1e-3
    qps_rank = int(qp_counts.sum().item())
    ipm_rank = int(ipm_counts.sum().item())
    qp_per_step = qp_counts.cpu().numpy()
    ipm_per_step = ipm_counts.cpu().numpy()

    # ---------------- end-to-end through the public API with host buffers (e2e) ----------------
    # The caller's state lives in host memory (a pinned HostIO mirror of the batch's I/O arena: x0, u0, veh, poly, dsafe, warm
    # start u in; x0, u0, u, U, traj, obj, max_violation, iteration counts, status out).  Every step sends the inputs
    # (one contiguous host -> device copy), runs the controller stage, reads every result back (one device -> host copy)
    # and synchronises; the next step's x0 / u0 / u are the values just read back, in place in the host buffer.
    hio = bs.host_io()
    for k in ("x0", "u0", "veh", "poly", "dsafe"):
        getattr(hio, k)[...] = np.asarray(getattr(cb, k)).reshape(getattr(hio, k).shape)
    hio.u[...] = 0.0
    h2d, d2h = hio.nbytes_in, hio.nbytes_out

    def e2e_step(s):
        bs.upload(hio)
        bs.params.noise_counter = s
        bs.setup()
        bs.solve()
        bs.advance_linear(uMax, duLim)
        bs.download(hio)
        torch.cuda.synchronize(dev)
        return int(hio.scp_iters.sum())

    for s in range(args.warmup):
        e2e_step(s)
    barrier()
    t0 = time.perf_counter()
    e2e_qps = 0
    for k in range(args.steps):
        e2e_qps += e2e_step(args.warmup + k)
    barrier()
    t_e2e = time.perf_counter() - t0

    # ---------------- rollout entry: the same closed loop, all timed steps in ONE launch ----------------
    # (instances are independent across MPC steps: a CTA that finishes an instance's step re-queues it for the next one,
    # so no step of the batch waits for its longest chain of QPs; results are bit-identical to the per-step calls)
    reset()
    bs.params.noise_counter = 0
    if args.warmup > 0:
        bs.rollout(args.warmup, uMax, duLim)
    torch.cuda.synchronize(dev)
    barrier()
    bs.params.noise_counter = args.warmup
    flush.zero_()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    ro = bs.rollout(args.steps, uMax, duLim)
    r1.record()
    barrier()
    t_roll = r0.elapsed_time(r1) * 1e-3
    roll_qps, roll_ipm = int(ro["qp_total"].sum().item()), int(ro["ipm_total"].sum().item())

    # ---------------- assembly kernel (K2) against the HBM roofline ----------------
    asm = None
    if rank == 0 and not args.skip_assembly:
        outbuf = bs.assemble_dense()
        torch.cuda.synchronize(dev)
        a0 = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
        a1 = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
        for i in range(3):
            bs.assemble_dense_into(bs.u, outbuf)
        for i in range(10):
            flush.zero_()
            a0[i].record(); bs.assemble_dense_into(bs.u, outbuf); a1[i].record()
        torch.cuda.synchronize(dev)
        ams = float(np.mean([a0[i].elapsed_time(a1[i]) for i in range(10)]))
        peaks, how = measured_peaks()
        abytes = assembly_bytes_per_qp(nVeh, Hp) * B
        asm = {"bound": "hbm", "kernel": "k_assemble", "achieved": abytes / (ams * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
               "unit": "GB/s", "frac": abytes / (ams * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": ncu_traffic("k_assemble", args),
               "peak_source": how,
               "ms_per_launch": ams, "algorithmic_bytes_per_launch": abytes}
        if asm["traffic"]:                                  # DRAM-level: bytes that reached HBM inside the launch (ncu capture) / time
            asm["frac_dram"] = asm["traffic"] / (ams * 1e-3) / 1e9 / peaks["hbm_gbs"]
        del outbuf

    # ---------------- north-star strong configuration + sharding bit-identity ----------------
    st_t, st_qps, st_ipm, st_B, st_p50, st_asm, st_ro_t, st_ro_qps = strong_leg(args, mods, dev, rank, world, flush, barrier, total=args.strong_total)
    shard = sharding_check(args, mods, dev, rank, world)

    # ---------------- reduce over ranks ----------------
    t_max, qps_all, e2e_max, e2e_all = t_dev, qps_rank, t_e2e, e2e_qps
    st_t_max, st_qps_all, st_ipm_all = st_t, st_qps, st_ipm
    st_ro_t_max, st_ro_qps_all = st_ro_t, st_ro_qps
    roll_t_max, roll_qps_all, roll_ipm_all = t_roll, roll_qps, roll_ipm
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, t_roll, st_t, st_ro_t], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        cc = torch.tensor([qps_rank, e2e_qps, ipm_rank, roll_qps, roll_ipm, st_qps, st_ipm, st_ro_qps], dtype=torch.int64, device=dev)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        t_max, e2e_max, roll_t_max, st_t_max, st_ro_t_max = (float(v) for v in tt)
        qps_all, e2e_all, ipm_all, roll_qps_all, roll_ipm_all, st_qps_all, st_ipm_all, st_ro_qps_all = (int(v) for v in cc)
        # the optional all-gather of trajectories / statistics, off the timed path
        gathered = [torch.empty_like(bs.U) for _ in range(world)]
        dist.all_gather(gathered, bs.U)
    else:
        ipm_all = ipm_rank

    if rank == 0:
        value = qps_all / t_max
        fit = algorithmic_flops_per_ipm_iteration(nVeh, Hp)
        # dominant kernel: k_scp_solve (FP64 pipe).  Flops per launch = F_it x IPM iterations executed in that launch
        # (the start-point factorisations of cold-started QPs are not counted: conservative); duration = CUDA events
        # around the launch on its stream.
        fl = fit * ipm_per_step
        ach = float(fl.sum() / (solve_ms.sum() * 1e-3) / 1e12)
        fp64_pk, fp64_src = fp64_peak()
        # K1 (set-up): bytes in + out per instance over the CUDA-event time of its launch
        k1_bytes = sum(getattr(bs, k).numel() * 8 for k in ("x0", "u0", "veh", "poly", "ref", "g", "cterm", "H", "qv", "gamma0", "abe"))
        peaks_, how_ = measured_peaks()
        k1_ach = k1_bytes / (float(setup_ms.mean()) * 1e-3) / 1e9
        setup_roof = {"bound": "hbm", "kernel": "k_mpc_setup", "achieved": k1_ach, "peak": peaks_["hbm_gbs"], "unit": "GB/s",
                      "frac": k1_ach / peaks_["hbm_gbs"], "ms_per_launch": float(setup_ms.mean()), "algorithmic_bytes_per_launch": k1_bytes,
                      "share_of_step": float(setup_ms.sum() / step_ms.sum()), "peak_source": how_,
                      "note": "instruction-bound, not memory-bound (ncu, profiles/r02_end_k_mpc_setup_details.txt: IPC 2.07, issue slots 52 % busy, DRAM "
                              "throughput 0.1 %): one warp per vehicle, 8x8 Pade products on DMMA, reference sampler and Toeplitz recurrences by "
                              "lanes; bytes = inputs + every K1 output"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, B),
            "e2e": {"value": e2e_all / e2e_max, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "bound_detail": "FP64 pipe: DMMA m8n8k4 (the FP64 tensor path) and DFMA share one "
                         "peak on B200; the kernel is latency / issue bound far below it (DESIGN.md section 4)",
                         "kernel": "k_scp_solve", "achieved": ach, "peak": fp64_pk, "unit": "TFLOP/s",
                         "frac": ach / fp64_pk, "traffic": ncu_traffic("k_scp_solve", args),
                         "peak_source": "profiles/fp64_peak.json: " + fp64_src,
                         "algorithmic_flops_per_ipm_iteration": fit, "solve_share_of_step": float(solve_ms.sum() / step_ms.sum())},
            "roofline_assembly": asm,
            "roofline_setup": setup_roof,
            "north_star_strong": {"value": st_qps_all / st_t_max, "unit": UNIT, "batch_total": args.strong_total,
                                  "batch_per_gpu": st_B, "ms_per_step": 1e3 * st_t_max / args.steps, "p50_ms_per_mpc_step_rank0": st_p50,
                                  "scaling": "strong", "qps_total": st_qps_all, "ipm_iterations_total": st_ipm_all,
                                  "roofline_frac": float(fit * st_ipm_all / world / st_t_max / 1e12 / fp64_pk),
                                  "assembly_frac_hbm_rank0": st_asm,
                                  "rollout_value": st_ro_qps_all / st_ro_t_max, "rollout_ms_per_step": 1e3 * st_ro_t_max / args.steps,
                                  "rollout_qps_total": st_ro_qps_all,
                                  "note": "BASELINE.json north star: 4096 scenarios in total, contiguous shards of 4096/N per GPU, the same "
                                          "closed-loop MPC steps (max over ranks of the summed CUDA-event step times); rollout_value: the same shards and "
                                          "steps through scpb200_mpc_rollout (one launch per rank, bit-identical results)"},
            "sharding_bitwise_ok": shard,
            "rollout": {"value": roll_qps_all / roll_t_max, "unit": UNIT, "ms_per_step": 1e3 * roll_t_max / args.steps,
                        "steps_per_launch": args.steps, "qps_total": roll_qps_all, "ipm_iterations_total": roll_ipm_all,
                        "roofline_frac": float(fit * roll_ipm_all / world / roll_t_max / 1e12 / fp64_pk),
                        "note": "scpb200_mpc_rollout: the same closed-loop workload with all timed MPC steps of every instance in one "
                                "launch (instances re-queued across steps inside the kernel); bit-identical results, no per-step tail"},
            "stats": {"qps_total": qps_all, "ipm_iterations_total": ipm_all, "qp_per_instance_step": qps_all / (world * B * args.steps),
                      "ipm_per_qp": ipm_all / max(1, qps_all),
                      "status_counts_rank0": dict(zip(["qp_maxiter", "qp_pivot", "scp_maxiter", "infeasible", "setup", "qp_dres_floor"],
                                                      [int(v) for v in stat_counts.cpu()])),
                      "wall_s_bracket": t_wall, "p50_ms_per_mpc_step": float(np.median(step_ms)),
                      "plan": bs.plan()},
        }
        if not args.skip_cpu and world == 1:                   # the CPU baseline is reported at N = 1 only
            # in a child process: the product arm's own process never imports, links or maps anything under oracle/
            sample = args.cpu_sample if args.cpu_sample > 0 else 256
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(args.steps), "--warmup", "0",
                   "--cpu-sample", str(sample), "--batch", str(B), "--hp", str(Hp), "--nveh", str(nVeh),
                   "--noise-sigma", str(args.noise_sigma), "--seed", str(args.seed), "--step-lo", str(args.step_lo),
                   "--step-hi", str(args.step_hi)]
            try:
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=900).stdout.strip().splitlines()
                line["cpu_baseline"] = json.loads(out[-1])["cpu_baseline"]
            except Exception as e:                              # report, do not fail the product line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                        "sample": f"child process failed: {e!r}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="instances per GPU")
    ap.add_argument("--nveh", type=int, default=8)
    ap.add_argument("--hp", type=int, default=10)
    ap.add_argument("--noise-sigma", dest="noise_sigma", type=float, default=3e-6)
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--step-lo", dest="step_lo", type=int, default=4)
    ap.add_argument("--step-hi", dest="step_hi", type=int, default=7)
    ap.add_argument("--cpu-sample", dest="cpu_sample", type=int, default=0,
                    help="instances of the CPU run (0: --impl reference runs the product arm's full batch; the cpu_baseline leg 256)")
    ap.add_argument("--trust-radius-frac", dest="trust_frac", type=float, default=0.0,
                    help="BASELINE configs[3]: trust region |u - ubar|_inf <= frac * uLim folded into the box (0 = off, the reference)")
    ap.add_argument("--max-scp-iter", dest="max_scp_iter", type=int, default=0, help="SCP iteration cap (0 = the reference's 20)")
    ap.add_argument("--strong-total", dest="strong_total", type=int, default=4096,
                    help="total instances of the north-star strong-scaling leg (sharded over the ranks)")
    ap.add_argument("--skip-cpu", dest="skip_cpu", action="store_true")
    ap.add_argument("--skip-assembly", dest="skip_assembly", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
