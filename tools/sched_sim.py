"""Event-driven model of K4's work queue on the measured per-QP interior-point iteration counts
(gpurun_out/straggler_dump.npz from tools/straggler_stats.py): compares scheduling policies against the two lower
bounds of a step (longest chain, total work / CTAs)."""
import heapq, sys
import numpy as np
D = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/straggler_dump.npz")
G = int(__import__('os').environ.get('SIM_CTAS', 296))
MPIN = int(__import__('os').environ.get('SIM_MPIN', 99))      # rr: an instance that has done this many QPs is no longer parked
A_US = float(sys.argv[2]) if len(sys.argv) > 2 else 47.0     # per interior-point iteration
B_US = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0     # per QP (linearise, evaluate, park / resume)
steps = sorted(int(k.split("_")[1]) for k in D.files if k.startswith("scp_"))

def qp_costs(s):
    per, scp = D[f"perqp_{s}"].astype(float), D[f"scp_{s}"]
    return [A_US * per[b, : scp[b]] + B_US for b in range(len(scp))]

def simulate(costs, order, policy, npinned=G // 2, pred=None, prio=None):
    """policy: 'rr' (FIFO ring, first npinned of `order` run to completion), 'lpt' (run to completion in `order`),
    'lrpt' (preemptive at QP granularity, highest predicted remaining work first)."""
    nB = len(costs)
    done = [0] * nB
    t_free = [(0.0, c) for c in range(G)]
    heapq.heapify(t_free)
    if policy in ("rr", "lpt"):
        from collections import deque
        ring = deque(order)
        pinned = set(order[:npinned]) if policy == "rr" else set(order)
        pending = []          # (time available, seq, b): parked instances become visible when their QP ends
        seq = 0
        while ring or pending:
            t, c = heapq.heappop(t_free)
            while pending and pending[0][0] <= t:
                ring.append(heapq.heappop(pending)[2])
            if not ring:
                t = pending[0][0]
                heapq.heappush(t_free, (t, c))
                continue
            b = ring.popleft()
            if b in pinned or done[b] >= MPIN:
                t += costs[b][done[b]:].sum(); done[b] = len(costs[b])
            else:
                t += costs[b][done[b]]; done[b] += 1
                if done[b] < len(costs[b]):
                    heapq.heappush(pending, (t, seq, b)); seq += 1
            heapq.heappush(t_free, (t, c))
        return max(t for t, _ in t_free)
    # lrpt with predicted remaining QPs: pred[b] QPs expected; beyond the prediction assume the cap of 20
    rem = prio(done) if prio else (lambda b: (pred[b] - done[b]) if done[b] < pred[b] else (20 - done[b]) + 0.5)
    avail = [(-rem(b), b) for b in range(nB)]
    heapq.heapify(avail)
    pending = []
    while avail or pending:
        t, c = heapq.heappop(t_free)
        while pending and pending[0][0] <= t:
            _, b = heapq.heappop(pending)
            heapq.heappush(avail, (-rem(b), b))
        if not avail:
            t = pending[0][0]
            heapq.heappush(t_free, (t, c))
            continue
        _, b = heapq.heappop(avail)
        t += costs[b][done[b]]; done[b] += 1
        if done[b] < len(costs[b]):
            heapq.heappush(pending, (t, b))
        heapq.heappush(t_free, (t, c))
    return max(t for t, _ in t_free)

print(f"model: {A_US} us / ipm iteration, {B_US} us / QP, {G} CTAs")
print("step  measured   rr(cur)   lpt   lrpt(pred=prev scp)  lrpt(oracle)   bound: chain / work")
prev_ipm = prev_scp = None
for s in steps:
    costs = qp_costs(s)
    nB = len(costs)
    ipm = np.array([D[f"perqp_{s}"][b, : D[f'scp_{s}'][b]].sum() for b in range(nB)])
    order = list(np.argsort(-prev_ipm, kind="stable")) if prev_ipm is not None else list(range(nB))
    pred = prev_scp if prev_scp is not None else np.full(nB, 10)
    chain = max(c.sum() for c in costs); work = sum(c.sum() for c in costs) / G
    rank = np.empty(nB); rank[np.array(order)] = np.arange(nB) / nB          # 0 = predicted longest
    lef = lambda done: (lambda b: done[b] - rank[b])
    lef2 = lambda done: (lambda b: (done[b] if done[b] >= 6 else 0) - rank[b])
    dl = np.abs(D[f"log_{s}"][:, :, 4].astype(float)) + 1e-30
    def pred_delta(done):
        def f(b):
            it = done[b]
            if it < 2: return 8.0 - it - rank[b]
            d1, d0 = dl[b, it - 1], dl[b, it - 2]
            ratio = d0 / d1
            if ratio <= 1.5: est = 20 - it
            else: est = min(20 - it, max(1.0, np.ceil(np.log(d1 / 1e-3) / np.log(ratio))))
            return est - rank[b]
        return f
    r = [simulate(costs, order, "lrpt", prio=pred_delta), simulate(costs, order, "lrpt", prio=lef2), simulate(costs, order, "rr"), simulate(costs, order, "lpt"), simulate(costs, order, "lrpt", pred=pred),
         simulate(costs, order, "lrpt", pred=D[f"scp_{s}"])]
    print(f"{s:3d}  {float(D[f'ms_{s}']):8.2f}  " + "  ".join(f"{x / 1e3:7.2f}" for x in r) + f"     {chain / 1e3:6.2f} / {work / 1e3:6.2f}")
    prev_ipm, prev_scp = ipm, D[f"scp_{s}"]
