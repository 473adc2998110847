"""Experiment (tuning aid): accuracy of u against the __float128 oracle and interior-point iterations per QP on the
configs[1] workload for a list of stopping tolerances.   python tools/exp_tolerance.py [B]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle
import test_gpu_workloads as W
import torch
PKG = "senquential-convex-programming-for-trajectory-planning_b200"
mods = dict(torch=torch, capi=importlib.import_module(PKG + "._capi"), batch=importlib.import_module(PKG + ".batch"),
            scen=importlib.import_module(PKG + ".scenarios"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Hp = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cb = mods["scen"].circle_batch(B, nVeh=8, Hp=Hp, step_lo=4, step_hi=7)
for kw in [dict(), dict(qp_reltol=1e-11), dict(qp_reltol=1e-12), dict(qp_reltol=1e-13, qp_abstol=1e-7), dict(qp_reltol=1e-13, qp_abstol=3e-8),
           dict(qp_reltol=1e-12, qp_dres_floor_factor=10)]:
    t = time.time()
    bs, recs, final = W._capture_chain(mods, cb, Hp, 20, **kw)
    res = W._oracle_check_qps(oracle, bs, cb, recs)
    nqp = res["n"]
    print(f"== {kw}: QPs {nqp}, ipm/QP {final['ipm_iters'].sum() / nqp:.2f}, max|u-u*| {res['du'].max():.2e}, p99.9 {np.quantile(res['du'], 0.999):.2e}, "
          f"rel obj {res['df'].max():.1e}, viol {res['viol'].max():.1e}, floor QPs {int(((res['gpu_qp_status'] & 32) != 0).sum())}, "
          f"status bits {np.bincount(final['status'] & 3, minlength=4).tolist()}, {time.time() - t:.0f}s", flush=True)
