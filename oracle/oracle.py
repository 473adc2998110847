"""ctypes front-end of oracle/scp_oracle.c — the CPU restatement of the SCP-QP hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs.  Nothing under the product package imports this module.

Parity status: set-up/assembly/evaluation are pinned by tests/golden (generated from the reference's own
Python by oracle/make_golden.py); the QP solve is PARITY UNPINNED against CVXOPT itself (not installable
here; see the header of scp_oracle.c) and is certified per solve by KKT residuals instead.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "scp_oracle.c")
_INC = os.path.join(_HERE, "coneqp_core.inc")
_LIB = os.path.join(_HERE, "libscp_oracle.so")
_lib = None

DBL = C.POINTER(C.c_double)


class QPOpts(C.Structure):
    _fields_ = [("abstol", C.c_double), ("reltol", C.c_double), ("feastol", C.c_double), ("maxiters", C.c_int)]


#: CVXOPT's documented defaults (solvers.options): abstol 1e-7, reltol 1e-6, feastol 1e-7, maxiters 100.
CVXOPT_DEFAULT = dict(abstol=1e-7, reltol=1e-6, feastol=1e-7, maxiters=100)
#: The parity setting: the minimiser is only pinned to 1e-5 in u when the gap is driven down.
TIGHT = dict(abstol=1e-10, reltol=1e-10, feastol=1e-10, maxiters=100)


def build(force: bool = False) -> str:
    """Compile scp_oracle.c with gcc into oracle/libscp_oracle.so (git-ignored, travels with gpurun)."""
    newest = max(os.path.getmtime(_SRC), os.path.getmtime(_INC))
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < newest:
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-fno-fast-math", "-o", _LIB, _SRC, "-lquadmath", "-lm"]
        subprocess.run(cmd, check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_log_width.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(DBL)


def _c(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _opts(opts):
    d = dict(TIGHT)
    if opts:
        d.update(opts)
    return QPOpts(d["abstol"], d["reltol"], d["feastol"], int(d["maxiters"]))


# ------------------------------------------------------------------------------------------------ small pieces
def expm(A):
    A = _c(A)
    n = A.shape[0]
    E = np.empty_like(A)
    rc = lib().orc_expm(C.c_int(n), _p(A), _p(E))
    assert rc == 0
    return E


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    u32 = C.POINTER(C.c_uint32)
    lib().orc_philox4x32_10(ctr.ctypes.data_as(u32), key.ctypes.data_as(u32), out.ctypes.data_as(u32))
    return out


def noise_pair(seed, instance, vehicle, counter):
    out = np.empty(2)
    lib().orc_noise_pair(C.c_uint64(seed), C.c_uint32(instance), C.c_uint32(vehicle), C.c_uint32(counter), _p(out))
    return out


def jacobian(x, u, Lf, Lr, noise=None):
    x = _c(x)
    Ac, Bc, Ec = np.empty((6, 6)), np.empty(6), np.empty(6)
    nz = None if noise is None else _c(noise)
    lib().orc_jacobian(_p(x), C.c_double(u), C.c_double(Lf), C.c_double(Lr), _p(nz), _p(Ac), _p(Bc), _p(Ec))
    return Ac, Bc, Ec


def discretize(x0, u0, Lf, Lr, dt, noise=None):
    x0 = _c(x0)
    Ad, Bd, Ed = np.empty((6, 6)), np.empty(6), np.empty(6)
    nz = None if noise is None else _c(noise)
    rc = lib().orc_discretize(_p(x0), C.c_double(u0), C.c_double(Lf), C.c_double(Lr), C.c_double(dt), _p(nz),
                              _p(Ad), _p(Bd), _p(Ed))
    assert rc == 0
    return Ad, Bd, Ed


def sample_reference(nSamples, poly, x, y, step):
    poly = _c(poly)
    out = np.empty((nSamples, 2))
    rc = lib().orc_sample_reference(C.c_int(nSamples), C.c_int(poly.shape[0]), _p(poly), C.c_double(x),
                                    C.c_double(y), C.c_double(step), _p(out))
    if rc:
        raise IndexError("reference would index past the polyline (index_min == nPts)")
    return out


# ------------------------------------------------------------------------------------------------ K1
def mpc_setup(x0, u0, veh, poly, Hp, dt, noise_sigma=0.0, seed=0, instance0=0, noise_counter=0):
    """a1-a5 + a14 for a batch.  Shapes: x0[B,nVeh,6] u0[B,nVeh] veh[B,nVeh,5] poly[B,nVeh,nPts,2]."""
    x0 = _c(x0)
    B, nVeh = x0.shape[0], x0.shape[1]
    u0 = _c(u0, (B, nVeh))
    veh = _c(veh, (B, nVeh, 5))
    poly = _c(poly)
    nPts = poly.shape[2]
    out = dict(
        ref=np.empty((B, nVeh, Hp, 2)), g=np.empty((B, nVeh, Hp, 2)), cterm=np.empty((B, nVeh, Hp, 2)),
        H=np.empty((B, nVeh, Hp, Hp)), qv=np.empty((B, nVeh, Hp)), gamma0=np.empty(B), abe=np.empty((B, nVeh, 48)),
    )
    rc = lib().orc_mpc_setup(
        C.c_int(B), C.c_int(nVeh), C.c_int(Hp), C.c_int(nPts), C.c_double(dt), C.c_double(noise_sigma),
        C.c_uint64(seed), C.c_uint32(instance0), C.c_uint32(noise_counter), _p(x0), _p(u0), _p(veh), _p(poly),
        _p(out["ref"]), _p(out["g"]), _p(out["cterm"]), _p(out["H"]), _p(out["qv"]), _p(out["gamma0"]), _p(out["abe"]))
    out["rc"] = rc
    return out


def forward(g, cterm, u):
    """pos[nVeh,Hp,2] for one instance (SCP_controller.py:199-213)."""
    g = _c(g)
    nVeh, Hp = g.shape[0], g.shape[1]
    pos = np.empty((nVeh, Hp, 2))
    lib().orc_forward(C.c_int(nVeh), C.c_int(Hp), _p(g), _p(_c(cterm)), _p(_c(u).ravel()), _p(pos))
    return pos


def _obst_args(nVeh, Hp, dsafe_obst, obst):
    if obst is None:
        return 0, None, None
    obst = _c(obst)
    nObst = obst.shape[-3]
    return nObst, _c(dsafe_obst), obst


# ------------------------------------------------------------------------------------------------ K2
def assemble_dense(g, cterm, H, qv, ubar, dsafe, dsafeExtra, uLim, omega_weight=1e5, omega_ub=1e25,
                   trust_radius=np.inf, dsafe_obst=None, obst=None):
    """One instance: dense (P,q,Aineq,bineq,lb,ub) in the layout of SCP_controller.py:93-128."""
    g = _c(g)
    nVeh, Hp = g.shape[0], g.shape[1]
    nObst, dso, ob = _obst_args(nVeh, Hp, dsafe_obst, obst)
    n1 = nVeh * Hp + 1
    mc = Hp * (nVeh * (nVeh - 1) // 2 + nVeh * nObst)
    P, q = np.empty((n1, n1)), np.empty(n1)
    A, b = np.empty((mc, n1)), np.empty(mc)
    lb, ub = np.empty(n1), np.empty(n1)
    tr = 1e308 if not np.isfinite(trust_radius) else float(trust_radius)
    lib().orc_assemble_dense(
        C.c_int(nVeh), C.c_int(Hp), C.c_int(nObst), _p(g), _p(_c(cterm)), _p(_c(H)), _p(_c(qv)),
        _p(_c(ubar).ravel()), _p(_c(dsafe)), _p(dso), _p(ob), C.c_double(dsafeExtra), C.c_double(uLim),
        C.c_double(omega_weight), C.c_double(omega_ub), C.c_double(tr), _p(P), _p(q), _p(A), _p(b), _p(lb), _p(ub))
    return P, q, A, b, lb, ub


def qcqp_evaluate(g, cterm, H, qv, gamma0, u, dsafe, dsafeExtra, tol=2 * 2.1 * 1e-3, dsafe_obst=None, obst=None,
                  obstacle_mode=0):
    g = _c(g)
    nVeh, Hp = g.shape[0], g.shape[1]
    nObst, dso, ob = _obst_args(nVeh, Hp, dsafe_obst, obst)
    feas = C.c_int(0)
    obj, mv, sv = C.c_double(0), C.c_double(0), C.c_double(0)
    ci = np.empty((nVeh, nVeh, Hp))
    cio = np.empty((nVeh, max(nObst, 1), Hp))
    lib().orc_qcqp_evaluate(
        C.c_int(nVeh), C.c_int(Hp), C.c_int(nObst), _p(g), _p(_c(cterm)), _p(_c(H)), _p(_c(qv)), C.c_double(gamma0),
        _p(_c(u).ravel()), _p(_c(dsafe)), _p(dso), _p(ob), C.c_double(dsafeExtra), C.c_double(tol),
        C.c_int(obstacle_mode), C.byref(feas), C.byref(obj), C.byref(mv), C.byref(sv), _p(ci), _p(cio))
    return dict(feasible=bool(feas.value), obj=obj.value, max_violation=mv.value, sum_violations=sv.value,
                ci=ci, ci_obst=cio[:, :nObst])


# ------------------------------------------------------------------------------------------------ a8
def coneqp(P, q, G, h, opts=None, quad=False):
    P, q, G, h = _c(P), _c(q).ravel(), _c(G), _c(h).ravel()
    n, m = P.shape[0], G.shape[0]
    x, s, z, info = np.empty(n), np.empty(m), np.empty(m), np.empty(8)
    o = _opts(opts)
    fn = lib().orc_coneqp_q if quad else lib().orc_coneqp
    st = fn(C.c_int(n), C.c_int(m), _p(P), _p(q), _p(G), _p(h), C.byref(o), _p(x), _p(s), _p(z), _p(info))
    return dict(x=x, s=s, z=z, status=st, pcost=info[0], dcost=info[1], gap=info[2], relgap=info[3], pres=info[4],
                dres=info[5], iterations=int(info[6]))


def kkt_residuals(P, q, G, h, x, z):
    P, q, G, h = _c(P), _c(q).ravel(), _c(G), _c(h).ravel()
    res = np.empty(4)
    lib().orc_kkt_residuals(C.c_int(P.shape[0]), C.c_int(G.shape[0]), _p(P), _p(q), _p(G), _p(h), _p(_c(x).ravel()),
                            _p(_c(z).ravel()), _p(res))
    return dict(stationarity=res[0], primal=res[1], dual=res[2], complementarity=res[3])


def qp_boxed(P, q, A, b, lb, ub, opts=None, inf_bound=1e20, quad=False):
    """The QP as SCP_controller.py:135-141 poses it; returns x, fval (=prob.value), multipliers, info."""
    P, q, A, b = _c(P), _c(q).ravel(), _c(A), _c(b).ravel()
    lb, ub = _c(lb).ravel(), _c(ub).ravel()
    n1, mc = P.shape[0], A.shape[0]
    x, info = np.empty(n1), np.empty(8)
    zA, zub, zlb = np.empty(mc), np.empty(n1), np.empty(n1)
    fval = C.c_double(0)
    o = _opts(opts)
    st = lib().orc_qp_boxed(C.c_int(n1), C.c_int(mc), _p(P), _p(q), _p(A), _p(b), _p(lb), _p(ub), C.byref(o),
                            C.c_double(inf_bound), C.c_int(int(quad)), _p(x), C.byref(fval), _p(info), _p(zA), _p(zub),
                            _p(zlb))
    return dict(x=x, fval=fval.value, status=st, zA=zA, zub=zub, zlb=zlb, gap=info[2], relgap=info[3],
                pres=info[4], dres=info[5], iterations=int(info[6]))


def rate_rows(nVeh, Hp, u_prev, duLim):
    """Dense steering-rate rows of the product's `enable_rate_rows` extension (include/scpb200.h): A[2n, n1], b[2n] in the
    order (v, k, +), (v, k, -)."""
    n = nVeh * Hp
    A, b = np.empty((2 * n, n + 1)), np.empty(2 * n)
    lib().orc_rate_rows(C.c_int(nVeh), C.c_int(Hp), _p(_c(u_prev).ravel()), C.c_double(duLim), _p(A), _p(b))
    return A, b


def stack_G(A, b, lb, ub, inf_bound=1e20):
    """G=[Aineq; I; -I], h=[bineq; ub; -lb] with infinite bounds dropped, as orc_qp_boxed builds it."""
    n1 = A.shape[1]
    ku = [i for i in range(n1) if not (inf_bound > 0 and abs(ub[i]) >= inf_bound)]
    kl = [i for i in range(n1) if not (inf_bound > 0 and abs(lb[i]) >= inf_bound)]
    I = np.eye(n1)
    return np.vstack([A, I[ku], -I[kl]]), np.concatenate([np.ravel(b), np.ravel(ub)[ku], -np.ravel(lb)[kl]])


# ------------------------------------------------------------------------------------------------ a10-a12
def scp_optimizer(g, cterm, H, qv, gamma0, dsafe, u, dsafeExtra=1.0, uLim=3 * np.pi / 180, delta_tol=1e-3,
                  omega_weight=1e5, omega_ub=1e25, constraint_tol=2 * 2.1 * 1e-3, max_scp_iter=20,
                  trust_radius=np.inf, opts=None, inf_bound=1e20, dsafe_obst=None, obst=None, obstacle_mode=0,
                  quad=False, u_prev=None, duLim=0.0):
    """One instance of SCP_optimizer (SCP_controller.py:74-197).  Returns dict(u, feasible, obj, iters, log, u_hist).
    u_prev[nVeh] (with duLim): the product's steering-rate-row extension (rate_rows below) in every QP."""
    g = _c(g)
    nVeh, Hp = g.shape[0], g.shape[1]
    nObst, dso, ob = _obst_args(nVeh, Hp, dsafe_obst, obst)
    n = nVeh * Hp
    u = _c(u).ravel().copy()
    W = lib().orc_log_width()
    log = np.zeros((max_scp_iter, W))
    uh = np.zeros((max_scp_iter, n))
    feas = C.c_int(0)
    obj = C.c_double(0)
    o = _opts(opts)
    tr = 1e308 if not np.isfinite(trust_radius) else float(trust_radius)
    up = None if u_prev is None else _c(u_prev).ravel()
    its = lib().orc_scp_optimizer_rate(
        C.c_int(nVeh), C.c_int(Hp), C.c_int(nObst), _p(g), _p(_c(cterm)), _p(_c(H)), _p(_c(qv)), C.c_double(gamma0),
        _p(_c(dsafe)), _p(dso), _p(ob), C.c_double(dsafeExtra), C.c_double(uLim), C.c_double(delta_tol),
        C.c_double(omega_weight), C.c_double(omega_ub), C.c_double(constraint_tol), C.c_int(max_scp_iter),
        C.c_double(tr), C.c_int(obstacle_mode), C.byref(o), C.c_double(inf_bound), C.c_int(int(quad)), _p(u),
        C.byref(feas), C.byref(obj), _p(log), _p(uh), None if up is None else _p(up), C.c_double(duLim))
    return dict(u=u, feasible=bool(feas.value), obj=obj.value, iters=its, log=log[:its], u_hist=uh[:its])


def scp_controller_batch(g, cterm, H, qv, gamma0, dsafe, u, dsafeExtra=1.0, uLim=3 * np.pi / 180, delta_tol=1e-3,
                         omega_weight=1e5, omega_ub=1e25, constraint_tol=2 * 2.1 * 1e-3, max_scp_iter=20,
                         trust_radius=np.inf, opts=None, inf_bound=1e20, threads=1):
    """Controller stage of one MPC step for a batch (nObst=0), optionally over `threads` host threads."""
    g = _c(g)
    B, nVeh, Hp = g.shape[0], g.shape[1], g.shape[2]
    n = nVeh * Hp
    cterm, H, qv, gamma0, dsafe = _c(cterm), _c(H), _c(qv), _c(gamma0), _c(dsafe)
    u = _c(u, (B, n)).copy()
    traj, U, stats = np.empty((B, Hp, 2, nVeh)), np.empty((B, Hp, nVeh)), np.empty((B, 4))
    o = _opts(opts)
    tr = 1e308 if not np.isfinite(trust_radius) else float(trust_radius)
    L = lib()

    def run(lo, hi):
        if hi <= lo:
            return
        L.orc_scp_controller_batch(
            C.c_int(hi - lo), C.c_int(nVeh), C.c_int(Hp), C.c_int(0), _p(g[lo:hi]), _p(cterm[lo:hi]), _p(H[lo:hi]),
            _p(qv[lo:hi]), _p(gamma0[lo:hi]), _p(dsafe[lo:hi]), None, None, C.c_double(dsafeExtra), C.c_double(uLim),
            C.c_double(delta_tol), C.c_double(omega_weight), C.c_double(omega_ub), C.c_double(constraint_tol),
            C.c_int(max_scp_iter), C.c_double(tr), C.c_int(0), C.byref(o), C.c_double(inf_bound), _p(u[lo:hi]),
            _p(traj[lo:hi]), _p(U[lo:hi]), _p(stats[lo:hi]))

    threads = max(1, min(int(threads), B))
    if threads == 1:
        run(0, B)
    else:
        # interleave chunks so that expensive instances spread across threads
        chunk = max(1, B // (threads * 8))
        spans = [(lo, min(B, lo + chunk)) for lo in range(0, B, chunk)]
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda s: run(*s), spans))
    return dict(u=u, traj=traj, U=U, scp_iters=stats[:, 0].astype(np.int64), ipm_iters=stats[:, 1].astype(np.int64),
                feasible=stats[:, 2].astype(bool), obj=stats[:, 3])


def ode_predict(x, u_ref, Lf, Lr, T, steps=10, tol=1e-12):
    """Delay-compensation prediction for one vehicle (MPC_Iter.py:25-33); returns [steps, 6]."""
    out = np.empty((steps, 6))
    lib().orc_ode_predict(_p(_c(x)), C.c_double(u_ref), C.c_double(Lf), C.c_double(Lr), C.c_double(T), C.c_int(steps),
                          C.c_double(tol), _p(out))
    return out
