"""Python front-end of the kernel-logic emulator (tests/emu/emu.cpp).  TEST INFRASTRUCTURE ONLY — see emu.cpp."""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = "senquential-convex-programming-for-trajectory-planning_b200"
SRC = os.path.join(HERE, "emu.cpp")
LIB = os.path.join(HERE, "libscpb200_emu.so")
CSRC = os.path.join(ROOT, PKG, "csrc")

capi = importlib.import_module(PKG + "._capi")
Dims, Params = capi.Dims, capi.Params
_lib = None
DBL = C.POINTER(C.c_double)
I32 = C.POINTER(C.c_int32)


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    newest = max(os.path.getmtime(d) for d in deps)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", LIB, SRC],
                       check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.emu_scp_shared_bytes.restype = C.c_size_t
    return _lib


def config(nt=128, reverse=False, force_global_S=False, alpha_slots=-1):
    lib().emu_config(C.c_int(nt), C.c_int(int(reverse)), C.c_int(int(force_global_S)), C.c_int(alpha_slots))


def _d(a):
    return None if a is None else a.ctypes.data_as(DBL)


def _i(a):
    return None if a is None else a.ctypes.data_as(I32)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def mpc_setup(x0, u0, veh, poly, Hp, params):
    x0, u0, veh, poly = _c(x0), _c(u0), _c(veh), _c(poly)
    B, nVeh, nPts = x0.shape[0], x0.shape[1], poly.shape[2]
    d = Dims(B, nVeh, Hp, 0, nPts)
    out = dict(ref=np.zeros((B, nVeh, Hp, 2)), g=np.zeros((B, nVeh, Hp, 2)), cterm=np.zeros((B, nVeh, Hp, 2)),
               H=np.zeros((B, nVeh, Hp, Hp)), qv=np.zeros((B, nVeh, Hp)), gamma0=np.zeros(B),
               abe=np.zeros((B, nVeh, 48)), setup_status=np.zeros(B, dtype=np.int32))
    lib().emu_mpc_setup(C.byref(d), C.byref(params), _d(x0), _d(u0), _d(veh), _d(poly), _d(out["ref"]), _d(out["g"]),
                        _d(out["cterm"]), _d(out["H"]), _d(out["qv"]), _d(out["gamma0"]), _d(out["abe"]),
                        _i(out["setup_status"]))
    return out


def _obst(dsafe_obst, obst):
    if obst is None:
        return 0, None, None
    obst = _c(obst)
    return obst.shape[1], _c(dsafe_obst), obst


def assemble_dense(g, cterm, H, qv, ubar, dsafe, params, dsafe_obst=None, obst=None):
    g, cterm, H, qv, ubar, dsafe = _c(g), _c(cterm), _c(H), _c(qv), _c(ubar), _c(dsafe)
    B, nVeh, Hp = g.shape[0], g.shape[1], g.shape[2]
    nObst, dso, ob = _obst(dsafe_obst, obst)
    d = Dims(B, nVeh, Hp, nObst, 2)
    n1 = nVeh * Hp + 1
    mc = Hp * (nVeh * (nVeh - 1) // 2 + nVeh * nObst)
    P, q, A, b = np.full((B, n1, n1), np.nan), np.full((B, n1), np.nan), np.full((B, mc, n1), np.nan), np.full((B, mc), np.nan)
    lb, ub = np.full((B, n1), np.nan), np.full((B, n1), np.nan)
    lib().emu_assemble_dense(C.byref(d), C.byref(params), _d(g), _d(cterm), _d(H), _d(qv), _d(ubar), _d(dsafe), _d(dso),
                             _d(ob), _d(P), _d(q), _d(A), _d(b), _d(lb), _d(ub))
    return P, q, A, b, lb, ub


def qcqp_evaluate(g, cterm, H, qv, gamma0, u, dsafe, params, dsafe_obst=None, obst=None):
    g, cterm, H, qv, gamma0, u, dsafe = _c(g), _c(cterm), _c(H), _c(qv), _c(gamma0), _c(u), _c(dsafe)
    B, nVeh, Hp = g.shape[0], g.shape[1], g.shape[2]
    nObst, dso, ob = _obst(dsafe_obst, obst)
    d = Dims(B, nVeh, Hp, nObst, 2)
    obj, mv, sv, feas = np.zeros(B), np.zeros(B), np.zeros(B), np.zeros(B, dtype=np.int32)
    ci = np.zeros((B, nVeh, nVeh, Hp))
    cio = np.zeros((B, nVeh, max(nObst, 1), Hp))
    lib().emu_qcqp_evaluate(C.byref(d), C.byref(params), _d(g), _d(cterm), _d(H), _d(qv), _d(gamma0), _d(u), _d(dsafe),
                            _d(dso), _d(ob), _d(obj), _d(mv), _d(sv), _i(feas), _d(ci), _d(cio) if nObst else None)
    return dict(obj=obj, max_violation=mv, sum_violations=sv, feasible=feas.astype(bool), ci=ci, ci_obst=cio[:, :, :nObst])


def qp_solve_dense(P, q, A, b, lb, ub, params):
    P, q, A, b, lb, ub = _c(P), _c(q), _c(A), _c(b), _c(lb), _c(ub)
    B, n1, mc = P.shape[0], P.shape[1], A.shape[1]
    d = Dims(B, 1, 1, 0, 2)
    x, fval = np.zeros((B, n1)), np.zeros(B)
    iters, status, zA = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32), np.zeros((B, mc))
    lib().emu_qp_solve_dense(C.byref(d), C.byref(params), C.c_int32(n1), C.c_int32(mc), _d(P), _d(q), _d(A), _d(b), _d(lb),
                             _d(ub), _d(x), _d(fval), _i(iters), _i(status), _d(zA))
    return dict(x=x, fval=fval, iters=iters, status=status, zA=zA)


def scp_solve(g, cterm, H, qv, gamma0, dsafe, u, params, dsafe_obst=None, obst=None, u_prev=None):
    g, cterm, H, qv, gamma0, dsafe = _c(g), _c(cterm), _c(H), _c(qv), _c(gamma0), _c(dsafe)
    B, nVeh, Hp = g.shape[0], g.shape[1], g.shape[2]
    nObst, dso, ob = _obst(dsafe_obst, obst)
    d = Dims(B, nVeh, Hp, nObst, 2)
    n = nVeh * Hp
    u = _c(u).reshape(B, n).copy()
    traj, U = np.zeros((B, Hp, 2, nVeh)), np.zeros((B, Hp, nVeh))
    log = np.zeros((B, params.max_scp_iter, capi.LOG_W))
    si, ii, st = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    obj, mv = np.zeros(B), np.zeros(B)
    up = None if u_prev is None else _c(u_prev).reshape(B, nVeh)
    rc = lib().emu_scp_solve_rate(C.byref(d), C.byref(params), _d(g), _d(cterm), _d(H), _d(qv), _d(gamma0), _d(dsafe), _d(dso),
                                  _d(ob), _d(u), _d(traj), _d(U), _d(log), _i(si), _i(ii), _i(st), _d(obj), _d(mv), _d(up))
    assert rc == 0, "emu_scp_solve_rate: enable_rate_rows needs u_prev"
    return dict(u=u, traj=traj, U=U, log=log, scp_iters=si, ipm_iters=ii, status=st, obj=obj, max_violation=mv)


def advance_linear(abe, U, uMax, duLim, x0, u0):
    """scpb200_advance_linear in NumPy order-preserving form (one vehicle at a time through the kernel's own function is
    not exported; this mirrors scp_advance_vehicle exactly: clamp, then x <- Ad x + Bd u + Ed accumulated left to right)."""
    B, nVeh = u0.shape
    xn, un = x0.copy(), u0.copy()
    for b in range(B):
        for v in range(nVeh):
            ua = min(U[b, 0, v], uMax); ua = max(ua, -uMax)
            ua = min(ua, u0[b, v] + duLim); ua = max(ua, u0[b, v] - duLim)
            a = abe[b, v]
            for i in range(6):
                acc = a[42 + i] + a[36 + i] * ua
                for j in range(6):
                    acc += a[i * 6 + j] * x0[b, v, j]
                xn[b, v, i] = acc
            un[b, v] = ua
    return xn, un


def mpc_rollout(x0, u0, veh, poly, dsafe, Hp, params, nsteps, uMax, duLim, u=None, x_meas=None, u_act=None, mech_limit=0.0,
                lat_acc_limit=0.0, delay=0.0, nsub_delay=144, nsub_plant=64):
    x0, u0, veh, poly, dsafe = _c(x0).copy(), _c(u0).copy(), _c(veh), _c(poly), _c(dsafe)
    B, nVeh = x0.shape[0], x0.shape[1]
    d = Dims(B, nVeh, Hp, 0, poly.shape[2])
    n = nVeh * Hp
    mode = 0 if x_meas is None else 1
    xm = None if x_meas is None else _c(x_meas).copy()
    ua = None if u_act is None else _c(u_act).copy()
    W = dict(ref=np.zeros((B, nVeh, Hp, 2)), g=np.zeros((B, nVeh, Hp, 2)), cterm=np.zeros((B, nVeh, Hp, 2)),
             H=np.zeros((B, nVeh, Hp, Hp)), qv=np.zeros((B, nVeh, Hp)), gamma0=np.zeros(B), abe=np.zeros((B, nVeh, 48)))
    uu = np.zeros((B, n)) if u is None else _c(u).reshape(B, n).copy()
    traj, U = np.zeros((B, Hp, 2, nVeh)), np.zeros((B, Hp, nVeh))
    si, ii, st = (np.zeros(B, dtype=np.int32) for _ in range(3))
    qt, it, so = (np.zeros(B, dtype=np.int32) for _ in range(3))
    sh, sth = np.zeros((B, nsteps), dtype=np.int32), np.zeros((B, nsteps), dtype=np.int32)
    Uh, xh = np.zeros((B, nsteps, Hp, nVeh)), np.zeros((B, nsteps + 1, nVeh, 6))
    lib().emu_mpc_rollout(C.byref(d), C.byref(params), C.c_int(nsteps), C.c_int(mode), _d(veh), _d(poly), _d(dsafe), _d(x0), _d(u0),
                          _d(xm), _d(ua), _d(W["ref"]), _d(W["g"]), _d(W["cterm"]), _d(W["H"]), _d(W["qv"]), _d(W["gamma0"]),
                          _d(W["abe"]), _d(uu), _d(traj), _d(U), _i(si), _i(ii), _i(st), C.c_double(uMax), C.c_double(duLim),
                          C.c_double(mech_limit), C.c_double(lat_acc_limit), C.c_double(delay), C.c_int(nsub_delay),
                          C.c_int(nsub_plant), _i(qt), _i(it), _i(so), _i(sh), _i(sth), _d(Uh), _d(xh))
    return dict(x0=x0, u0=u0, x_meas=xm, u_act=ua, u=uu, U=U, traj=traj, qp_total=qt, ipm_total=it, status_or=so, scp_iters_hist=sh,
                status_hist=sth, U_hist=Uh, x_hist=xh, **W)


def plant_step(x_meas, u_act, veh, U, mech_limit, lat_acc_limit, duLim, T, nsub, params):
    """returns (x_next, u_next, uMax, U_clamped)"""
    x, ua, veh, U = _c(x_meas).copy(), _c(u_act).copy(), _c(veh), _c(U)
    B, Hp, nVeh = U.shape
    d = Dims(B, nVeh, Hp, 0, 2)
    umax, uc = np.zeros((B, nVeh)), np.zeros((B, Hp, nVeh))
    lib().emu_plant_step(C.byref(d), C.byref(params), _d(veh), _d(U), C.c_double(mech_limit), C.c_double(lat_acc_limit),
                         C.c_double(duLim), C.c_double(T), C.c_int32(nsub), _d(x), _d(ua), _d(umax), _d(uc))
    return x, ua, umax, uc


def ode_predict(x, u_ref, veh, T, steps, nsub, params):
    x, u_ref, veh = _c(x), _c(u_ref), _c(veh)
    B, nVeh = x.shape[0], x.shape[1]
    d = Dims(B, nVeh, 1, 0, 2)
    out = np.zeros((B, nVeh, steps, 6))
    lib().emu_ode_predict(C.byref(d), C.byref(params), _d(x), _d(u_ref), _d(veh), C.c_double(T), C.c_int32(steps),
                          C.c_int32(nsub), _d(out))
    return out
