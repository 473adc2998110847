"""Shared test helpers (TEST INFRASTRUCTURE ONLY): closed-form known-answer QPs, a pure-NumPy restatement of the
reference's QP data, an independent Philox4x32-10, and the KKT certificate in extended precision.

Nothing here calls oracle/scp_oracle.c: these are the checks that certify the oracle (and the CUDA path) from outside
("analytic known-answer QPs (box-only, single active row, omega-active) with closed-form solutions",
"certified by KKT residuals rather than by trust")."""
from __future__ import annotations

import numpy as np

LD = np.longdouble


# ------------------------------------------------------------------------------------------------ known-answer QPs
def kat_box_only(n=9, seed=1):
    """min 1/2 x'Dx + q'x, lb <= x <= ub with D diagonal: x_i = clip(-q_i / D_i, lb_i, ub_i).  One inactive general row
    keeps the (mc >= 1) code paths busy; `kat_box_only_no_rows` is the mc = 0 edge case."""
    rng = np.random.default_rng(seed)
    D = rng.uniform(0.5, 50.0, n)
    q = rng.uniform(-40.0, 40.0, n)
    lb, ub = -rng.uniform(0.1, 1.0, n), rng.uniform(0.1, 1.0, n)
    x = np.clip(-q / D, lb, ub)
    A = np.ones((1, n))
    b = np.array([n * 10.0])                                       # a'x <= 10 n: never active inside the box
    return dict(name="box_only", P=np.diag(D), q=q, A=A, b=b, lb=lb, ub=ub, x=x, fval=float(0.5 * x @ (D * x) + q @ x))


def kat_box_only_no_rows(n=5, seed=2):
    d = kat_box_only(n, seed)
    d.update(name="box_only_no_rows", A=np.zeros((0, n)), b=np.zeros(0))
    return d


def kat_single_active_row(n=12, seed=3):
    """min 1/2 |x - c|^2 s.t. a'x <= b (box far away): the projection of c on the half space,
    x = c - a (a'c - b) / |a|^2, multiplier (a'c - b) / |a|^2."""
    rng = np.random.default_rng(seed)
    c = rng.normal(size=n)
    a = rng.normal(size=n)
    b = float(a @ c - 0.7 * np.linalg.norm(a))                      # c violates the row by 0.7 |a|
    lam = (a @ c - b) / (a @ a)
    x = c - lam * a
    return dict(name="single_active_row", P=np.eye(n), q=-c, A=a[None], b=np.array([b]), lb=np.full(n, -50.0),
                ub=np.full(n, 50.0), x=x, fval=float(0.5 * x @ x - c @ x), zA=np.array([lam]))


def kat_omega_active(n=10, w=1e5, seed=4):
    """The SCP QP's shape (SCP_controller.py:118-127) in closed form: variables (u_1..u_n, omega), cost
    sum(1/2 p_i u_i^2 + q_i u_i) + w omega, rows  u_i - omega <= -1  and  -u_i - omega <= -1  (infeasible without
    slack), omega in [0, 1e25], |u_i| <= 3.  omega = 1 + max|u_i|; with sum|q_i| <= w the minimiser is u = 0,
    omega = 1, cost w (every row active: the degenerate vertex the reference's conflict steps sit at)."""
    rng = np.random.default_rng(seed)
    p = rng.uniform(1.0, 10.0, n)
    q = rng.uniform(-1.0, 1.0, n) * (0.5 * w / n)                   # sum |q| <= w / 2
    P = np.zeros((n + 1, n + 1))
    P[:n, :n] = np.diag(p)
    A = np.zeros((2 * n, n + 1))
    A[:n, :n], A[n:, :n] = np.eye(n), -np.eye(n)
    A[:, n] = -1.0
    b = -np.ones(2 * n)
    lb = np.concatenate([np.full(n, -3.0), [0.0]])
    ub = np.concatenate([np.full(n, 3.0), [1e25]])
    x = np.concatenate([np.zeros(n), [1.0]])
    return dict(name="omega_active", P=P, q=np.concatenate([q, [w]]), A=A, b=b, lb=lb, ub=ub, x=x, fval=float(w))


def kat_omega_active_1d(pw=4.0, q=-3.0e5, w=1e5):
    """One steering variable, |q| > w: u = -sign(q)(|q| - w)/p, omega = 1 + |u|."""
    u = -np.sign(q) * (abs(q) - w) / pw
    om = 1.0 + abs(u)
    P = np.array([[pw, 0.0], [0.0, 0.0]])
    A = np.array([[1.0, -1.0], [-1.0, -1.0]])
    b = -np.ones(2)
    big = 10.0 * abs(u) + 10.0
    return dict(name="omega_active_1d", P=P, q=np.array([q, w]), A=A, b=b, lb=np.array([-big, 0.0]), ub=np.array([big, 1e25]),
                x=np.array([u, om]), fval=float(0.5 * pw * u * u + q * u + w * om))


KAT_CASES = [kat_box_only, kat_box_only_no_rows, kat_single_active_row, kat_omega_active, kat_omega_active_1d]


# ------------------------------------------------------------------------------------------------ KKT certificate
def kkt_certificate(P, q, A, b, lb, ub, x, zA, zub, zlb, inf_bound=1e20):
    """KKT residuals of (x, z) for  min 1/2 x'Px + q'x, Ax <= b, lb <= x <= ub  in NumPy longdouble (80-bit on x86).
    Independent of the oracle's C code.  Returns absolute residuals and the bound on |u - u*|_2 they imply."""
    P, q, A, b, lb, ub, x, zA, zub, zlb = (np.asarray(a, dtype=LD) for a in (P, q, A, b, lb, ub, x, zA, zub, zlb))
    hasu, hasl = np.abs(ub) < inf_bound, np.abs(lb) < inf_bound
    zub, zlb = np.where(hasu, zub, 0), np.where(hasl, zlb, 0)
    r = P @ x + q + (A.T @ zA if A.shape[0] else 0) + zub - zlb                          # stationarity
    sA = b - A @ x if A.shape[0] else np.zeros(0, dtype=LD)
    sU, sL = np.where(hasu, ub - x, 1), np.where(hasl, x - lb, 1)
    primal = max(float(np.max(-sA, initial=0.0)), float(np.max(-sU)), float(np.max(-sL)), 0.0)
    dual = max(float(np.max(-zA, initial=0.0)), float(np.max(-zub)), float(np.max(-zlb)), 0.0)
    gap = float(np.abs(zA * sA).sum() + np.abs(zub * sU).sum() + np.abs(zlb * sL).sum())
    comp = max(float(np.max(np.abs(zA * sA), initial=0.0)), float(np.max(np.abs(zub * sU))), float(np.max(np.abs(zlb * sL))))
    return dict(stationarity=float(np.max(np.abs(r))), stationarity_2=float(np.sqrt((r * r).sum())), primal=primal,
                dual=dual, gap=gap, complementarity=comp)


def minimiser_distance_bound(cert, lam_min, radius_u, row_norm_max, w_omega, mrows):
    """Distance |u - u*|_2 to the exact minimiser that a KKT certificate guarantees.  For a feasible x and z >= 0:
        f(x*) >= L(x*, z) >= f(x) - gap + r'(x* - x) + 1/2 (x* - x)'P(x* - x)   and   f(x*) <= f(x),
    hence  lam_min/2 |du|^2 <= gap + |r|_2 |dx|_2  (lam_min = smallest eigenvalue of P on the u-block; omega has no
    curvature).  |dx| <= |du| + |d omega|, and omega tracks u on both sides: omega* = max(0, max_r(a_r'u* - b_r)) exactly,
    while for x complementarity pins omega to within gap * mrows / w of max_r(a_r'u - b_r) (some row carries a multiplier
    >= (w - |r|)/mrows), so |d omega| <= row_norm_max |du| + eps.  Start from the box diameter and iterate the
    inequality to its fixed point."""
    eps_om = cert["gap"] * mrows / max(w_omega - cert["stationarity"], 1.0) + cert["primal"]
    du = radius_u
    for _ in range(60):
        dx = (1.0 + row_norm_max) * du + eps_om
        du = min(du, float(np.sqrt(2.0 * (cert["gap"] + cert["stationarity_2"] * dx) / lam_min)))
    return du


# ------------------------------------------------------------------------------------------------ QP data in NumPy
def qp_from_reference_arrays(G, ubar, trust_radius=np.inf):
    """The QP of SCP_controller.py:93-128 about `ubar`, built in NumPy from the REFERENCE's own MPCclass arrays stored in
    a golden record (Mathcal_B, const_term, Phi_0, Psi_0): rows via the structured identity of SURVEY 8(a7).  Checked
    against the dense (P, q, Aineq, bineq) the reference logged wherever a record holds them."""
    nVeh, Hp = int(G["sc_nVeh"]), int(G["sc_Hp"])
    n = nVeh * Hp
    MB, ct = G["Mathcal_B"], G["const_term"]                       # [2Hp,Hp,nVeh], [2Hp,nVeh]
    ubar = np.asarray(ubar, float).ravel()
    pos = np.stack([ct[:, v] + MB[:, :, v] @ ubar[v * Hp:(v + 1) * Hp] for v in range(nVeh)])     # [nVeh, 2Hp]
    sbar = G["sc_dsafeVehicles"] + float(G["sc_dsafeExtra"])
    rows, rhs = [], []
    for i in range(nVeh):
        for j in range(i + 1, nVeh):
            for k in range(Hp):
                d = pos[i, 2 * k:2 * k + 2] - pos[j, 2 * k:2 * k + 2]
                c = ct[2 * k:2 * k + 2, i] - ct[2 * k:2 * k + 2, j]
                a = np.zeros(n + 1)
                a[i * Hp:(i + 1) * Hp] = -2.0 * d @ MB[2 * k:2 * k + 2, :, i]
                a[j * Hp:(j + 1) * Hp] = 2.0 * d @ MB[2 * k:2 * k + 2, :, j]
                a[np.abs(a) <= 1e-20] = 0.0
                a[n] = -1.0
                rows.append(a)
                rhs.append(-sbar[i, j] ** 2 - d @ d + 2.0 * d @ c)
    P = np.zeros((n + 1, n + 1))
    q = np.zeros(n + 1)
    for v in range(nVeh):
        P[v * Hp:(v + 1) * Hp, v * Hp:(v + 1) * Hp] = 2.0 * G["Phi_0"][:, :, v]
        q[v * Hp:(v + 1) * Hp] = G["Psi_0"][:, v]
    q[n] = 1e5
    uLim = float(G["sc_uLim"])
    lb = np.concatenate([np.maximum(-uLim, ubar - trust_radius), [0.0]])
    ub = np.concatenate([np.minimum(uLim, ubar + trust_radius), [1e25]])
    return P, q, np.array(rows), np.array(rhs), lb, ub


# ------------------------------------------------------------------------------------------------ Philox4x32-10
#: Random123 known-answer vectors for philox4x32, 10 rounds (kat_vectors of the Random123 distribution):
#: (counter[4], key[2]) -> output[4]
PHILOX_KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10_py(ctr, key):
    """Philox4x32-10 from the published algorithm (Salmon, Moraes, Dror, Shaw, SC'11), plain Python integers."""
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c = [int(v) & MASK for v in ctr]
    k = [int(v) & MASK for v in key]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k[1]) & MASK, p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return tuple(c)


def noise_pair_py(seed, instance, vehicle, counter, stream=0):
    """The N(0,1) pair the kernels draw for (seed; instance, vehicle, counter, stream): Philox words -> two 53-bit
    uniforms -> Box-Muller (scp_kernels.cuh: scp_noise_pair)."""
    import math
    r = philox4x32_10_py((instance, vehicle, counter, (0x5C9B200 + stream) & 0xFFFFFFFF), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u1 = (float((r[0] << 21) ^ (r[1] >> 11)) + 1.0) / 9007199254740992.0
    u2 = float((r[2] << 21) ^ (r[3] >> 11)) / 9007199254740992.0
    rad, ang = math.sqrt(-2.0 * math.log(u1)), 6.283185307179586476925286766559 * u2
    return rad * math.cos(ang), rad * math.sin(ang)
