// ops_pair.cuh — the constraint operator of the fused SCP kernel: linearised vehicle-pair and obstacle rows
// in their rank structure (never materialised).
//
// SCP_controller.py:308-317 builds, per (pair i<j, step k), Phi = -M'M with M = [B_i(k), -B_j(k)] (2 x n),
// and :100-101 the row Psi' + 2 ubar'Phi.  With dbar = p_i(k) - p_j(k) at the linearisation point that row is
//     A[r, i*Hp + a] = -2 dbar . g_i[k-a],   A[r, j*Hp + a] = +2 dbar . g_j[k-a]   (a <= k),   A[r, n] = -1
// (g_v[l] = C A^l B, the Toeplitz generator of Mathcal_B, MPC_Iter.py:146-147).  Hence
//     A x      = -2 dbar . (resp_i(k) - resp_j(k)) - omega,     resp_v(k) = sum_{a<=k} g_v[k-a] x_v[a]
//     A' w     : F_v(k) = sum_{rows of v at k} (-/+ 2) w_r dbar_r ;  (A'w)[v,a] = sum_{k>=a} g_v[k-a] . F_v(k)
//     A' D A   : entry ((i,a),(j,b)) = sum_{k>=max(a,b)} D_r coef_r(i,a) coef_r(j,b) over the rows shared by i and j
// Obstacle rows (SCP_controller.py:106-114, :321-326) have a single block: A[r, v*Hp+a] = -2 dbar . g_v[k-a].
// P = blkdiag(2 Phi_0, 0) (SCP_controller.py:120,124) with Phi_0 = H from K1.
#pragma once
#include "scp_common.cuh"

struct PairOp {
    int nVeh, Hp, n, nObst, mcv, mc;   // mcv = vehicle-pair rows, mc = all rows
    const double *g;      // [nVeh][Hp][2]   (shared)
    const double *H;      // [nVeh][Hp][Hp]  (global, read-only)
    const double *dbar;   // [mc][2]         (shared)
    double *resp;         // [n][2] scratch  (shared)
    double *red;          // reduction scratch

    SCP_MFN int pair_index(int i, int j) const { return i * nVeh - (i * (i + 1) >> 1) + (j - i - 1); }

    SCP_MFN void mul_P(Cta &cta, const double *x, double *y) const
    {
        CTA_PHASE(tid)
            for (int c = tid; c <= n; c += cta.nt) {
                double acc = 0.0;
                if (c < n) {
                    const int v = c / Hp;
                    const double *Hr = H + (size_t)c * Hp;
                    const double *xv = x + v * Hp;
                    for (int b = 0; b < Hp; ++b) acc += Hr[b] * xv[b];
                    acc *= 2.0;
                }
                y[c] = acc;
            }
        CTA_PHASE_END
    }

    SCP_MFN void add_P(Cta &cta, double *S) const
    {
        CTA_PHASE(tid)
            const int per = Hp * (Hp + 1) >> 1;
            for (int e = tid; e < nVeh * per; e += cta.nt) {
                const int v = e / per;
                int a, b;
                scp_tri_decode(e - v * per, &a, &b);
                S[scp_sidx(v * Hp + a, v * Hp + b)] += 2.0 * H[((size_t)v * Hp + a) * Hp + b];
            }
        CTA_PHASE_END
    }

    // resp[(v,k)] = sum_{a<=k} g_v[k-a] x[v*Hp+a]
    SCP_MFN void response(Cta &cta, const double *x) const
    {
        CTA_PHASE(tid)
            for (int c = tid; c < n; c += cta.nt) {
                const int v = c / Hp, k = c - v * Hp;
                const double *gv = g + (size_t)v * Hp * 2;
                const double *xv = x + v * Hp;
                double rx = 0.0, ry = 0.0;
                for (int a = 0; a <= k; ++a) {
                    rx += gv[(k - a) * 2] * xv[a];
                    ry += gv[(k - a) * 2 + 1] * xv[a];
                }
                resp[c * 2] = rx;
                resp[c * 2 + 1] = ry;
            }
        CTA_PHASE_END
    }

    SCP_MFN void mul_A(Cta &cta, const double *x, double *y) const
    {
        response(cta, x);
        CTA_PHASE(tid)
            const double om = x[n];
            for (int r = tid; r < mc; r += cta.nt) {
                double dx, dy;
                if (r < mcv) {
                    const int p = r / Hp, k = r - p * Hp;
                    int i = 0, rem = p;                       // p -> (i, j)
                    while (rem >= nVeh - 1 - i) { rem -= nVeh - 1 - i; ++i; }
                    const int j = i + 1 + rem;
                    dx = resp[(i * Hp + k) * 2] - resp[(j * Hp + k) * 2];
                    dy = resp[(i * Hp + k) * 2 + 1] - resp[(j * Hp + k) * 2 + 1];
                } else {
                    const int q = r - mcv, v = q / (nObst * Hp), k = q % Hp;
                    dx = resp[(v * Hp + k) * 2];
                    dy = resp[(v * Hp + k) * 2 + 1];
                }
                y[r] = -2.0 * (dbar[r * 2] * dx + dbar[r * 2 + 1] * dy) - om;
            }
        CTA_PHASE_END
    }

    // resp[(v,k)] = F_v(k) = sum over the rows of vehicle v at step k of sign * 2 * w_r * dbar_r ; returns sum w
    SCP_MFN double forces(Cta &cta, const double *w) const
    {
        CTA_RED_BEGIN(cta, 1)
        CTA_PHASE(tid)
            double sw = 0.0;
            for (int r = tid; r < mc; r += cta.nt) sw += w[r];
            for (int c = tid; c < n; c += cta.nt) {
                const int v = c / Hp, k = c - v * Hp;
                double fx = 0.0, fy = 0.0;
                for (int o = 0; o < nVeh; ++o) {
                    if (o == v) continue;
                    const int i = v < o ? v : o, j = v < o ? o : v;
                    const int r = pair_index(i, j) * Hp + k;
                    const double sw2 = (v == i ? -2.0 : 2.0) * w[r];
                    fx += sw2 * dbar[r * 2];
                    fy += sw2 * dbar[r * 2 + 1];
                }
                for (int o = 0; o < nObst; ++o) {
                    const int r = mcv + (v * nObst + o) * Hp + k;
                    fx -= 2.0 * w[r] * dbar[r * 2];
                    fy -= 2.0 * w[r] * dbar[r * 2 + 1];
                }
                resp[c * 2] = fx;
                resp[c * 2 + 1] = fy;
            }
            CTA_RED_SUM(cta, red, 0, tid, sw)
        CTA_PHASE_END_RED(cta, red, 1)
        return cta_red_sum(cta, red, 0);
    }

    SCP_MFN void add_At(Cta &cta, const double *w, double *vout) const
    {
        const double sw = forces(cta, w);
        CTA_PHASE(tid)
            for (int c = tid; c < n; c += cta.nt) {
                const int v = c / Hp, a = c - v * Hp;
                const double *gv = g + (size_t)v * Hp * 2;
                double acc = 0.0;
                for (int k = a; k < Hp; ++k)
                    acc += gv[(k - a) * 2] * resp[(v * Hp + k) * 2] + gv[(k - a) * 2 + 1] * resp[(v * Hp + k) * 2 + 1];
                vout[c] += acc;
            }
            if (tid == 0) vout[n] -= sw;
        CTA_PHASE_END
    }

    // coefficient of row r (step k) on u[v, a] without its sign:  2 dbar_r . g_v[k-a]
    SCP_MFN double coef2(const double *gv, int r, int l) const
    {
        return 2.0 * (dbar[r * 2] * gv[l * 2] + dbar[r * 2 + 1] * gv[l * 2 + 1]);
    }

    SCP_MFN void add_AtDA(Cta &cta, const double *dd, double *S) const
    {
        const double sd = forces(cta, dd);            // resp = A'dd in force form (for the omega row)
        CTA_PHASE(tid)
            // omega row: S[n][(v,a)] -= (A'dd)[(v,a)] ; S[n][n] += sum dd
            for (int c = tid; c < n; c += cta.nt) {
                const int v = c / Hp, a = c - v * Hp;
                const double *gv = g + (size_t)v * Hp * 2;
                double acc = 0.0;
                for (int k = a; k < Hp; ++k)
                    acc += gv[(k - a) * 2] * resp[(v * Hp + k) * 2] + gv[(k - a) * 2 + 1] * resp[(v * Hp + k) * 2 + 1];
                S[scp_sidx(n, c)] -= acc;
            }
            if (tid == 0) S[scp_sidx(n, n)] += sd;
            // u block, lower triangle
            const int tot = n * (n + 1) >> 1;
            for (int e = tid; e < tot; e += cta.nt) {
                int ci, cj;
                scp_tri_decode(e, &ci, &cj);
                const int i = ci / Hp, a = ci - i * Hp, j = cj / Hp, b = cj - j * Hp;
                const double *gi = g + (size_t)i * Hp * 2, *gj = g + (size_t)j * Hp * 2;
                double acc = 0.0;
                if (i == j) {                          // a >= b
                    for (int o = 0; o < nVeh; ++o) {
                        if (o == i) continue;
                        const int r0 = (i < o ? pair_index(i, o) : pair_index(o, i)) * Hp;
                        for (int k = a; k < Hp; ++k)
                            acc += dd[r0 + k] * coef2(gi, r0 + k, k - a) * coef2(gi, r0 + k, k - b);
                    }
                    for (int o = 0; o < nObst; ++o) {
                        const int r0 = mcv + (i * nObst + o) * Hp;
                        for (int k = a; k < Hp; ++k)
                            acc += dd[r0 + k] * coef2(gi, r0 + k, k - a) * coef2(gi, r0 + k, k - b);
                    }
                } else {                               // i > j : rows of pair (j, i); signs -(j) and +(i)
                    const int r0 = pair_index(j, i) * Hp;
                    for (int k = (a > b ? a : b); k < Hp; ++k)
                        acc -= dd[r0 + k] * coef2(gi, r0 + k, k - a) * coef2(gj, r0 + k, k - b);
                }
                S[scp_sidx(ci, cj)] += acc;
            }
        CTA_PHASE_END
    }
};
