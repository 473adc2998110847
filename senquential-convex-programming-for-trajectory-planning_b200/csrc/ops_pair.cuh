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
    double *resp;         // [n][2] scratch: response of the last prep'd x   (shared)
    double *frc;          // [n][2] scratch: forces of the last prep'd w     (shared)
    double xom, wsum;     // omega component of x / sum of w of the last prep
    double *red;          // reduction scratch
    double *Msm;          // [n][3] scratch: per (vehicle, step) sum of 4 dd dbar dbar' (xx, xy, yy)
    const int *rowtab;    // [mc] (i*Hp + k) | (j*Hp + k) << 16 per row (shared; built once per instance)
    bool coh;             // H was written during this launch (rollout entry): read it past L1
    int alpha_slots;      // pair-block mode: > 0 tensor path (pair_block_mma, Hp <= 64), 0 entry by entry

    SCP_MFN int pair_index(int i, int j) const { return i * nVeh - (i * (i + 1) >> 1) + (j - i - 1); }

    // (P x)[c],  P = blkdiag(2H, 0)
    SCP_MFN double P_col(int c, const double *x) const
    {
        if (c >= n) return 0.0;
        const int v = c / Hp;
        const double *Hr = H + (size_t)c * Hp, *xv = x + v * Hp;
        double a0 = 0.0, a1 = 0.0;
        int b = 0;
        for (; b + 1 < Hp; b += 2) { a0 += scp_ldc(Hr + b, coh) * xv[b]; a1 += scp_ldc(Hr + b + 1, coh) * xv[b + 1]; }
        if (b < Hp) a0 += scp_ldc(Hr + b, coh) * xv[b];
        return 2.0 * (a0 + a1);
    }

    // One phase: resp[(v,k)] = sum_{a<=k} g_v[k-a] x[v*Hp+a] (if x) and frc[(v,k)] = F_v(k) = sum over the rows of
    // vehicle v at step k of sign * 2 * w_r * dbar_r, wsum = sum w (if w).  Afterwards row_dot / col_dot are inline.
    SCP_MFN void prep(Cta &cta, const double *x, const double *w)
    {
        xom = x ? x[n] : 0.0;
        CTA_RED_BEGIN(cta, 1)
        CTA_PHASE(tid)
            double sw = 0.0;
            if (w)
                for (int r = tid; r < mc; r += cta.nt) sw += w[r];
            for (int c = tid; c < n; c += cta.nt) {
                const int v = c / Hp, k = c - v * Hp;
                if (x) {
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
                if (w) {
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
                    frc[c * 2] = fx;
                    frc[c * 2 + 1] = fy;
                }
            }
            CTA_RED_SUM(cta, red, 0, tid, sw)
        CTA_PHASE_END_RED(cta, red, 1)
        wsum = cta_red_sum(cta, red, 0);
    }

    // (A x)[r] for the x of the last prep
    SCP_MFN double row_dot(int r) const
    {
        double dx, dy;
        const int t = rowtab[r], ci = t & 0xffff;
        if (r < mcv) {
            const int cj = t >> 16;
            dx = resp[ci * 2] - resp[cj * 2];
            dy = resp[ci * 2 + 1] - resp[cj * 2 + 1];
        } else {
            dx = resp[ci * 2];
            dy = resp[ci * 2 + 1];
        }
        return -2.0 * (dbar[r * 2] * dx + dbar[r * 2 + 1] * dy) - xom;
    }

    // (A' w)[c] for the w of the last prep
    SCP_MFN double col_dot(int c) const
    {
        if (c >= n) return -wsum;
        const int v = c / Hp, a = c - v * Hp;
        const double *gv = g + (size_t)v * Hp * 2;
        double acc = 0.0;
        for (int k = a; k < Hp; ++k)
            acc += gv[(k - a) * 2] * frc[(v * Hp + k) * 2] + gv[(k - a) * 2 + 1] * frc[(v * Hp + k) * 2 + 1];
        return acc;
    }

    // resp[(v,k)] = F_v(k) for weights w (used by form_normal for the omega row); returns sum w
    SCP_MFN double forces(Cta &cta, const double *w)
    {
        prep(cta, (const double *)0, w);
        return wsum;
    }

    // coefficient of row r (step k) on u[v, a] without its sign:  2 dbar_r . g_v[k-a]
    SCP_MFN double coef2(const double *gv, int r, int l) const
    {
        return 2.0 * (dbar[r * 2] * gv[l * 2] + dbar[r * 2 + 1] * gv[l * 2 + 1]);
    }

    // One pair block on the FP64 tensor path, operands generated in registers (no scratch, no intra-warp exchange):
    //     S[(j,b),(i,a)] = -sum_k cj[k][b] (dd_r ci[k][a]),   cv[k][l] = 2 dbar_r . g_v[k-l]  (l <= k, else 0),  r = r0 + k
    // i.e. C = -Aj' (D Ai) with Hp x Hp causal factors.  Per m8n8k4 step lane l supplies A[row b = l>>2][col k = l&3] and
    // B[row k = l&3][col a = l>>2] and owns C[b = l>>2][a = 2(l&3), 2(l&3)+1].  Causality (k >= max(a, b)) skips the
    // k-steps below the tile diagonal: Hp = 10 needs 6 DMMAs per pair.
#define SCP_PAIR_NA 8      /* a-tiles held in accumulators: Hp <= 64 */
    SCP_MFN void pair_block_mma(int lane, double *S, int i, int j, int r0, const double *dd) const
    {
        const int nt = (Hp + 7) >> 3;
        if (nt <= 2) pair_block_mma_n<2>(lane, S, i, j, r0, dd);          // the unrolled a-tile loop matches the horizon
        else if (nt <= 4) pair_block_mma_n<4>(lane, S, i, j, r0, dd);
        else pair_block_mma_n<SCP_PAIR_NA>(lane, S, i, j, r0, dd);
    }
    template <int NA>
    SCP_MFN void pair_block_mma_n(int lane, double *S, int i, int j, int r0, const double *dd) const
    {
        const double *gi = g + (size_t)i * Hp * 2, *gj = g + (size_t)j * Hp * 2;
#if SCP_DEVICE_BUILD
        const double2 *__restrict__ gi2 = reinterpret_cast<const double2 *>(gi);
        const double2 *__restrict__ gj2 = reinterpret_cast<const double2 *>(gj);
        const double2 *__restrict__ db2 = reinterpret_cast<const double2 *>(dbar) + r0;
        const double *__restrict__ ddr = dd + r0;
        const int q = lane >> 2, kq = lane & 3;
        const int nt = (Hp + 7) >> 3, nk = (Hp + 3) >> 2;
        for (int tb = 0; tb < nt; ++tb) {
            double acc[NA][2];
#pragma unroll
            for (int ta = 0; ta < NA; ++ta) acc[ta][0] = acc[ta][1] = 0.0;
            const int b = 8 * tb + q;
            for (int s = 2 * tb; s < nk; ++s) {
                const int k = 4 * s + kq;
                double dxk = 0.0, dyk = 0.0, ddk = 0.0;
                if (k < Hp) { const double2 t = db2[k]; dxk = 2.0 * t.x; dyk = 2.0 * t.y; ddk = ddr[k]; }
                double ae = 0.0;
                if (k < Hp && b <= k) { const double2 t = gj2[k - b]; ae = dxk * t.x + dyk * t.y; }
                dxk *= ddk; dyk *= ddk;
#pragma unroll
                for (int ta = 0; ta < NA; ++ta)
                    if (ta < nt && 2 * ta <= s) {                       // warp-uniform
                        const int a = 8 * ta + q;
                        double be = 0.0;
                        if (k < Hp && a <= k) { const double2 t = gi2[k - a]; be = dxk * t.x + dyk * t.y; }
                        scp_dmma(acc[ta][0], acc[ta][1], ae, be);
                    }
            }
            if (b < Hp) {
#pragma unroll
                for (int ta = 0; ta < NA; ++ta)
                    if (ta < nt) {
                        const int a0 = 8 * ta + 2 * kq;
                        if (a0 < Hp) S[scp_sidx(j * Hp + b, i * Hp + a0)] = -acc[ta][0];
                        if (a0 + 1 < Hp) S[scp_sidx(j * Hp + b, i * Hp + a0 + 1)] = -acc[ta][1];
                    }
            }
        }
#else
        if (lane == 0)
            for (int b = 0; b < Hp; ++b)
                for (int a = 0; a < Hp; ++a) {
                    double acc = 0.0;
                    for (int k = (a > b ? a : b); k < Hp; ++k)
                        acc += coef2(gj, r0 + k, k - b) * (dd[r0 + k] * coef2(gi, r0 + k, k - a));
                    S[scp_sidx(j * Hp + b, i * Hp + a)] = -acc;
                }
#endif
    }

    // One diagonal (vehicle) block on the FP64 tensor path:
    //     S[(v,a),(v,b)] = 2 H_v[a][b] + dg delta_ab + sum_{k >= max(a,b)} g_v[k-a]' M_v(k) g_v[k-b]        (a >= b)
    // as C = G' (M G) with the contraction index kappa = 2k + c over (step, component): lane l supplies
    // A[a = l>>2][kappa] = g_v[k-a][c] and B[kappa][b = l>>2] = (M_v(k) g_v[k-b])[c] (two multiply-adds from the
    // aggregated 2x2 blocks), and owns C[a = l>>2][b = 2(l&3), 2(l&3)+1].  Causality skips the kappa-steps below the tile
    // row: Hp = 10 needs 7 DMMAs per vehicle (the entry-by-entry loop it replaces was ~40 % of form_normal).
    // rsub (or null): sub-diagonal additions of the difference rows, S[(v,a),(v,a-1)] += rsub[v*Hp + a].
    SCP_MFN void diag_block_mma(int lane, double *S, int v, const double *dg, const double *rsub) const
    {
        const double *gv = g + (size_t)v * Hp * 2;
        const double *Mv = Msm + (size_t)v * Hp * 3;
        const double *Hv = H + (size_t)v * Hp * Hp;
#if SCP_DEVICE_BUILD
        const double2 *__restrict__ gv2 = reinterpret_cast<const double2 *>(gv);
        const int q = lane >> 2, kq = lane & 3, c = kq & 1, kh = kq >> 1;
        const int nt = (Hp + 7) >> 3, nks = (2 * Hp + 3) >> 2;
        for (int ta = 0; ta < nt; ++ta) {
            const int a = 8 * ta + q;
            for (int tb = 0; tb <= ta; ++tb) {
                const int b = 8 * tb + q;
                double c0 = 0.0, c1 = 0.0;
                for (int s = 4 * ta; s < nks; ++s) {
                    const int k = 2 * s + kh;
                    double ae = 0.0, be = 0.0;
                    if (k < Hp) {
                        if (a <= k) ae = gv[(k - a) * 2 + c];
                        if (b <= k) { const double2 t = gv2[k - b]; be = Mv[k * 3 + c] * t.x + Mv[k * 3 + c + 1] * t.y; }
                    }
                    scp_dmma(c0, c1, ae, be);
                }
                const int b0 = 8 * tb + 2 * kq;
                if (a < Hp) {
                    if (rsub) {
                        if (a == b0 + 1) c0 += rsub[v * Hp + a];
                        if (a == b0 + 2) c1 += rsub[v * Hp + a];
                    }
                    if (b0 <= a) S[scp_sidx(v * Hp + a, v * Hp + b0)] = c0 + 2.0 * scp_ldc(Hv + a * Hp + b0, coh) + (a == b0 ? dg[v * Hp + a] : 0.0);
                    if (b0 + 1 <= a) S[scp_sidx(v * Hp + a, v * Hp + b0 + 1)] = c1 + 2.0 * scp_ldc(Hv + a * Hp + b0 + 1, coh) + (a == b0 + 1 ? dg[v * Hp + a] : 0.0);
                }
            }
        }
#else
        if (lane == 0)
            for (int a = 0; a < Hp; ++a)
                for (int b = 0; b <= a; ++b) {
                    double acc = 0.0;
                    for (int k = a; k < Hp; ++k) {
                        const double gbx = gv[(k - b) * 2], gby = gv[(k - b) * 2 + 1];
                        const double bx = Mv[k * 3] * gbx + Mv[k * 3 + 1] * gby, by = Mv[k * 3 + 1] * gbx + Mv[k * 3 + 2] * gby;
                        acc += gv[(k - a) * 2] * bx + gv[(k - a) * 2 + 1] * by;
                    }
                    if (rsub && a == b + 1) acc += rsub[v * Hp + a];
                    S[scp_sidx(v * Hp + a, v * Hp + b)] = acc + 2.0 * scp_ldc(Hv + a * Hp + b, coh) + (a == b ? dg[v * Hp + a] : 0.0);
                }
#endif
    }

    // S(lower) = blkdiag(2H, 0) + A' diag(dd) A + diag(dg), every entry written exactly once (no clear, no
    // read-modify-write):
    //   diagonal blocks   S[(v,a),(v,b)] = 2H_v[a][b] + dg + sum_{k>=a} g_v[k-a]' M_v(k) g_v[k-b],
    //                     M_v(k) = sum over the rows of v at step k of 4 dd_r dbar_r dbar_r'   (2x2, aggregated)
    //   omega row         S[n][(v,a)] = -(A'dd)[(v,a)],  S[n][n] = sum dd + dg[n]
    //   pair blocks (j>i) S[(j,b),(i,a)] = -sum_{k>=max(a,b)} aj[k][b] ai[k][a],  ai = 2 dd_r dbar_r.g_i[k-a],
    //                     aj = 2 dbar_r.g_j[k-b]: one warp per block on the tensor path (pair_block_mma).
    // Two phases: (1) per (vehicle, step): forces of dd (for the omega row), M_v(k), sum dd, padding; (2) every entry of S
    // — the pair blocks by warps, then omega row and diagonal blocks by threads, without a barrier in between (they
    // read the same inputs and write disjoint entries).
    // The scalings come as functions (ddf(r): row scaling, dgf(c): box term of the diagonal) evaluated inside phase 1 —
    // no separate pass over the rows; dd[r] and dg[c] are stored for phase 2.
    // rhs_row (n1 entries, or null): stored as row n1p - 1 of S with a huge diagonal, so that the factorisation
    // forward-substitutes it on the way (chol_factor).
    template <class Mem, class DD, class DG>
    SCP_MFN void form_normal(Cta &cta, const Mem &m, double *dd, double *dg, const double *rhs_row, DD ddf, DG dgf SCP_TIMER_ARG)
    {
        double *S = m.S;
        const double *rsub = m.nr > 0 ? m.rsub : (const double *)0;     // written by dgf in phase 1, read in phase 2
        CTA_RED_BEGIN(cta, 1)
        CTA_PHASE(tid)
            double sw = 0.0;
            for (int c = tid; c < m.n1p; c += cta.nt) dg[c] = dgf(c);
            for (int c = tid; c < n; c += cta.nt) {
                const int v = c / Hp, k = c - v * Hp;
                double fx = 0.0, fy = 0.0, mxx = 0.0, mxy = 0.0, myy = 0.0;
                for (int o = 0; o < nVeh; ++o) {
                    if (o == v) continue;
                    const int r = (v < o ? pair_index(v, o) : pair_index(o, v)) * Hp + k;
                    const double d = ddf(r);
                    const double w2 = 2.0 * d, dx = dbar[r * 2], dy = dbar[r * 2 + 1];
                    const double wx = w2 * dx, wy = w2 * dy;
                    if (v < o) { fx -= wx; fy -= wy; dd[r] = d; sw += d; } else { fx += wx; fy += wy; }
                    mxx += 2.0 * wx * dx; mxy += 2.0 * wx * dy; myy += 2.0 * wy * dy;
                }
                for (int o = 0; o < nObst; ++o) {
                    const int r = mcv + (v * nObst + o) * Hp + k;
                    const double d = ddf(r);
                    const double w2 = 2.0 * d, dx = dbar[r * 2], dy = dbar[r * 2 + 1];
                    const double wx = w2 * dx, wy = w2 * dy;
                    fx -= wx; fy -= wy; dd[r] = d; sw += d;
                    mxx += 2.0 * wx * dx; mxy += 2.0 * wx * dy; myy += 2.0 * wy * dy;
                }
                frc[c * 2] = fx; frc[c * 2 + 1] = fy;
                Msm[c * 3] = mxx; Msm[c * 3 + 1] = mxy; Msm[c * 3 + 2] = myy;
            }
            // padding rows/columns of S: zero off-diagonal, unit diagonal (the factorisation keeps them so); the last
            // one carries the right-hand side
            for (int c = m.n1 + tid; c < m.n1p - 1; c += cta.nt) {
                for (int j = 0; j < c; ++j) S[scp_sidx(c, j)] = 0.0;
                S[scp_sidx(c, c)] = 1.0;
            }
            for (int j = tid; j < m.n1p; j += cta.nt)
                S[scp_sidx(m.n1p - 1, j)] = j < m.n1 ? (rhs_row ? rhs_row[j] : 0.0) : (j == m.n1p - 1 ? 1e300 : 0.0);
            CTA_RED_SUM(cta, red, 0, tid, sw)
        CTA_PHASE_END_RED(cta, red, 1)
        const double sd = cta_red_sum(cta, red, 0);
        wsum = sd;
        SCP_TIMER(13)
        if (alpha_slots > 0) {
            const int npair = nVeh * (nVeh - 1) >> 1;
            WARP_SECTION(w, nw)
                WARP_PHASE(lane)
                    // pair blocks, then the vehicle blocks, dealt round-robin to the warps
                    for (int p = w; p < npair + nVeh; p += nw) {
                        if (p < npair) {
                            int i = 0, q = p;
                            while (q >= nVeh - 1 - i) { q -= nVeh - 1 - i; ++i; }
                            pair_block_mma(lane, S, i, i + 1 + q, p * Hp, dd);
                        } else {
                            diag_block_mma(lane, S, p - npair, dg, rsub);
                        }
                    }
                WARP_PHASE_END
            WARP_SECTION_END
        }
        CTA_PHASE(tid)
            // omega row
            for (int c = tid; c < n; c += cta.nt) {
                const int v = c / Hp, a = c - v * Hp;
                const double *gv = g + (size_t)v * Hp * 2;
                double acc = 0.0;
                for (int k = a; k < Hp; ++k)
                    acc += gv[(k - a) * 2] * frc[(v * Hp + k) * 2] + gv[(k - a) * 2 + 1] * frc[(v * Hp + k) * 2 + 1];
                S[scp_sidx(n, c)] = -acc;
            }
            if (tid == 0) S[scp_sidx(n, n)] = sd + dg[n];
            // diagonal blocks entry by entry (long horizons only)
            const int per = Hp * (Hp + 1) >> 1;
            for (int e = tid; e < (alpha_slots == 0 ? nVeh * per : 0); e += cta.nt) {
                const int v = e / per;
                int a, b;
                scp_tri_decode(e - v * per, &a, &b);
                const double *gv = g + (size_t)v * Hp * 2;
                const double *Mv = Msm + (size_t)v * Hp * 3;
                double acc = 2.0 * scp_ldc(H + ((size_t)v * Hp + a) * Hp + b, coh) + (a == b ? dg[v * Hp + a] : 0.0);
                if (rsub && a == b + 1) acc += rsub[v * Hp + a];
                for (int k = a; k < Hp; ++k) {
                    const double gbx = gv[(k - b) * 2], gby = gv[(k - b) * 2 + 1];
                    const double bx = Mv[k * 3] * gbx + Mv[k * 3 + 1] * gby, by = Mv[k * 3 + 1] * gbx + Mv[k * 3 + 2] * gby;
                    acc += gv[(k - a) * 2] * bx + gv[(k - a) * 2 + 1] * by;
                }
                S[scp_sidx(v * Hp + a, v * Hp + b)] = acc;
            }
            // pair blocks entry by entry (horizons beyond the accumulator budget of pair_block_mma)
            if (alpha_slots == 0) {
                const int npair = nVeh * (nVeh - 1) >> 1;
                for (int e = tid; e < npair * Hp * Hp; e += cta.nt) {
                    const int p = e / (Hp * Hp), rem = e - p * Hp * Hp, b = rem / Hp, a = rem - b * Hp;
                    int i = 0, q = p;
                    while (q >= nVeh - 1 - i) { q -= nVeh - 1 - i; ++i; }
                    const int j = i + 1 + q, r0 = p * Hp;
                    const double *gi = g + (size_t)i * Hp * 2, *gj = g + (size_t)j * Hp * 2;
                    double acc = 0.0;
                    for (int k = (a > b ? a : b); k < Hp; ++k)
                        acc -= dd[r0 + k] * coef2(gi, r0 + k, k - a) * coef2(gj, r0 + k, k - b);
                    S[scp_sidx(j * Hp + b, i * Hp + a)] = acc;
                }
            }
        CTA_PHASE_END
        SCP_TIMER(14)
    }
};
