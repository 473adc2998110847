// ipm_core.cuh — CTA-resident primal-dual interior-point method for
//     min 1/2 x'Px + q'x   s.t.  A x <= b,   lb <= x <= ub                      (SCP_controller.py:135-141)
// one QP per CTA, everything the iteration touches resident in shared memory (or, for sizes that do not
// fit, in an L2-resident per-CTA slice of the workspace).
//
// Algorithm: Mehrotra predictor-corrector on the reduced normal equations
//     (P + A' D A + D_ub + D_lb) dx = rhs,      D = 1 / (s/z + delta)
// with the same starting point, step rule (0.99), centring exponent (3) and stopping rule as CVXOPT's coneqp
// (the solver BASELINE.json names), so that iterates are comparable with the CPU oracle iteration by
// iteration.  delta (params.qp_dual_reg) is a proximal regularisation of the dual block that keeps the
// normal matrix factorisable when active rows reach z/s ~ 1e15 (omega active); residuals are always the
// unregularised ones, so it perturbs the Newton direction, not the solution.
// Box rows never enter A: they only add to the diagonal of the normal matrix.  Bounds with
// |value| >= inf_bound (the omega <= 1e25 row of SCP_controller.py:127) are treated as absent.
//
// The normal matrix S lives as a tile-packed lower triangle (8x8 tiles) and is factorised by a blocked
// right-looking Cholesky: per tile column (a) factor + invert the diagonal tile, (b) panel = panel * Lkk^-T,
// (c) trailing tiles -= panel panel'.   Triangular solves walk the same tiles and use the stored inverses
// of the diagonal tiles.
//
// The constraint operator `Op` supplies the problem-specific pieces (structured pair rows for the fused SCP
// kernel, dense rows for the CVXOPT-replacement entry):
//     void Op::mul_P(cta, x, y)            y[0..n1) = P x
//     void Op::add_P(cta, S)               S(lower) += P
//     void Op::mul_A(cta, x, y)            y[0..mc) = A x
//     void Op::add_At(cta, w, v)           v[0..n1) += A' w
//     void Op::add_AtDA(cta, dd, S)        S(lower) += A' diag(dd) A
// each a sequence of complete phases (entered and left with the CTA synchronised).
#pragma once
#include "scp_common.cuh"

struct IpmCtl {
    double abstol, reltol, feastol, dual_reg, inf_bound;
    int max_iter;
};

// Pointers into the CTA's working set.  n1p = n1 rounded up to the tile size; vectors of length n1p have
// zero padding, the padded diagonal of S is 1.
struct IpmMem {
    int n1, n1p, T, mc;
    double *S;      // [T(T+1)/2 * 64]   tile-packed lower triangle of the normal matrix / its Cholesky factor
    double *Linv;   // [T * 64]          inverses of the diagonal tiles of the factor
    double *x, *q, *rx, *dx, *tn;                    // [n1p]
    double *bA, *sA, *zA, *rzA, *dsA, *dzA, *ccA;    // [mc]    collision rows
    double *ub, *sU, *zU, *dsU, *dzU, *ccU;          // [n1p]   x <= ub rows
    double *lb, *sL, *zL, *dsL, *dzL, *ccL;          // [n1p]   x >= lb rows
    double *red;    // [8 * SCP_MAX_WARPS] reduction scratch
    double *t8;     // [16] tile-solve scratch (8) + flags
};

struct IpmResult {
    double fval, gap, relgap, pres, dres;
    int iters, status;
};

// ------------------------------------------------------------------------------------------------ 8x8 tile leaf
// Factor the diagonal tile in place (lower triangle) and write the inverse of the factor to Linv (lower,
// upper part zero).  Executed by the first 8 threads of the CTA inside a phase (other threads idle).
// Returns 1 (in *fixed) if a pivot had to be repaired.
SCP_FN void tile_potrf_inv(int tid, double *Tkk, double *Linv, int *fixed)
{
#if SCP_DEVICE_BUILD
    // lanes 0..7 each own one row of the tile in registers; the other lanes of warp 0 shadow lane (tid & 7)
    // so that the full-mask shuffles stay convergent.
    if (tid < 32) {
        const int r = tid & 7;
        double row[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) row[c] = Tkk[r * 8 + c];
        double dinv[8];
        int bad = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double d = __shfl_sync(0xffffffffu, row[c], c);
            if (!(d > 1e-300)) { d = 1e300; bad = 1; }
            const double inv = rsqrt(d);
            dinv[c] = inv;
            if (r == c) row[c] = d * inv;
            else if (r > c) row[c] *= inv;
#pragma unroll
            for (int c2 = c + 1; c2 < 8; ++c2) {
                const double l = __shfl_sync(0xffffffffu, row[c], c2);   // L[c2][c]
                if (r >= c2) row[c2] -= row[c] * l;
            }
        }
        if (tid < 8) {
#pragma unroll
            for (int c = 0; c < 8; ++c) Tkk[r * 8 + c] = (c <= r) ? row[c] : 0.0;
        }
        __syncwarp();
        // inverse: lane j (< 8) solves L X[:,j] = e_j by forward substitution; L is read back from the tile
        if (tid < 8) {
            const int j = tid;
            double xcol[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < i && k >= j) acc -= Tkk[i * 8 + k] * xcol[k];
                xcol[i] = (i >= j) ? acc * dinv[i] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) Linv[i * 8 + j] = xcol[i];
        }
        if (tid == 0 && bad) *fixed = 1;
    }
#else
    if (tid == 0) {
        double dinv[8];
        for (int c = 0; c < 8; ++c) {
            double d = Tkk[c * 8 + c];
            if (!(d > 1e-300)) { d = 1e300; *fixed = 1; }
            const double inv = 1.0 / sqrt(d);
            dinv[c] = inv;
            Tkk[c * 8 + c] = d * inv;
            for (int r = c + 1; r < 8; ++r) Tkk[r * 8 + c] *= inv;
            for (int c2 = c + 1; c2 < 8; ++c2) {
                const double l = Tkk[c2 * 8 + c];
                for (int r = c2; r < 8; ++r) Tkk[r * 8 + c2] -= Tkk[r * 8 + c] * l;
            }
        }
        for (int r = 0; r < 8; ++r)
            for (int c = r + 1; c < 8; ++c) Tkk[r * 8 + c] = 0.0;
        for (int j = 0; j < 8; ++j) {
            double xcol[8];
            for (int i = 0; i < 8; ++i) {
                double acc = (i == j) ? 1.0 : 0.0;
                for (int k = j; k < i; ++k) acc -= Tkk[i * 8 + k] * xcol[k];
                xcol[i] = (i >= j) ? acc * dinv[i] : 0.0;
            }
            for (int i = 0; i < 8; ++i) Linv[i * 8 + j] = xcol[i];
        }
    }
#endif
}

// ------------------------------------------------------------------------------------------------ Cholesky
// In-place blocked Cholesky of the tile-packed lower triangle.  *fixed is set if any pivot was repaired.
SCP_FN void chol_tiles(Cta &cta, const IpmMem &m, int *fixed)
{
    const int T = m.T;
    double *S = m.S;
    for (int K = 0; K < T; ++K) {
        double *Skk = S + scp_tile_off(K, K);
        double *Lki = m.Linv + K * SCP_TILE2;
        CTA_PHASE(tid)
            tile_potrf_inv(tid, Skk, Lki, fixed);
        CTA_PHASE_END
        const int Tr = T - K - 1;
        if (Tr == 0) break;
        // (b) panel rows: L_IK[r][:] = S_IK[r][:] * Lkk^-T       one thread per panel row
        CTA_PHASE(tid)
            for (int pr = tid; pr < Tr * 8; pr += cta.nt) {
                const int I = K + 1 + (pr >> 3), r = pr & 7;
                double *row = S + scp_tile_off(I, K) + r * 8;
                double s[8], o[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) s[c] = row[c];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    double acc = 0.0;
#pragma unroll
                    for (int c2 = 0; c2 < 8; ++c2)
                        if (c2 <= c) acc += s[c2] * Lki[c * 8 + c2];
                    o[c] = acc;
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) row[c] = o[c];
            }
        CTA_PHASE_END
        // (c) trailing update: S_IJ -= L_IK L_JK'   (K < J <= I), a 4x4 quadrant per thread
        CTA_PHASE(tid)
            const int ntask = (Tr * (Tr + 1) >> 1) * 4;
            for (int t = tid; t < ntask; t += cta.nt) {
                int ii, jj;
                scp_tri_decode(t >> 2, &ii, &jj);
                const int I = K + 1 + ii, J = K + 1 + jj;
                const int r0 = (t & 2) ? 4 : 0, c0 = (t & 1) ? 4 : 0;
                const double *Ai = S + scp_tile_off(I, K) + r0 * 8;
                const double *Bj = S + scp_tile_off(J, K) + c0 * 8;
                double *C = S + scp_tile_off(I, J) + r0 * 8 + c0;
                double acc[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    double a[4], b[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) { a[r] = Ai[r * 8 + k]; b[r] = Bj[r * 8 + k]; }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[r][c] += a[r] * b[c];
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) C[r * 8 + c] -= acc[r][c];
            }
        CTA_PHASE_END
    }
}

// v := S^-1 v  using the factor in m.S / m.Linv (v has length n1p)
SCP_FN void chol_solve_tiles(Cta &cta, const IpmMem &m, double *v)
{
    const int T = m.T;
    const double *S = m.S;
    double *t8 = m.t8;
    for (int K = 0; K < T; ++K) {                       // forward: L y = v
        const double *Lki = m.Linv + K * SCP_TILE2;
        CTA_PHASE(tid)
            if (tid < 8) {
                double acc = 0.0;
                for (int c = 0; c <= tid; ++c) acc += Lki[tid * 8 + c] * v[K * 8 + c];
                t8[tid] = acc;
            }
        CTA_PHASE_END
        CTA_PHASE(tid)
            if (tid < 8) v[K * 8 + tid] = t8[tid];
            for (int pr = tid; pr < (T - K - 1) * 8; pr += cta.nt) {
                const int I = K + 1 + (pr >> 3), r = pr & 7;
                const double *row = S + scp_tile_off(I, K) + r * 8;
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < 8; ++c) acc += row[c] * t8[c];
                v[I * 8 + r] -= acc;
            }
        CTA_PHASE_END
    }
    for (int K = T - 1; K >= 0; --K) {                  // backward: L' x = y
        const double *Lki = m.Linv + K * SCP_TILE2;
        CTA_PHASE(tid)
            if (tid < 8) {
                double acc = 0.0;
                for (int r = tid; r < 8; ++r) acc += Lki[r * 8 + tid] * v[K * 8 + r];
                t8[tid] = acc;
            }
        CTA_PHASE_END
        CTA_PHASE(tid)
            if (tid < 8) v[K * 8 + tid] = t8[tid];
            for (int j = tid; j < K * 8; j += cta.nt) {
                const double *col = S + scp_tile_off(K, j >> 3) + (j & 7);
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < 8; ++r) acc += col[r * 8] * t8[r];
                v[j] -= acc;
            }
        CTA_PHASE_END
    }
}

// ------------------------------------------------------------------------------------------------ helpers
SCP_FN bool ipm_has_ub(const IpmMem &m, const IpmCtl &ctl, int c) { return c < m.n1 && fabs(m.ub[c]) < ctl.inf_bound; }
SCP_FN bool ipm_has_lb(const IpmMem &m, const IpmCtl &ctl, int c) { return c < m.n1 && fabs(m.lb[c]) < ctl.inf_bound; }

// S := 0 with unit diagonal on the padding; then the caller adds P, A'DA and the box diagonal
SCP_FN void ipm_clear_S(Cta &cta, const IpmMem &m)
{
    CTA_PHASE(tid)
        const int tot = (m.T * (m.T + 1) >> 1) * SCP_TILE2;
        for (int e = tid; e < tot; e += cta.nt) m.S[e] = 0.0;
    CTA_PHASE_END
    CTA_PHASE(tid)
        for (int c = m.n1 + tid; c < m.n1p; c += cta.nt) m.S[scp_sidx(c, c)] = 1.0;
    CTA_PHASE_END
}

// ------------------------------------------------------------------------------------------------ the solver
// On entry m.q, m.bA, m.ub, m.lb hold the problem data (padding of q/ub/lb beyond n1 is ignored).
// On exit m.x holds the solution.
template <class Op>
SCP_FN void ipm_solve(Cta &cta, Op &op, const IpmMem &m, const IpmCtl &ctl, IpmResult *res)
{
    const int n1 = m.n1, n1p = m.n1p, mc = m.mc;
    double *red = m.red;
    int *fixed_p = (int *)(m.t8 + 8);   // pivot-repair flag lives in shared scratch (t8 has 16 slots, 8 used)

    // ---- constants: row count, residual scales, zero the padding ---------------------------------
    CTA_RED_BEGIN(cta, 3)
    CTA_PHASE(tid)
        double cnt = 0.0, hq = 0.0, hh = 0.0;
        if (tid == 0) *fixed_p = 0;
        for (int c = tid; c < n1p; c += cta.nt) {
            if (c >= n1) { m.q[c] = 0.0; m.x[c] = 0.0; m.dx[c] = 0.0; m.rx[c] = 0.0; m.tn[c] = 0.0; }
            else hq += m.q[c] * m.q[c];
            if (ipm_has_ub(m, ctl, c)) { cnt += 1.0; hh += m.ub[c] * m.ub[c]; }
            if (ipm_has_lb(m, ctl, c)) { cnt += 1.0; hh += m.lb[c] * m.lb[c]; }
        }
        for (int r = tid; r < mc; r += cta.nt) hh += m.bA[r] * m.bA[r];
        CTA_RED_SUM(cta, red, 0, tid, cnt)
        CTA_RED_SUM(cta, red, 1, tid, hq)
        CTA_RED_SUM(cta, red, 2, tid, hh)
    CTA_PHASE_END_RED(cta, red, 3)
    const double mrows = cta_red_sum(cta, red, 0) + (double)mc;
    const double resx0 = fmax(1.0, sqrt(cta_red_sum(cta, red, 1)));
    const double resz0 = fmax(1.0, sqrt(cta_red_sum(cta, red, 2)));

    // ---- starting point (coneqp): (P + G'G) x = G'h - q ; z = Gx - h ; s = -z ; shift ------------
    ipm_clear_S(cta, m);
    op.add_P(cta, m.S);
    CTA_PHASE(tid)
        for (int r = tid; r < mc; r += cta.nt) m.dsA[r] = 1.0;            // dd = 1
        for (int c = tid; c < n1p; c += cta.nt) {
            double d = 0.0, rhs = 0.0;
            if (ipm_has_ub(m, ctl, c)) { d += 1.0; rhs += m.ub[c]; }
            if (ipm_has_lb(m, ctl, c)) { d += 1.0; rhs += m.lb[c]; }
            if (c < n1) { m.S[scp_sidx(c, c)] += d; m.x[c] = rhs - m.q[c]; }
        }
    CTA_PHASE_END
    op.add_AtDA(cta, m.dsA, m.S);
    op.add_At(cta, m.bA, m.x);
    chol_tiles(cta, m, fixed_p);
    chol_solve_tiles(cta, m, m.x);
    op.mul_A(cta, m.x, m.rzA);                                              // rzA = A x (scratch)
    CTA_RED_BEGIN(cta, 3)
    CTA_PHASE(tid)
        double nrm = 0.0, ts = -1e300, tz = -1e300;
        for (int r = tid; r < mc; r += cta.nt) {
            const double z = m.rzA[r] - m.bA[r];
            m.zA[r] = z; m.sA[r] = -z;
            nrm += z * z; ts = fmax(ts, z); tz = fmax(tz, -z);
        }
        for (int c = tid; c < n1p; c += cta.nt) {
            if (ipm_has_ub(m, ctl, c)) {
                const double z = m.x[c] - m.ub[c];
                m.zU[c] = z; m.sU[c] = -z; nrm += z * z; ts = fmax(ts, z); tz = fmax(tz, -z);
            } else { m.zU[c] = 0.0; m.sU[c] = 1.0; }
            if (ipm_has_lb(m, ctl, c)) {
                const double z = m.lb[c] - m.x[c];
                m.zL[c] = z; m.sL[c] = -z; nrm += z * z; ts = fmax(ts, z); tz = fmax(tz, -z);
            } else { m.zL[c] = 0.0; m.sL[c] = 1.0; }
        }
        CTA_RED_SUM(cta, red, 0, tid, nrm)
        CTA_RED_MAX(cta, red, 1, tid, ts)
        CTA_RED_MAX(cta, red, 2, tid, tz)
    CTA_PHASE_END_RED(cta, red, 3)
    {
        const double nrm = sqrt(cta_red_sum(cta, red, 0));
        const double ts = cta_red_max(cta, red, 1), tz = cta_red_max(cta, red, 2);
        const double thr = -1e-8 * fmax(nrm, 1.0);
        const double as = (ts >= thr) ? 1.0 + ts : 0.0, az = (tz >= thr) ? 1.0 + tz : 0.0;
        CTA_PHASE(tid)
            for (int r = tid; r < mc; r += cta.nt) { m.sA[r] += as; m.zA[r] += az; }
            for (int c = tid; c < n1p; c += cta.nt) {
                if (ipm_has_ub(m, ctl, c)) { m.sU[c] += as; m.zU[c] += az; }
                if (ipm_has_lb(m, ctl, c)) { m.sL[c] += as; m.zL[c] += az; }
            }
        CTA_PHASE_END
    }

    int iters = 0, status = SCPB200_ST_QP_MAXITER;
    double f0 = 0.0, gap = 0.0, relgap = -1.0, pres = 0.0, dres = 0.0;
    for (iters = 0; iters <= ctl.max_iter; ++iters) {
        // ---- residuals: rx = Px + q + G'z ; rz = s + Gx - h ; costs --------------------------------
        op.mul_P(cta, m.x, m.tn);                                            // tn = P x
        op.mul_A(cta, m.x, m.rzA);                                           // rzA = A x
        CTA_RED_BEGIN(cta, 4)
        CTA_PHASE(tid)
            double pf = 0.0, pg = 0.0, prz = 0.0, pzr = 0.0;
            for (int c = tid; c < n1p; c += cta.nt) {
                double r = 0.0;
                if (c < n1) {
                    pf += m.x[c] * (0.5 * m.tn[c] + m.q[c]);
                    r = m.tn[c] + m.q[c];
                }
                if (ipm_has_ub(m, ctl, c)) {
                    const double rz = m.sU[c] + m.x[c] - m.ub[c];
                    r += m.zU[c]; pg += m.sU[c] * m.zU[c]; prz += rz * rz; pzr += m.zU[c] * rz;
                }
                if (ipm_has_lb(m, ctl, c)) {
                    const double rz = m.sL[c] - m.x[c] + m.lb[c];
                    r -= m.zL[c]; pg += m.sL[c] * m.zL[c]; prz += rz * rz; pzr += m.zL[c] * rz;
                }
                m.rx[c] = r;
            }
            for (int r = tid; r < mc; r += cta.nt) {
                const double rz = m.sA[r] + m.rzA[r] - m.bA[r];
                m.rzA[r] = rz;
                pg += m.sA[r] * m.zA[r]; prz += rz * rz; pzr += m.zA[r] * rz;
            }
            CTA_RED_SUM(cta, red, 0, tid, pf)
            CTA_RED_SUM(cta, red, 1, tid, pg)
            CTA_RED_SUM(cta, red, 2, tid, prz)
            CTA_RED_SUM(cta, red, 3, tid, pzr)
        CTA_PHASE_END_RED(cta, red, 4)
        f0 = cta_red_sum(cta, red, 0);
        gap = cta_red_sum(cta, red, 1);
        const double resz = sqrt(cta_red_sum(cta, red, 2));
        const double zrz = cta_red_sum(cta, red, 3);
        op.add_At(cta, m.zA, m.rx);                                          // rx += A' zA
        CTA_RED_BEGIN(cta, 1)
        CTA_PHASE(tid)
            double p = 0.0;
            for (int c = tid; c < n1; c += cta.nt) p += m.rx[c] * m.rx[c];
            CTA_RED_SUM(cta, red, 0, tid, p)
        CTA_PHASE_END_RED(cta, red, 1)
        const double resx = sqrt(cta_red_sum(cta, red, 0));
        const double dcost = f0 + zrz - gap;
        if (f0 < 0.0) relgap = gap / -f0;
        else if (dcost > 0.0) relgap = gap / dcost;
        else relgap = -1.0;
        pres = resz / resz0;
        dres = resx / resx0;
        if (pres <= ctl.feastol && dres <= ctl.feastol &&
            (gap <= ctl.abstol || (relgap >= 0.0 && relgap <= ctl.reltol))) { status = 0; break; }
        if (iters == ctl.max_iter) break;

        // ---- normal matrix with D = 1/(s/z + delta) and its factor ---------------------------------
        ipm_clear_S(cta, m);
        op.add_P(cta, m.S);
        CTA_PHASE(tid)
            for (int r = tid; r < mc; r += cta.nt) m.dzA[r] = 1.0 / (m.sA[r] / m.zA[r] + ctl.dual_reg);   // dd in dzA
            for (int c = tid; c < n1; c += cta.nt) {
                double d = 0.0;
                if (ipm_has_ub(m, ctl, c)) d += 1.0 / (m.sU[c] / m.zU[c] + ctl.dual_reg);
                if (ipm_has_lb(m, ctl, c)) d += 1.0 / (m.sL[c] / m.zL[c] + ctl.dual_reg);
                m.S[scp_sidx(c, c)] += d;
            }
        CTA_PHASE_END
        op.add_AtDA(cta, m.dzA, m.S);
        chol_tiles(cta, m, fixed_p);

        const double mu = gap / mrows;
        double sigma = 0.0, step = 1.0;
        for (int pass = 0; pass < 2; ++pass) {
            // w1 = D (rz + bs/z), bs = -s z + sigma mu - [pass 1] (ds_a dz_a);  rhs = -rx - G'w1
            CTA_PHASE(tid)
                for (int r = tid; r < mc; r += cta.nt) {
                    const double s = m.sA[r], z = m.zA[r];
                    double bs = -s * z + sigma * mu;
                    if (pass == 1) bs -= m.ccA[r];
                    m.dzA[r] = (m.rzA[r] + bs / z) / (s / z + ctl.dual_reg);          // w1 in dzA
                }
                for (int c = tid; c < n1p; c += cta.nt) {
                    double rhs = -m.rx[c];
                    if (ipm_has_ub(m, ctl, c)) {
                        const double s = m.sU[c], z = m.zU[c];
                        double bs = -s * z + sigma * mu;
                        if (pass == 1) bs -= m.ccU[c];
                        const double w1 = (s + m.x[c] - m.ub[c] + bs / z) / (s / z + ctl.dual_reg);
                        m.dzU[c] = w1; rhs -= w1;
                    }
                    if (ipm_has_lb(m, ctl, c)) {
                        const double s = m.sL[c], z = m.zL[c];
                        double bs = -s * z + sigma * mu;
                        if (pass == 1) bs -= m.ccL[c];
                        const double w1 = (s - m.x[c] + m.lb[c] + bs / z) / (s / z + ctl.dual_reg);
                        m.dzL[c] = w1; rhs += w1;
                    }
                    m.dx[c] = (c < n1) ? rhs : 0.0;
                    m.tn[c] = 0.0;
                }
            CTA_PHASE_END
            op.add_At(cta, m.dzA, m.tn);                                     // tn = A' w1A
            CTA_PHASE(tid)
                for (int c = tid; c < n1; c += cta.nt) m.dx[c] -= m.tn[c];
            CTA_PHASE_END
            chol_solve_tiles(cta, m, m.dx);
            op.mul_A(cta, m.dx, m.dsA);                                      // dsA = A dx (scratch)
            CTA_RED_BEGIN(cta, 3)
            CTA_PHASE(tid)
                double pdd = 0.0, ts = 0.0, tz = 0.0;
                for (int r = tid; r < mc; r += cta.nt) {
                    const double s = m.sA[r], z = m.zA[r];
                    double bs = -s * z + sigma * mu;
                    if (pass == 1) bs -= m.ccA[r];
                    const double dz = m.dzA[r] + m.dsA[r] / (s / z + ctl.dual_reg);
                    const double ds = (bs - s * dz) / z;
                    m.dzA[r] = dz; m.dsA[r] = ds;
                    pdd += ds * dz; ts = fmax(ts, -ds / s); tz = fmax(tz, -dz / z);
                    if (pass == 0) m.ccA[r] = ds * dz;
                }
                for (int c = tid; c < n1p; c += cta.nt) {
                    if (ipm_has_ub(m, ctl, c)) {
                        const double s = m.sU[c], z = m.zU[c];
                        double bs = -s * z + sigma * mu;
                        if (pass == 1) bs -= m.ccU[c];
                        const double dz = m.dzU[c] + m.dx[c] / (s / z + ctl.dual_reg);
                        const double ds = (bs - s * dz) / z;
                        m.dzU[c] = dz; m.dsU[c] = ds;
                        pdd += ds * dz; ts = fmax(ts, -ds / s); tz = fmax(tz, -dz / z);
                        if (pass == 0) m.ccU[c] = ds * dz;
                    }
                    if (ipm_has_lb(m, ctl, c)) {
                        const double s = m.sL[c], z = m.zL[c];
                        double bs = -s * z + sigma * mu;
                        if (pass == 1) bs -= m.ccL[c];
                        const double dz = m.dzL[c] - m.dx[c] / (s / z + ctl.dual_reg);
                        const double ds = (bs - s * dz) / z;
                        m.dzL[c] = dz; m.dsL[c] = ds;
                        pdd += ds * dz; ts = fmax(ts, -ds / s); tz = fmax(tz, -dz / z);
                        if (pass == 0) m.ccL[c] = ds * dz;
                    }
                }
                CTA_RED_SUM(cta, red, 0, tid, pdd)
                CTA_RED_MAX(cta, red, 1, tid, ts)
                CTA_RED_MAX(cta, red, 2, tid, tz)
            CTA_PHASE_END_RED(cta, red, 3)
            const double dsdz = cta_red_sum(cta, red, 0);
            const double t = fmax(0.0, fmax(cta_red_max(cta, red, 1), cta_red_max(cta, red, 2)));
            if (t == 0.0) step = 1.0;
            else step = fmin(1.0, (pass == 0 ? 1.0 : 0.99) / t);
            if (pass == 0) {
                const double base = fmin(1.0, fmax(0.0, 1.0 - step + dsdz / gap * step * step));
                sigma = base * base * base;
            }
        }
        CTA_PHASE(tid)
            for (int c = tid; c < n1p; c += cta.nt) {
                if (c < n1) m.x[c] += step * m.dx[c];
                if (ipm_has_ub(m, ctl, c)) { m.sU[c] += step * m.dsU[c]; m.zU[c] += step * m.dzU[c]; }
                if (ipm_has_lb(m, ctl, c)) { m.sL[c] += step * m.dsL[c]; m.zL[c] += step * m.dzL[c]; }
            }
            for (int r = tid; r < mc; r += cta.nt) { m.sA[r] += step * m.dsA[r]; m.zA[r] += step * m.dzA[r]; }
        CTA_PHASE_END
    }
    if (*fixed_p) status |= SCPB200_ST_QP_PIVOT;
    res->fval = f0; res->gap = gap; res->relgap = relgap; res->pres = pres; res->dres = dres;
    res->iters = iters; res->status = status;
}
