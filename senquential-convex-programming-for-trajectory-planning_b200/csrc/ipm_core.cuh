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
// The normal matrix S lives as a tile-packed lower triangle (8x8 tiles, layout scp_tphys) and is factorised by a
// blocked right-looking Cholesky (chol_factor); the diagonal tiles of the factor are then inverted in place
// (chol_invert_diag) and every solve is a tile-wise substitution inside one warp (chol_solve).
//
// The constraint operator `Op` supplies the problem-specific pieces (structured pair rows for the fused SCP
// kernel, dense rows for the CVXOPT-replacement entry).  Products with A and A' are split into a CTA-wide
// preparation phase and an inline per-row / per-column evaluation, so that the solver can fuse them into its own
// row and column loops instead of paying a phase (barrier + loop overhead) per product:
//     void   Op::prep(cta, x, w)                phases; afterwards row_dot refers to x, col_dot to w (either may be null)
//     double Op::row_dot(r)                     (A x)[r]
//     double Op::col_dot(c)                     (A' w)[c]
//     double Op::P_col(c, x)                    (P x)[c]
//     void   Op::form_normal(cta, m, dd, dg, ddf, dgf)   S(lower) = P + A' diag(dd) A + diag(dg), with dd[r] = ddf(r) and
//                                               dg[c] = dgf(c) (box terms) evaluated and stored by the operator
// form_normal must leave the padding of S (rows >= n1) as unit diagonal / zero off-diagonal.
#pragma once
#include "scp_common.cuh"

struct IpmCtl {
    double abstol, reltol, feastol, dual_reg, inf_bound;
    double dres_floor;   // accept at the precision floor of the dual residual when it is <= dres_floor (0 = never)
    int max_iter;
    // Warm start across the QPs of one SCP loop (consecutive QPs differ only in the linearisation point):
    //   snap      global memory for one interior iterate (x, s, z of every row; ipm_snap_doubles()), or null
    //   warm      start from the iterate in `snap` instead of the coneqp starting point
    //   snap_relgap  the iterate saved is the first one with relative gap <= snap_relgap: centred, still far enough
    //             from the boundary for the next QP's active set to change (a converged iterate would jam)
    double *snap;
    double snap_relgap;
    int warm;
    int snap_min_iter;   // tuning: the saved iterate is not taken before this iteration
};

// Pointers into the CTA's working set.  n1p = n1 rounded up to the tile size; vectors of length n1p have
// zero padding, the padded diagonal of S is 1.
struct IpmMem {
    int n1, n1p, T, mc;
    bool S_far;     // S lives in the L2-resident workspace slice, not in shared memory (long horizons): chol_factor_left
    double *S;      // [T(T+1)/2 * 64]   tile-packed lower triangle of the normal matrix / its Cholesky factor;
                    //                   (diagonal tiles replaced by their inverses after chol_invert_diag)
    double *x, *q, *rx, *dx, *tn;                    // [n1p]
    double *bA, *sA, *zA, *rzA, *dsA, *dzA, *ccA, *eA;   // [mc]    collision rows (e = 1/(s + delta z))
    double *ub, *sU, *zU, *dsU, *dzU, *ccU, *eU;         // [n1p]   x <= ub rows
    double *lb, *sL, *zL, *dsL, *dzL, *ccL, *eL;         // [n1p]   x >= lb rows
    double *dinv;   // [n1p]             reciprocal pivots of the factor
    // difference rows (steering-rate bounds, params.enable_rate_rows): for c < nr, with the difference operator
    //     (D x)[c] = x[c] - (c % rper ? x[c-1] : 0)      (rper = Hp: the chain restarts at every vehicle)
    // the two rows  (D x)[c] <= hP[c]  and  -(D x)[c] <= hM[c].  Like the box rows they never enter A: they add a
    // tridiagonal term to the vehicle blocks of the normal matrix (diagonal through dgf, sub-diagonal through rsub).
    // nr = 0: none (a compile-time zero in the fixed-shape kernels, so every loop below folds away).
    int nr, rper;
    double *hP, *sP, *zP, *dsP, *dzP, *ccP, *eP;         // [nr]
    double *hM, *sM, *zM, *dsM, *dzM, *ccM, *eM;         // [nr]
    double *rsub;   // [nr]              S[c][c-1] += rsub[c]  (0 where the chain restarts)
    double *red;    // [SCP_RED_DOUBLES] reduction scratch (double-buffered, see scp_common.cuh)
    double *t8;     // [16] tile-solve scratch (8) + flags
};

struct IpmResult {
    double fval, gap, relgap, pres, dres;
    int iters, status;
    int snap_saved;      // an iterate was written to ctl.snap during this solve
};

// Padded order of the normal matrix: a multiple of the tile size with AT LEAST ONE padding row.  The last row
// (n1p - 1) carries a right-hand side through the factorisation (see chol_factor).
SCP_HDFN int ipm_padded(int n1) { return scp_round_up(n1 + 1, SCP_TILE); }

SCP_HDFN size_t ipm_snap_doubles(int n1p, int mc, int nr = 0)
{
    return (size_t)5 * n1p + 2 * (size_t)((mc + 1) & ~1) + 4 * (size_t)((nr + 1) & ~1);
}

// difference rows: (D x)[c], and column c of D' w for w = wP - wM
SCP_FN bool ipm_rfirst(const IpmMem &m, int c) { return c % m.rper == 0; }
SCP_FN double ipm_rdiff(const IpmMem &m, const double *x, int c) { return x[c] - (ipm_rfirst(m, c) ? 0.0 : x[c - 1]); }
SCP_FN double ipm_rcol(const IpmMem &m, const double *wP, const double *wM, int c)
{
    double a = wP[c] - wM[c];
    if (c + 1 < m.nr && !ipm_rfirst(m, c + 1)) a -= wP[c + 1] - wM[c + 1];
    return a;
}

// Copy the interior iterate (x, sA, zA, sU, zU, sL, zL) to / from the snapshot area (one phase).
SCP_FN void ipm_snapshot(Cta &cta, const IpmMem &m, double *snap, bool save)
{
    const int n1p = m.n1p, mc = m.mc;
    const size_t rows = (size_t)((mc + 1) & ~1);
    double *px = snap, *psA = px + n1p, *pzA = psA + rows, *psU = pzA + rows, *pzU = psU + n1p, *psL = pzU + n1p,
           *pzL = psL + n1p;
    const size_t rr = (size_t)((m.nr + 1) & ~1);
    double *psP = pzL + n1p, *pzP = psP + rr, *psM = pzP + rr, *pzM = psM + rr;
    CTA_PHASE(tid)
        if (save) {
            for (int c = tid; c < m.nr; c += cta.nt) { psP[c] = m.sP[c]; pzP[c] = m.zP[c]; psM[c] = m.sM[c]; pzM[c] = m.zM[c]; }
            for (int c = tid; c < n1p; c += cta.nt) {
                px[c] = m.x[c]; psU[c] = m.sU[c]; pzU[c] = m.zU[c]; psL[c] = m.sL[c]; pzL[c] = m.zL[c];
            }
            for (int r = tid; r < mc; r += cta.nt) { psA[r] = m.sA[r]; pzA[r] = m.zA[r]; }
        } else {
            for (int c = tid; c < m.nr; c += cta.nt) {
                m.sP[c] = SCP_LD_COHERENT(psP + c); m.zP[c] = SCP_LD_COHERENT(pzP + c);
                m.sM[c] = SCP_LD_COHERENT(psM + c); m.zM[c] = SCP_LD_COHERENT(pzM + c);
            }
            for (int c = tid; c < n1p; c += cta.nt) {
                m.x[c] = SCP_LD_COHERENT(px + c);
                m.sU[c] = SCP_LD_COHERENT(psU + c); m.zU[c] = SCP_LD_COHERENT(pzU + c);
                m.sL[c] = SCP_LD_COHERENT(psL + c); m.zL[c] = SCP_LD_COHERENT(pzL + c);
            }
            for (int r = tid; r < mc; r += cta.nt) { m.sA[r] = SCP_LD_COHERENT(psA + r); m.zA[r] = SCP_LD_COHERENT(pzA + r); }
        }
    CTA_PHASE_END
}

// ------------------------------------------------------------------------------------------------ 8x8 tile leaves
// All 8x8 tiles (normal matrix, scratch) use the half-row-swapped layout of scp_tphys().
//
// Cholesky factor of one diagonal tile, in place (lower triangle only), reciprocal
// pivots to dinv[0..8).  The factorisation is ONE dependent chain of 8 pivots (rsqrt 63 + mul 9 + fma 9 cycles
// each on B200), so a single lane runs it out of registers with every other update off the chain.
#define SCP_TRI(r, c) ((r) * ((r) + 1) / 2 + (c))
SCP_FN void tile_potrf(double *Tkk, double *dinv, int *fixed)
{
    double a[36];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[SCP_TRI(r, c)] = Tkk[scp_tphys(r, c)];
    int bad = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double d = a[SCP_TRI(c, c)];
        if (!(d > 1e-300)) { d = 1e300; bad = 1; }
#if SCP_DEVICE_BUILD
        const double inv = rsqrt(d);
#else
        const double inv = 1.0 / sqrt(d);
#endif
        dinv[c] = inv;
        a[SCP_TRI(c, c)] = d * inv;
#pragma unroll
        for (int r = c + 1; r < 8; ++r) a[SCP_TRI(r, c)] *= inv;
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2)
#pragma unroll
            for (int r = c2; r < 8; ++r) a[SCP_TRI(r, c2)] -= a[SCP_TRI(r, c)] * a[SCP_TRI(c2, c)];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) Tkk[scp_tphys(r, c)] = a[SCP_TRI(r, c)];     // the upper triangle is never read
    if (bad) *fixed = 1;
}

// Column j of the inverse of a lower-triangular 8x8 tile L (reciprocal diagonal in dinv), by forward
// substitution:  X[i][j] = (delta_ij - sum_{k=j}^{i-1} L[i][k] X[k][j]) dinv[i].  The caller stores xcol.
SCP_FN void tile_trtri_column(const double *L, const double *dinv, int j, double xcol[8])
{
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < i && k >= j) acc -= L[scp_tphys(i, k)] * xcol[k];
        xcol[i] = (i >= j) ? acc * dinv[i] : 0.0;
    }
}

// ---- warp-level tile products on the FP64 tensor path (DMMA m8n8k4; 37 TFLOP/s measured on B200, the same
// peak as the DFMA pipe, but two instructions and four 8-byte loads per lane replace ~300 scalar instructions
// per 8x8x8 product).  Fragment ownership for lane l: A[l>>2][(l&3) + 4h], B[(l&3) + 4h][l>>2],
// C[l>>2][2(l&3) .. 2(l&3)+1].  The host build (kernel-logic emulator) runs plain loops on lane 0.
#if SCP_DEVICE_BUILD
SCP_FN void scp_dmma(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
SCP_FN int scp_frag_rowmajor(int lane, int h)     // element (row = lane>>2, col = (lane&3) + 4h)
{
    const int r = lane >> 2;
    return (r << 3) + (((h ^ (r >> 1)) & 1) << 2) + (lane & 3);
}
SCP_FN int scp_frag_colmajor(int lane, int h)     // element (row = (lane&3) + 4h, col = lane>>2)
{
    return scp_tphys((lane & 3) + 4 * h, lane >> 2);
}
SCP_FN int scp_frag_c(int lane)                   // elements (row = lane>>2, cols 2(lane&3), 2(lane&3)+1), 16-byte aligned
{
    const int r = lane >> 2, cp = lane & 3;
    return (r << 3) + ((((cp >> 1) ^ (r >> 1)) & 1) << 2) + ((cp & 1) << 1);
}
#endif

// C -= A B'   (A, B, C tiles)
SCP_FN void warp_tile_syrk(int lane, double *C, const double *A, const double *B)
{
#if SCP_DEVICE_BUILD
    double2 *cp = reinterpret_cast<double2 *>(C + scp_frag_c(lane));
    const double a0 = A[scp_frag_rowmajor(lane, 0)], a1 = A[scp_frag_rowmajor(lane, 1)];
    const double b0 = B[scp_frag_rowmajor(lane, 0)], b1 = B[scp_frag_rowmajor(lane, 1)];
    double2 c = *cp;
    double n0 = 0.0, n1 = 0.0;
    scp_dmma(n0, n1, a0, b0);
    scp_dmma(n0, n1, a1, b1);
    c.x -= n0; c.y -= n1;
    *cp = c;
#else
    if (lane == 0)
        for (int r = 0; r < 8; ++r)
            for (int c = 0; c < 8; ++c) {
                double acc = 0.0;
                for (int k = 0; k < 8; ++k) acc += A[scp_tphys(r, k)] * B[scp_tphys(c, k)];
                C[scp_tphys(r, c)] -= acc;
            }
#endif
}

// Two independent tiles at once: the loads of both are in flight before the first product issues (a single tile is a
// dependent chain load -> DMMA -> DMMA -> store of ~250 cycles).
SCP_FN void warp_tile_syrk2(int lane, double *C, const double *A, const double *B, double *C2, const double *A2, const double *B2)
{
#if SCP_DEVICE_BUILD
    const int ia0 = scp_frag_rowmajor(lane, 0), ia1 = scp_frag_rowmajor(lane, 1), ic = scp_frag_c(lane);
    double2 *cp = reinterpret_cast<double2 *>(C + ic), *cp2 = reinterpret_cast<double2 *>(C2 + ic);
    const double a0 = A[ia0], a1 = A[ia1], b0 = B[ia0], b1 = B[ia1];
    const double e0 = A2[ia0], e1 = A2[ia1], f0 = B2[ia0], f1 = B2[ia1];
    double2 c = *cp, c2 = *cp2;
    double n0 = 0.0, n1 = 0.0, m0 = 0.0, m1 = 0.0;
    scp_dmma(n0, n1, a0, b0);
    scp_dmma(m0, m1, e0, f0);
    scp_dmma(n0, n1, a1, b1);
    scp_dmma(m0, m1, e1, f1);
    c.x -= n0; c.y -= n1; c2.x -= m0; c2.y -= m1;
    *cp = c;
    *cp2 = c2;
#else
    warp_tile_syrk(lane, C, A, B);
    warp_tile_syrk(lane, C2, A2, B2);
#endif
}

// ------------------------------------------------------------------------------------------------ Cholesky
// (Round 2, long horizons with the matrix in the L2-resident workspace: FOUR trailing tiles in flight per warp instead of
// two made the update 3 x slower, 3.3 M against 1.06 M cycles per factorisation at Hp = 50.  The update is not bound by
// the latency of a round trip: it moves 1.8 KB per tile product through the SM's 64 B / cycle path to L2, >= 620 k cycles
// per factorisation, and the paired order keeps the panel tiles in L1 where the strided order of four does not.  What
// helps there is less traffic — accumulators in registers over the tile columns, the shared operand row staged in shared
// memory — not more loads in flight.  profiles/r02_hp50_four_tiles_in_flight_and_prefetch_experiment_phase_timers.txt)
// In-place blocked right-looking Cholesky of the tile-packed lower triangle: on exit m.S holds L, m.dinv the
// reciprocal pivots.
//   per tile column K:  (a) lane 0 factors the diagonal tile (in the shadow of the previous trailing update),
//                       (b) panel rows are solved against it by substitution (one thread per row, right-looking
//                           inside the row: 18 cycles per unknown on the dependent chain),
//                       (c) trailing tiles -= panel panel' (one warp per pair of tiles, DMMA).
// The factor is NOT inverted (the previous version spent n1^3/6 multiply-adds and 2 T barriers per factorisation on
// X = L^-1; see chol_invert_diag / chol_solve for what replaced it), and the forward substitution of ONE right-hand side
// is free — stored as the last row of the matrix (row n1p - 1, a padding row: callers put rhs' there and a huge
// diagonal), it is solved by the panel phases like any other row and comes out as y' = (L^-1 rhs)'.
// *fixed is set if any pivot had to be repaired.
SCP_FN void chol_factor(Cta &cta, const IpmMem &m, int *fixed SCP_TIMER_ARG)
{
    const int T = m.T;
    double *S = m.S, *dinv = m.dinv;
    CTA_PHASE(tid)
        if (tid == 0) tile_potrf(S + scp_tile_off(0, 0), dinv, fixed);
    CTA_PHASE_END
    SCP_TIMER(2)
    for (int K = 0; K < T - 1; ++K) {
        const double *Lkk = S + scp_tile_off(K, K);
        const int Tr = T - K - 1;
        // (b) panel rows: x L_KK' = s
        CTA_PHASE(tid)
            for (int pr = tid; pr < Tr * 8; pr += cta.nt) {
                const int I = K + 1 + (pr >> 3), r = pr & 7;
                double *row = S + scp_tile_off(I, K) + (r << 3);
                const int h0 = ((r >> 1) & 1) << 2, h1 = h0 ^ 4;       // where logical columns 0-3 / 4-7 live
                double x[8];
#pragma unroll
                for (int c = 0; c < 4; ++c) { x[c] = row[h0 + c]; x[c + 4] = row[h1 + c]; }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    x[c] *= dinv[K * 8 + c];
#pragma unroll
                    for (int c2 = c + 1; c2 < 8; ++c2) x[c2] -= x[c] * Lkk[scp_tphys(c2, c)];
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) { row[h0 + c] = x[c]; row[h1 + c] = x[c + 4]; }
            }
        CTA_PHASE_END
        SCP_TIMER(3)
        // (c) trailing update: S_IJ -= L_IK L_JK'   (K < J <= I), one warp per output tile; tile index 0 is the next
        // diagonal tile: warp 0 takes it first and factors it while the others work through the rest
        WARP_SECTION(w, nw)
            const int ntile = Tr * (Tr + 1) >> 1;
            WARP_PHASE(lane)
                if (w == 0) warp_tile_syrk(lane, S + scp_tile_off(K + 1, K + 1), S + scp_tile_off(K + 1, K), S + scp_tile_off(K + 1, K));
            WARP_PHASE_END
            WARP_PHASE(lane)
                if (w == 0 && lane == 0) tile_potrf(S + scp_tile_off(K + 1, K + 1), dinv + (K + 1) * 8, fixed);
                if (w > 0 || nw == 1) {
                    // tiles 1 .. ntile-1 over warps 1 .. nw-1 (a single-warp CTA does them itself), (ii, jj) advanced incrementally
                    const int nwork = nw > 1 ? nw - 1 : 1;
                    const double *Lk = S + scp_tile_off(K + 1, K);               // panel tile of row K+1+ii: + ii(ii+2K+3)/2 tiles
                    for (int t = nw > 1 ? w : 1; t < ntile; t += 2 * nwork) {
                        int ii, jj, i2, j2;
                        scp_tri_lookup(t, &ii, &jj);
                        double *C1 = S + scp_tile_off(K + 1 + ii, K + 1 + jj);
                        const double *A1 = Lk + (ii * (ii + 2 * K + 3) >> 1) * SCP_TILE2, *B1 = Lk + (jj * (jj + 2 * K + 3) >> 1) * SCP_TILE2;
                        if (t + nwork < ntile) {                                   // warp-uniform
                            scp_tri_lookup(t + nwork, &i2, &j2);
                            warp_tile_syrk2(lane, C1, A1, B1, S + scp_tile_off(K + 1 + i2, K + 1 + j2),
                                            Lk + (i2 * (i2 + 2 * K + 3) >> 1) * SCP_TILE2, Lk + (j2 * (j2 + 2 * K + 3) >> 1) * SCP_TILE2);
                        } else {
                            warp_tile_syrk(lane, C1, A1, B1);
                        }
                    }
                }
            WARP_PHASE_END
        WARP_SECTION_END
        CTA_SYNC
        SCP_TIMER(4)
    }
}

// ---- left-looking variant for a factor that lives in the L2-resident workspace (long horizons, Hp = 50: 679 KB) -------------
// The right-looking update above moves 1.8 KB per tile product through the SM's path to L2 (operands, C in, C out) and was
// measured to be bound by exactly that (1.06 M of the 2.23 M cycles of an iteration).  Left-looking, tile column K is brought up
// to date in one go,  S_IK -= sum_{J<K} L_IJ L_KJ'  (I = K .. T-1), with the sum ACCUMULATED IN REGISTERS over J: per product
// two operand fragments are read (the one of tile row K is shared by all the tiles a warp works on), nothing is written; C makes
// one round trip per tile instead of one per product.  Up to four tiles per warp and pass, so that their loads overlap.
// Warp 0 takes the diagonal tile first and factors it while the others finish the column.  Same products, summed before they are
// subtracted instead of one by one (rounding-level differences from chol_factor).
#define SCP_LEFT_NT 4
#if SCP_DEVICE_BUILD
// NT tiles (rows I0, I0 + dI, ...) of tile column K, UJ tile columns J per chunk: every operand fragment of a chunk is requested
// before its first product issues (2 UJ (1 + NT) loads in flight per lane) — with one J per round trip to L2 the update ran at
// exactly one L2 latency per product (ncu: 16 % of the samples on these loads, 14 % at the barrier behind them).
template <int NT, int UJ>
SCP_NOINLINE_FN void warp_column_update_dev(int lane, double *S, int K, int I0, int dI, int T)
{
    const int ia0 = scp_frag_rowmajor(lane, 0), ia1 = scp_frag_rowmajor(lane, 1), ic = scp_frag_c(lane);
    double acc[NT][2];
    const double *Ap[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        acc[t][0] = acc[t][1] = 0.0;
        const int I = I0 + t * dI;
        Ap[t] = S + scp_tile_off(I < T ? I : K, 0) + ia0;      // tiles (I, 0), (I, 1), ... are contiguous
    }
    const double *Bp = S + scp_tile_off(K, 0) + ia0;
    const int d1 = ia1 - ia0;
    // (Measured and not kept: a predicated last chunk instead of the remainder loop, with C requested before the products —
    // 1.07 M cycles per factorisation against 0.96 M for this version.)
    int J = 0;
    for (; J + UJ <= K; J += UJ) {
        double b[UJ][2], a[NT][UJ][2];
#pragma unroll
        for (int u = 0; u < UJ; ++u) { b[u][0] = Bp[(J + u) * SCP_TILE2]; b[u][1] = Bp[(J + u) * SCP_TILE2 + d1]; }
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if (I0 + t * dI < T) {
#pragma unroll
                for (int u = 0; u < UJ; ++u) { a[t][u][0] = Ap[t][(J + u) * SCP_TILE2]; a[t][u][1] = Ap[t][(J + u) * SCP_TILE2 + d1]; }
            }
#pragma unroll
        for (int u = 0; u < UJ; ++u)
#pragma unroll
            for (int t = 0; t < NT; ++t)
                if (I0 + t * dI < T) { scp_dmma(acc[t][0], acc[t][1], a[t][u][0], b[u][0]); scp_dmma(acc[t][0], acc[t][1], a[t][u][1], b[u][1]); }
    }
    for (; J < K; ++J) {
        const double b0 = Bp[J * SCP_TILE2], b1 = Bp[J * SCP_TILE2 + d1];
        double a0[NT], a1[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if (I0 + t * dI < T) { a0[t] = Ap[t][J * SCP_TILE2]; a1[t] = Ap[t][J * SCP_TILE2 + d1]; }
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if (I0 + t * dI < T) { scp_dmma(acc[t][0], acc[t][1], a0[t], b0); scp_dmma(acc[t][0], acc[t][1], a1[t], b1); }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t)
        if (I0 + t * dI < T) {
            double2 *cp = reinterpret_cast<double2 *>(S + scp_tile_off(I0 + t * dI, K) + ic);
            double2 c = *cp;
            c.x -= acc[t][0]; c.y -= acc[t][1];
            *cp = c;
        }
}
#endif
// single = true: the tile (I0, K) alone (the diagonal tile: deeper chunks); else up to SCP_LEFT_NT tiles I0, I0 + dI, ...
SCP_FN void warp_column_update(int lane, double *S, int K, int I0, int dI, int T, bool single)
{
#if SCP_DEVICE_BUILD
    if (single) warp_column_update_dev<1, 8>(lane, S, K, I0, T, T);
    else warp_column_update_dev<SCP_LEFT_NT, 2>(lane, S, K, I0, dI, T);
#else
    if (lane == 0)
        for (int t = 0; t < (single ? 1 : SCP_LEFT_NT); ++t) {
            const int I = I0 + t * dI;
            if (I >= T) continue;
            double *C = S + scp_tile_off(I, K);
            for (int r = 0; r < 8; ++r)
                for (int c = 0; c < 8; ++c) {
                    double acc = 0.0;
                    for (int J = 0; J < K; ++J) {
                        const double *A = S + scp_tile_off(I, J), *B = S + scp_tile_off(K, J);
                        for (int k = 0; k < 8; ++k) acc += A[scp_tphys(r, k)] * B[scp_tphys(c, k)];
                    }
                    C[scp_tphys(r, c)] -= acc;
                }
        }
#endif
}

SCP_FN void chol_factor_left(Cta &cta, const IpmMem &m, int *fixed SCP_TIMER_ARG)
{
    const int T = m.T;
    double *S = m.S, *dinv = m.dinv;
    for (int K = 0; K < T; ++K) {
        // (a) column K up to date; its diagonal tile factored by lane 0 of warp 0 meanwhile
        WARP_SECTION(w, nw)
            WARP_PHASE(lane)
                if (w == 0) warp_column_update(lane, S, K, K, T, T, true);       // the diagonal tile alone
            WARP_PHASE_END
            WARP_PHASE(lane)
                if (w == 0 && lane == 0) tile_potrf(S + scp_tile_off(K, K), dinv + K * 8, fixed);
                if (w > 0 || nw == 1) {
                    const int nwork = nw > 1 ? nw - 1 : 1;
                    // one tile per call, eight tile columns J per chunk (32 operand loads in flight per lane): more round trips to L2
                    // overlap than with four tiles and two columns per chunk (measured, profiles/r02_left_looking_*)
                    for (int I0 = K + 1 + (nw > 1 ? w - 1 : 0); I0 < T; I0 += nwork)
                        warp_column_update(lane, S, K, I0, nwork, T, true);
                }
            WARP_PHASE_END
        WARP_SECTION_END
        CTA_SYNC
        SCP_TIMER(4)
        // (b) panel rows: x L_KK' = s   (as in chol_factor)
        const double *Lkk = S + scp_tile_off(K, K);
        const int Tr = T - K - 1;
        CTA_PHASE(tid)
            for (int pr = tid; pr < Tr * 8; pr += cta.nt) {
                const int I = K + 1 + (pr >> 3), r = pr & 7;
                double *row = S + scp_tile_off(I, K) + (r << 3);
                const int h0 = ((r >> 1) & 1) << 2, h1 = h0 ^ 4;
                double x[8];
#pragma unroll
                for (int c = 0; c < 4; ++c) { x[c] = row[h0 + c]; x[c + 4] = row[h1 + c]; }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    x[c] *= dinv[K * 8 + c];
#pragma unroll
                    for (int c2 = c + 1; c2 < 8; ++c2) x[c2] -= x[c] * Lkk[scp_tphys(c2, c)];
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) { row[h0 + c] = x[c]; row[h1 + c] = x[c + 4]; }
            }
        CTA_PHASE_END
        SCP_TIMER(3)
    }
}

// ---- solves: inverted diagonal tiles + tile-wise substitution by one warp -------------------------------------------
// Substitution against L is a dependent chain over the n1 unknowns.  What was measured on B200 (n1p = 88, 3 CTAs per SM):
//   * full inverse X = L^-1 (round 1): two mat-vecs per solve (1.8 k cycles each) but n1^3/6 multiply-adds and 2 T CTA
//     barriers per factorisation (16 k cycles), and the explicit inverse limits the attainable dual residual;
//   * plain tile-wise substitution by one warp, diagonal tile solved by one lane: 14 k cycles per triangular solve;
//   * 32 x 32 diagonal blocks inverted, block substitution with lane = row: 15 k (the per-lane row-dots serialise on
//     shared-memory latency: the compiler keeps one load in flight).
//   * the whole substitution inside ONE warp with the vector in registers (row i = lane + 32 m), the 8 unknowns of a tile
//     column gathered and broadcast by shuffles, no barrier (round 2): ~260 instructions per tile column, which a lone
//     warp issues at ~5 cycles each (every instruction waits on the one before: shuffle 25, LDS 30, DFMA 9) = 1.3 k cycles per
//     column against 715 for the version below; the step fell from 603 k to 433 k QP/s (ncu: 25 % of the warp samples were the
//     other three warps waiting for warp 0).  Not kept.
//   * one CTA barrier per BLOCK of four tile columns instead of one per column (round 2; warp 0 solves the block with
//     warp-level synchronisation only, the other warps take the previous block's 32 unknowns out of everything beyond it):
//     25.1 k cycles per iteration against 23.6 k, 578 k QP/s against 590-603 k.  The barrier is not what a column costs: it
//     is the three dependent shared-memory round trips inside warp 0 (row-dot -> store -> row-dot), which blocking keeps.
//     Not kept.
//   * long horizons (factor in the L2-resident workspace, Hp = 50): the next step's rows / columns fetched into registers
//     before the barrier that ends a step (round 2): 517 k cycles per iteration with and without.  Not kept.
// This version: only the 8 x 8 DIAGONAL TILES are inverted (in place, after the factorisation: one phase), and a
// triangular solve is T steps inside warp 0 with no CTA barrier: 8 lanes apply the tile inverse (an 8 x 8 mat-vec), every
// lane then takes the 8 new unknowns out of its rows below (forward) / columns to the left (backward).  All loads of a
// step are issued before its arithmetic.
#if SCP_DEVICE_BUILD
#define SCP_LDS_FENCE asm volatile("" ::: "memory");      /* keeps the loads above, the arithmetic below */
#endif

// Replace every diagonal tile L_KK of the factor by its inverse (lower triangular, zeros above the diagonal); tiles
// tile_base .. tile_base + 3 by this warp, lane = (tile, column).  yrow (n1p entries, or null): the warp that owns the last
// tile first copies out row n1p - 1 (the forward-substituted right-hand side, see chol_factor), whose last entries live in
// that tile.
SCP_FN void warp_invert_diag_tiles(int lane, double *S, const double *dinv, int T, int tile_base, double *yrow)
{
#if SCP_DEVICE_BUILD
    const int Kreal = tile_base + (lane >> 3), K = Kreal < T ? Kreal : T - 1, j = lane & 7;
    double *Tk = S + scp_tile_off(K, K);
    if (yrow && tile_base <= T - 1 && T - 1 < tile_base + 4) {
        const int n1p = T * 8;
        for (int c = lane; c < n1p; c += 32) yrow[c] = c < n1p - 1 ? S[scp_sidx(n1p - 1, c)] : 0.0;
    }
    double xcol[8];
    tile_trtri_column(Tk, dinv + K * 8, j, xcol);
    __syncwarp();
    if (Kreal < T) {
#pragma unroll
        for (int i = 0; i < 8; ++i) Tk[scp_tphys(i, j)] = xcol[i];
    }
#else
    if (lane == 0) {
        if (yrow && tile_base <= T - 1 && T - 1 < tile_base + 4) {
            const int n1p = T * 8;
            for (int c = 0; c < n1p; ++c) yrow[c] = c < n1p - 1 ? S[scp_sidx(n1p - 1, c)] : 0.0;
        }
        for (int K = tile_base; K < tile_base + 4 && K < T; ++K) {
            double *Tk = S + scp_tile_off(K, K), X[64];
            for (int j = 0; j < 8; ++j) {
                double xcol[8];
                tile_trtri_column(Tk, dinv + K * 8, j, xcol);
                for (int i = 0; i < 8; ++i) X[i * 8 + j] = xcol[i];
            }
            for (int i = 0; i < 8; ++i)
                for (int j = 0; j < 8; ++j) Tk[scp_tphys(i, j)] = X[i * 8 + j];
        }
    }
#endif
}

SCP_FN void chol_invert_diag(Cta &cta, const IpmMem &m, double *yrow)
{
    WARP_SECTION(w, nw)
        for (int base = 4 * w; base < m.T; base += 4 * nw) {
            WARP_PHASE(lane)
                warp_invert_diag_tiles(lane, m.S, m.dinv, m.T, base, yrow);
            WARP_PHASE_END
        }
    WARP_SECTION_END
    CTA_SYNC
}

// data a lane keeps in a register across the warp phases of one section (the host build runs the lanes of a phase one
// after the other, so there every lane needs its own copy)
#if SCP_DEVICE_BUILD
#define SCP_LANE_VAR(name) double name##_store = 0.0;
#define SCP_LANE_REF(name, lane) name##_store
#else
#define SCP_LANE_VAR(name) double name##_store[32] = {0.0};
#define SCP_LANE_REF(name, lane) name##_store[lane]
#endif

// row i of the factor against the 8 unknowns of tile column K:  sum_c L[i][8K + c] x[c]
SCP_FN double trsv_row_dot(const double *S, int i, int K, const double *x)
{
    const int r = i & 7, h0 = ((r >> 1) & 1) << 2, h1 = h0 ^ 4;
    const double *row = S + scp_tile_off(i >> 3, K) + (r << 3);
#if SCP_DEVICE_BUILD
    const double2 *row2 = reinterpret_cast<const double2 *>(row), *x2 = reinterpret_cast<const double2 *>(x);
    const double2 r0 = row2[h0 >> 1], r1 = row2[(h0 >> 1) + 1], r2 = row2[h1 >> 1], r3 = row2[(h1 >> 1) + 1];
    const double2 x0 = x2[0], x1 = x2[1], x2v = x2[2], x3 = x2[3];
    SCP_LDS_FENCE
    return ((r0.x * x0.x + r0.y * x0.y) + (r1.x * x1.x + r1.y * x1.y)) + ((r2.x * x2v.x + r2.y * x2v.y) + (r3.x * x3.x + r3.y * x3.y));
#else
    return ((row[h0] * x[0] + row[h0 + 1] * x[1]) + (row[h0 + 2] * x[2] + row[h0 + 3] * x[3])) +
           ((row[h1] * x[4] + row[h1 + 1] * x[5]) + (row[h1 + 2] * x[6] + row[h1 + 3] * x[7]));
#endif
}
// column j of tile row K of the factor against the 8 unknowns of that tile row:  sum_r L[8K + r][j] x[r]
SCP_FN double trsv_col_dot(const double *S, int K, int j, const double *x)
{
    const int c = j & 7, lo = c & 3, he = (c >> 2) << 2, ho = he ^ 4;       // rows 0,1,4,5 / rows 2,3,6,7
    const double *tl = S + scp_tile_off(K, j >> 3);
    const double c0 = tl[he + lo], c1 = tl[8 + he + lo], c2 = tl[16 + ho + lo], c3 = tl[24 + ho + lo];
    const double c4 = tl[32 + he + lo], c5 = tl[40 + he + lo], c6 = tl[48 + ho + lo], c7 = tl[56 + ho + lo];
#if SCP_DEVICE_BUILD
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    const double2 x0 = x2[0], x1 = x2[1], x2v = x2[2], x3 = x2[3];
    SCP_LDS_FENCE
    return ((c0 * x0.x + c1 * x0.y) + (c2 * x1.x + c3 * x1.y)) + ((c4 * x2v.x + c5 * x2v.y) + (c6 * x3.x + c7 * x3.y));
#else
    return ((c0 * x[0] + c1 * x[1]) + (c2 * x[2] + c3 * x[3])) + ((c4 * x[4] + c5 * x[5]) + (c6 * x[6] + c7 * x[7]));
#endif
}

// v := S^-1 v = L^-T (L^-1 v)   (have_y = false), or
// v := L^-T y with y (already in v) the forward-substituted right-hand side chol_invert_diag extracted   (have_y = true),
// for the factor prepared by chol_factor + chol_invert_diag.  v: n1p entries (zero padding).
//
// Tile-wise substitution across the CTA, ONE barrier per tile column, the dependent chain kept inside 8 lanes of warp 0:
// in step K those lanes bring the rows of tile K up to date with the unknowns of tile K-1 (found in the previous step),
// exchange them, and apply the inverse of the diagonal tile (an 8 x 8 mat-vec) to get the unknowns of tile K; meanwhile
// every other thread takes the unknowns of tile K-1 out of one row further down.  Backward likewise with columns.
#if SCP_DEVICE_BUILD
// ---- sweeps against a factor in the L2-resident workspace (long horizons) ---------------------------------------------------
// With the factor in shared memory a step of the sweep is bound by arithmetic and exchange inside warp 0 (the register-resident
// variant of this routine was measured there and not kept: profiles/r02_pipelined_sweeps_experiment.txt).  With the factor in L2
// every operand costs a round trip of ~700 cycles, and the phase version below pays two of them in sequence inside warp 0 plus
// one in every other thread per tile column (ncu at Hp = 50: 22 % of all warp samples wait at the barrier that ends a step).  The
// factor is constant during a sweep, so here EVERY operand of step K+1 — the rows (columns) of the off-diagonal and of the
// inverted diagonal tile for warp 0, one row (column) of the off-diagonal tile for every other thread — is requested before the
// barrier that ends step K, and a step touches only registers and shared memory.  Warp 0 runs nothing but the chain (all 32
// lanes, full-mask shuffles; lanes 0-7 own the tile); thread 32 + w owns row 8 + w (forward) / column w (backward) for the whole
// sweep; rows beyond that range (n1p > 8 + threads - 32) are handled without prefetch.  Same arithmetic as the phase version.
SCP_FN double trsv_tree8(const double (&a)[8], const double (&x)[8])
{
    return ((a[0] * x[0] + a[1] * x[1]) + (a[2] * x[2] + a[3] * x[3])) + ((a[4] * x[4] + a[5] * x[5]) + (a[6] * x[6] + a[7] * x[7]));
}
SCP_FN void trsv_load_row(const double *tl, int r, double (&a)[8])      // row r of a tile in logical column order
{
    const int h0 = ((r >> 1) & 1) << 2, h1 = h0 ^ 4;
    const double2 *row2 = reinterpret_cast<const double2 *>(tl + (r << 3));
    const double2 r0 = row2[h0 >> 1], r1 = row2[(h0 >> 1) + 1], r2 = row2[h1 >> 1], r3 = row2[(h1 >> 1) + 1];
    a[0] = r0.x; a[1] = r0.y; a[2] = r1.x; a[3] = r1.y; a[4] = r2.x; a[5] = r2.y; a[6] = r3.x; a[7] = r3.y;
}
SCP_FN void trsv_load_col(const double *tl, int c, double (&a)[8])      // column c of a tile in logical row order
{
    const int lo = c & 3, he = (c >> 2) << 2, ho = he ^ 4;               // rows 0,1,4,5 / rows 2,3,6,7
    a[0] = tl[he + lo]; a[1] = tl[8 + he + lo]; a[2] = tl[16 + ho + lo]; a[3] = tl[24 + ho + lo];
    a[4] = tl[32 + he + lo]; a[5] = tl[40 + he + lo]; a[6] = tl[48 + ho + lo]; a[7] = tl[56 + ho + lo];
}
SCP_FN void trsv_load_x(const double *xp, double (&x)[8])
{
    const double2 *x2 = reinterpret_cast<const double2 *>(xp);
    const double2 x0 = x2[0], x1 = x2[1], x2v = x2[2], x3 = x2[3];
    x[0] = x0.x; x[1] = x0.y; x[2] = x1.x; x[3] = x1.y; x[4] = x2v.x; x[5] = x2v.y; x[6] = x3.x; x[7] = x3.y;
}
SCP_FN void chol_solve_far(Cta &cta, const IpmMem &m, double *v, bool have_y)
{
    const int T = m.T, n1p = m.n1p;
    const double *S = m.S;
    const int tid = (int)threadIdx.x, lane = tid & 31, nw = cta.nt >> 5;
    const bool crit = tid < 32;
    const int l8 = lane & 7;
    const int nwork = nw > 1 ? cta.nt - 32 : 24;         // a single-warp CTA: lanes 8 .. 31 take the rows beyond after the chain
    const int wid = nw > 1 ? tid - 32 : lane - 8;        // < 0: not a worker
    for (int dir = have_y ? 1 : 0; dir < 2; ++dir) {
        if (dir == 1) {
            if (tid == 0) v[n1p - 1] = 0.0;              // the right-hand-side row is not part of the system
            __syncthreads();
        }
        // the worker's own row (forward: 8 + wid, beyond tile K while >= 8 (K + 1)) or column (backward: wid, beyond while < 8 K)
        const int wi = dir == 0 ? 8 + wid : wid;
        const bool wown = wid >= 0 && wi < n1p;
        double xk = 0.0, a[8], d[8], wa[8];
        if (crit) {
            const int K0 = dir == 0 ? 0 : T - 1;
            if (dir == 0) trsv_load_row(S + scp_tile_off(K0, K0), l8, d); else trsv_load_col(S + scp_tile_off(K0, K0), l8, d);
#pragma unroll
            for (int c = 0; c < 8; ++c) a[c] = 0.0;
        }
        if (wown && T > 1) {                             // operands of step 1
            if (dir == 0) { if (wi >= 16) trsv_load_row(S + scp_tile_off(wi >> 3, 0), wi & 7, wa); }
            else { if (wi < (T - 2) * 8) trsv_load_col(S + scp_tile_off(T - 1, wi >> 3), wi & 7, wa); }
        }
        for (int step = 0; step < T; ++step) {
            const int K = dir == 0 ? step : T - 1 - step;    // tile whose unknowns this step finds
            const int Kp = dir == 0 ? K - 1 : K + 1;         // tile found in the previous step
            const int Kn = dir == 0 ? K + 1 : K - 1;         // tile of the next step
            if (crit) {
                double r = v[K * 8 + l8];
                double x[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) x[c] = __shfl_sync(0xffffffffu, xk, c);
                if (step > 0) r -= trsv_tree8(a, x);
#pragma unroll
                for (int c = 0; c < 8; ++c) x[c] = __shfl_sync(0xffffffffu, r, c);
                xk = trsv_tree8(d, x);
                if (lane < 8) v[K * 8 + lane] = xk;
                if (step + 1 < T) {
                    if (dir == 0) {
                        trsv_load_row(S + scp_tile_off(Kn, K), l8, a);
                        trsv_load_row(S + scp_tile_off(Kn, Kn), l8, d);
                    } else {
                        trsv_load_col(S + scp_tile_off(K, Kn), l8, a);
                        trsv_load_col(S + scp_tile_off(Kn, Kn), l8, d);
                    }
                }
            }
            if (wid >= 0 && step > 0) {
                const double *xp = v + Kp * 8;
                const bool act = dir == 0 ? wi >= (K + 1) * 8 : wi < K * 8;
                if (wown && act) {
                    double x[8];
                    trsv_load_x(xp, x);
                    v[wi] -= trsv_tree8(wa, x);
                }
                // rows / columns beyond the one-per-thread range (not prefetched)
                if (dir == 0) {
                    for (int i = wi + nwork; i < n1p; i += nwork)
                        if (i >= (K + 1) * 8) v[i] -= trsv_row_dot(S, i, Kp, xp);
                } else {
                    for (int j = wi + nwork; j < K * 8; j += nwork) v[j] -= trsv_col_dot(S, Kp, j, xp);
                }
            }
            if (wown && step + 1 < T) {                      // the worker's operand of step K+1: tile column K (forward) / tile row K (backward)
                if (dir == 0) { if (wi >= (Kn + 1) * 8) trsv_load_row(S + scp_tile_off(wi >> 3, K), wi & 7, wa); }
                else { if (wi < Kn * 8) trsv_load_col(S + scp_tile_off(K, wi >> 3), wi & 7, wa); }
            }
            __syncthreads();
        }
    }
}
#endif

SCP_FN void chol_solve(Cta &cta, const IpmMem &m, double *v, bool have_y)
{
#if SCP_DEVICE_BUILD
    if (m.S_far) { chol_solve_far(cta, m, v, have_y); return; }
#endif
    const int T = m.T, n1p = m.n1p;
    const double *S = m.S;
    const int nwork = cta.nt - 8;                        // threads that work on the rows / columns beyond the current tile
    for (int dir = have_y ? 1 : 0; dir < 2; ++dir) {
        if (dir == 1) {
            CTA_PHASE(tid)
                if (tid == 0) v[n1p - 1] = 0.0;              // the right-hand-side row is not part of the system
            CTA_PHASE_END
        }
        for (int step = 0; step < T; ++step) {
            const int K = dir == 0 ? step : T - 1 - step;    // tile whose unknowns this step finds
            const int Kp = dir == 0 ? K - 1 : K + 1;         // tile found in the previous step
            WARP_SECTION(w, nw)
                (void)nw;
                SCP_LANE_VAR(xk)
                WARP_PHASE(lane)
                    const int tid = w * 32 + lane;
                    if (step > 0) {
                        const double *xp = v + Kp * 8;
                        if (tid < 8) {
                            v[K * 8 + tid] -= dir == 0 ? trsv_row_dot(S, K * 8 + tid, Kp, xp) : trsv_col_dot(S, Kp, K * 8 + tid, xp);
                        } else {
                            if (dir == 0) {
                                for (int i = (K + 1) * 8 + (tid - 8); i < n1p; i += nwork) v[i] -= trsv_row_dot(S, i, Kp, xp);
                            } else {
                                for (int j = tid - 8; j < K * 8; j += nwork) v[j] -= trsv_col_dot(S, Kp, j, xp);
                            }
                        }
                    }
                WARP_PHASE_END
                WARP_PHASE(lane)
                    const int tid = w * 32 + lane;
                    if (tid < 8)       // the diagonal tile holds its inverse
                        SCP_LANE_REF(xk, lane) = dir == 0 ? trsv_row_dot(S, K * 8 + tid, K, v + K * 8) : trsv_col_dot(S, K, K * 8 + tid, v + K * 8);
                WARP_PHASE_END
                WARP_PHASE(lane)
                    const int tid = w * 32 + lane;
                    if (tid < 8) v[K * 8 + tid] = SCP_LANE_REF(xk, lane);
                WARP_PHASE_END
            WARP_SECTION_END
            CTA_SYNC
        }
    }
}

// ------------------------------------------------------------------------------------------------ helpers
SCP_FN bool ipm_has_ub(const IpmMem &m, const IpmCtl &ctl, int c) { return c < m.n1 && fabs(m.ub[c]) < ctl.inf_bound; }
SCP_FN bool ipm_has_lb(const IpmMem &m, const IpmCtl &ctl, int c) { return c < m.n1 && fabs(m.lb[c]) < ctl.inf_bound; }

// S := 0 with unit diagonal on the padding (helper for operators that accumulate into S)
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
// Per-row arithmetic of one direction computation, shared by collision and box rows.  With
//     e = 1/(s + delta z),  dd = z e  (= 1/(s/z + delta)),  bs = -s z + sigma mu - cc
// the regularised Newton system gives  w1 = dd rz + bs e,  dz = w1 + dd (G dx),  ds = (bs - s dz)/z.
SCP_FN double ipm_w1(double s, double z, double rz, double e, double bs) { return z * e * rz + bs * e; }

// Right-hand side of one direction computation (pass 0: affine direction, sigma mu = 0; pass 1: centring + corrector):
//   w1 (per row, kept in dz) and  m.dx = -rx - G'w1,  padding zero.  Two phases and the operator's preparation.
template <class Op>
SCP_FN void ipm_pass_rhs(Cta &cta, Op &op, const IpmMem &m, const IpmCtl &ctl, int pass, double smu)
{
    const int n1 = m.n1, n1p = m.n1p, mc = m.mc;
    // w1 for the collision rows (input of A'w1); bs = -s z + sigma mu - [pass 1] (ds_a dz_a)
    CTA_PHASE(tid)
        for (int r = tid; r < mc; r += cta.nt) {
            const double s = m.sA[r], z = m.zA[r];
            double bs = smu - s * z;
            if (pass == 1) bs -= m.ccA[r];
            m.dzA[r] = ipm_w1(s, z, m.rzA[r], m.eA[r], bs);                       // w1 in dzA
        }
        for (int c = tid; c < m.nr; c += cta.nt) {
            const double df = ipm_rdiff(m, m.x, c);
            {
                const double s = m.sP[c], z = m.zP[c];
                double bs = smu - s * z;
                if (pass == 1) bs -= m.ccP[c];
                m.dzP[c] = ipm_w1(s, z, s + df - m.hP[c], m.eP[c], bs);
            }
            {
                const double s = m.sM[c], z = m.zM[c];
                double bs = smu - s * z;
                if (pass == 1) bs -= m.ccM[c];
                m.dzM[c] = ipm_w1(s, z, s - df - m.hM[c], m.eM[c], bs);
            }
        }
    CTA_PHASE_END
    op.prep(cta, (const double *)0, m.dzA);
    // rhs = -rx - G'w1
    CTA_PHASE(tid)
        for (int c = tid; c < n1p; c += cta.nt) {
            double rhs = 0.0;
            if (c < n1) rhs = -m.rx[c] - op.col_dot(c);
            if (c < m.nr) rhs -= ipm_rcol(m, m.dzP, m.dzM, c);
            if (ipm_has_ub(m, ctl, c)) {
                const double s = m.sU[c], z = m.zU[c];
                double bs = smu - s * z;
                if (pass == 1) bs -= m.ccU[c];
                const double w1 = ipm_w1(s, z, s + m.x[c] - m.ub[c], m.eU[c], bs);
                m.dzU[c] = w1; rhs -= w1;
            }
            if (ipm_has_lb(m, ctl, c)) {
                const double s = m.sL[c], z = m.zL[c];
                double bs = smu - s * z;
                if (pass == 1) bs -= m.ccL[c];
                const double w1 = ipm_w1(s, z, s - m.x[c] + m.lb[c], m.eL[c], bs);
                m.dzL[c] = w1; rhs += w1;
            }
            m.dx[c] = rhs;
        }
    CTA_PHASE_END
}

// On entry m.q, m.bA, m.ub, m.lb hold the problem data (padding of q/ub/lb beyond n1 is ignored).
// On exit m.x holds the solution.
//
// One iteration = residuals -> right-hand side of the affine direction -> normal matrix (that right-hand side rides as
// its last row) -> factorisation -> [back substitution, step length] -> right-hand side of the corrector -> [forward +
// back substitution, step length] -> update.
template <class Op>
SCP_FN void ipm_solve(Cta &cta, Op &op, const IpmMem &m, const IpmCtl &ctl, IpmResult *res)
{
    const int n1 = m.n1, n1p = m.n1p, mc = m.mc;
    double *red = m.red;
    SCP_TIMER_DECL
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
        for (int c = tid; c < m.nr; c += cta.nt) { cnt += 2.0; hh += m.hP[c] * m.hP[c] + m.hM[c] * m.hM[c]; }
        CTA_RED_SUM(cta, red, 0, tid, cnt)
        CTA_RED_SUM(cta, red, 1, tid, hq)
        CTA_RED_SUM(cta, red, 2, tid, hh)
    CTA_PHASE_END_RED(cta, red, 3)
    const double mrows = cta_red_sum(cta, red, 0) + (double)mc;
    const double resx0 = fmax(1.0, sqrt(cta_red_sum(cta, red, 1)));
    const double resz0 = fmax(1.0, sqrt(cta_red_sum(cta, red, 2)));

    if (ctl.warm) {
        ipm_snapshot(cta, m, ctl.snap, false);
    } else {
        // ---- starting point (coneqp): (P + G'G) x = G'h - q ; z = Gx - h ; s = -z ; shift ------------
        SCP_TIMER(0)
        op.prep(cta, (const double *)0, m.bA);
        CTA_PHASE(tid)
            for (int c = tid; c < n1p; c += cta.nt) {
                double rhs = 0.0;
                if (c < n1) rhs = op.col_dot(c) - m.q[c];
                if (c < m.nr) rhs += ipm_rcol(m, m.hP, m.hM, c);
                if (ipm_has_ub(m, ctl, c)) rhs += m.ub[c];
                if (ipm_has_lb(m, ctl, c)) rhs += m.lb[c];
                m.x[c] = rhs;
            }
        CTA_PHASE_END
        op.form_normal(cta, m, m.dsA, m.tn, m.x, [](int) { return 1.0; },                    // dd = 1
                       [&](int c) {
                           double d = (ipm_has_ub(m, ctl, c) ? 1.0 : 0.0) + (ipm_has_lb(m, ctl, c) ? 1.0 : 0.0);
                           if (c < m.nr) {                                                // D'D of both difference rows
                               d += 2.0 + ((c + 1 < m.nr && !ipm_rfirst(m, c + 1)) ? 2.0 : 0.0);
                               m.rsub[c] = ipm_rfirst(m, c) ? 0.0 : -2.0;
                           }
                           return d;
                       } SCP_TIMER_PASS);
        SCP_TIMER(1)
        if (m.S_far) chol_factor_left(cta, m, fixed_p SCP_TIMER_PASS); else chol_factor(cta, m, fixed_p SCP_TIMER_PASS);
        chol_invert_diag(cta, m, m.x);
        chol_solve(cta, m, m.x, true);
        SCP_TIMER(8)
        op.prep(cta, m.x, (const double *)0);
        CTA_RED_BEGIN(cta, 3)
        CTA_PHASE(tid)
            double nrm = 0.0, ts = -1e300, tz = -1e300;
            for (int r = tid; r < mc; r += cta.nt) {
                const double z = op.row_dot(r) - m.bA[r];
                m.zA[r] = z; m.sA[r] = -z;
                nrm += z * z; ts = fmax(ts, z); tz = fmax(tz, -z);
            }
            for (int c = tid; c < m.nr; c += cta.nt) {
                const double df = ipm_rdiff(m, m.x, c);
                const double zp = df - m.hP[c], zm = -df - m.hM[c];
                m.zP[c] = zp; m.sP[c] = -zp; m.zM[c] = zm; m.sM[c] = -zm;
                nrm += zp * zp + zm * zm; ts = fmax(ts, fmax(zp, zm)); tz = fmax(tz, fmax(-zp, -zm));
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
        double as_shift, az_shift;
        {
            const double nrm = sqrt(cta_red_sum(cta, red, 0));
            const double ts = cta_red_max(cta, red, 1), tz = cta_red_max(cta, red, 2);
            const double thr = -1e-8 * fmax(nrm, 1.0);
            as_shift = (ts >= thr) ? 1.0 + ts : 0.0;
            az_shift = (tz >= thr) ? 1.0 + tz : 0.0;
        }
        CTA_PHASE(tid)
            for (int r = tid; r < mc; r += cta.nt) { m.sA[r] += as_shift; m.zA[r] += az_shift; }
            for (int c = tid; c < m.nr; c += cta.nt) {
                m.sP[c] += as_shift; m.zP[c] += az_shift; m.sM[c] += as_shift; m.zM[c] += az_shift;
            }
            for (int c = tid; c < n1p; c += cta.nt) {
                if (ipm_has_ub(m, ctl, c)) { m.sU[c] += as_shift; m.zU[c] += az_shift; }
                if (ipm_has_lb(m, ctl, c)) { m.sL[c] += as_shift; m.zL[c] += az_shift; }
            }
        CTA_PHASE_END
    }

    int iters = 0, status = SCPB200_ST_QP_MAXITER, snap_saved = 0;
    double f0 = 0.0, gap = 0.0, relgap = -1.0, pres = 0.0, dres = 0.0, dres_prev = 1e300;
    for (iters = 0; iters <= ctl.max_iter; ++iters) {
        // ---- residuals: rx = Px + q + G'z ; rz = s + Gx - h ; costs; e = 1/(s + delta z) -----------
        SCP_TIMER(0)
        op.prep(cta, m.x, m.zA);
        CTA_RED_BEGIN(cta, 5)
        CTA_PHASE(tid)
            double pf = 0.0, pg = 0.0, prz = 0.0, pzr = 0.0, prx = 0.0;
            for (int c = tid; c < n1p; c += cta.nt) {
                double r = 0.0;
                if (c < n1) {
                    const double px = op.P_col(c, m.x);
                    pf += m.x[c] * (0.5 * px + m.q[c]);
                    r = px + m.q[c] + op.col_dot(c);
                }
                if (c < m.nr) {
                    r += ipm_rcol(m, m.zP, m.zM, c);
                    const double df = ipm_rdiff(m, m.x, c);
                    {
                        const double s = m.sP[c], z = m.zP[c], rz = s + df - m.hP[c];
                        pg += s * z; prz += rz * rz; pzr += z * rz;
                        m.eP[c] = 1.0 / (s + ctl.dual_reg * z);
                    }
                    {
                        const double s = m.sM[c], z = m.zM[c], rz = s - df - m.hM[c];
                        pg += s * z; prz += rz * rz; pzr += z * rz;
                        m.eM[c] = 1.0 / (s + ctl.dual_reg * z);
                    }
                }
                if (ipm_has_ub(m, ctl, c)) {
                    const double s = m.sU[c], z = m.zU[c], rz = s + m.x[c] - m.ub[c];
                    r += z; pg += s * z; prz += rz * rz; pzr += z * rz;
                    m.eU[c] = 1.0 / (s + ctl.dual_reg * z);
                }
                if (ipm_has_lb(m, ctl, c)) {
                    const double s = m.sL[c], z = m.zL[c], rz = s - m.x[c] + m.lb[c];
                    r -= z; pg += s * z; prz += rz * rz; pzr += z * rz;
                    m.eL[c] = 1.0 / (s + ctl.dual_reg * z);
                }
                m.rx[c] = r;
                prx += r * r;
            }
            for (int r = tid; r < mc; r += cta.nt) {
                const double s = m.sA[r], z = m.zA[r];
                const double rz = s + op.row_dot(r) - m.bA[r];
                m.rzA[r] = rz;
                pg += s * z; prz += rz * rz; pzr += z * rz;
                m.eA[r] = 1.0 / (s + ctl.dual_reg * z);
            }
            CTA_RED_SUM(cta, red, 0, tid, pf)
            CTA_RED_SUM(cta, red, 1, tid, pg)
            CTA_RED_SUM(cta, red, 2, tid, prz)
            CTA_RED_SUM(cta, red, 3, tid, pzr)
            CTA_RED_SUM(cta, red, 4, tid, prx)
        CTA_PHASE_END_RED(cta, red, 5)
        f0 = cta_red_sum(cta, red, 0);
        gap = cta_red_sum(cta, red, 1);
        const double resz = sqrt(cta_red_sum(cta, red, 2));
        const double zrz = cta_red_sum(cta, red, 3);
        const double resx = sqrt(cta_red_sum(cta, red, 4));
        const double dcost = f0 + zrz - gap;
        if (f0 < 0.0) relgap = gap / -f0;
        else if (dcost > 0.0) relgap = gap / dcost;
        else relgap = -1.0;
        pres = resz / resz0;
        dres = resx / resx0;
        const bool gap_ok = gap <= ctl.abstol || (relgap >= 0.0 && relgap <= ctl.reltol);
        if (pres <= ctl.feastol && dres <= ctl.feastol && gap_ok) { status = 0; break; }
        // Precision floor of the dual residual: gap and primal residual have converged, the dual residual is within
        // qp_dres_floor_factor x feastol and no longer decreasing.  Iterating on drives s.z to underflow and the residual back up (measured
        // at Hp = 20 / 50: gap 1e-90, dres 1e-7 after 60 iterations); accept the iterate and say so.
        if (gap_ok && pres <= ctl.feastol && dres <= ctl.dres_floor && iters > 0 && dres >= 0.5 * dres_prev) {
            status = SCPB200_ST_QP_DRES_FLOOR;
            break;
        }
        dres_prev = dres;
        if (iters == ctl.max_iter) break;
        if (ctl.snap && !snap_saved && iters >= ctl.snap_min_iter && relgap >= 0.0 && relgap <= ctl.snap_relgap) {
            ipm_snapshot(cta, m, ctl.snap, true);
            snap_saved = 1;
        }
        SCP_TIMER(9)

        // ---- affine right-hand side, then the normal matrix with dd = z e (the right-hand side as its last row) ----
        ipm_pass_rhs(cta, op, m, ctl, 0, 0.0);
        SCP_TIMER(10)
        op.form_normal(cta, m, m.dsA, m.tn, m.dx, [&](int r) { return m.zA[r] * m.eA[r]; },       // dd (kept in dsA)
                       [&](int c) {
                           double d = 0.0;
                           if (ipm_has_ub(m, ctl, c)) d += m.zU[c] * m.eU[c];
                           if (ipm_has_lb(m, ctl, c)) d += m.zL[c] * m.eL[c];
                           if (c < m.nr) {                          // D' diag(dP + dM) D: tridiagonal inside a vehicle block
                               const double tc = m.zP[c] * m.eP[c] + m.zM[c] * m.eM[c];
                               d += tc;
                               if (c + 1 < m.nr && !ipm_rfirst(m, c + 1)) d += m.zP[c + 1] * m.eP[c + 1] + m.zM[c + 1] * m.eM[c + 1];
                               m.rsub[c] = ipm_rfirst(m, c) ? 0.0 : -tc;
                           }
                           return d;
                       } SCP_TIMER_PASS);
        SCP_TIMER(1)
        if (m.S_far) chol_factor_left(cta, m, fixed_p SCP_TIMER_PASS); else chol_factor(cta, m, fixed_p SCP_TIMER_PASS);
        chol_invert_diag(cta, m, m.dx);
        SCP_TIMER(5)

        const double mu = gap / mrows;
        double sigma = 0.0, step = 1.0;
        for (int pass = 0; pass < 2; ++pass) {
            const double smu = sigma * mu;
            if (pass == 1) {
                ipm_pass_rhs(cta, op, m, ctl, 1, smu);
                SCP_TIMER(10)
            }
            chol_solve(cta, m, m.dx, pass == 0);
            SCP_TIMER(8)
            op.prep(cta, m.dx, (const double *)0);
            CTA_RED_BEGIN(cta, 3)
            CTA_PHASE(tid)
                double pdd = 0.0, ts = 0.0, tz = 0.0;
                for (int r = tid; r < mc; r += cta.nt) {
                    const double s = m.sA[r], z = m.zA[r];
                    double bs = smu - s * z;
                    if (pass == 1) bs -= m.ccA[r];
                    const double pinv = 1.0 / (s * z);                                    // 1/z = pinv s, 1/s = pinv z
                    const double dz = m.dzA[r] + z * m.eA[r] * op.row_dot(r);
                    const double ds = (bs - s * dz) * (pinv * s);
                    m.dzA[r] = dz; m.dsA[r] = ds;
                    pdd += ds * dz; ts = fmax(ts, -ds * (pinv * z)); tz = fmax(tz, -dz * (pinv * s));
                    if (pass == 0) m.ccA[r] = ds * dz;
                }
                for (int c = tid; c < m.nr; c += cta.nt) {
                    const double gdx = ipm_rdiff(m, m.dx, c);
                    {
                        const double s = m.sP[c], z = m.zP[c];
                        double bs = smu - s * z;
                        if (pass == 1) bs -= m.ccP[c];
                        const double pinv = 1.0 / (s * z);
                        const double dz = m.dzP[c] + z * m.eP[c] * gdx;
                        const double ds = (bs - s * dz) * (pinv * s);
                        m.dzP[c] = dz; m.dsP[c] = ds;
                        pdd += ds * dz; ts = fmax(ts, -ds * (pinv * z)); tz = fmax(tz, -dz * (pinv * s));
                        if (pass == 0) m.ccP[c] = ds * dz;
                    }
                    {
                        const double s = m.sM[c], z = m.zM[c];
                        double bs = smu - s * z;
                        if (pass == 1) bs -= m.ccM[c];
                        const double pinv = 1.0 / (s * z);
                        const double dz = m.dzM[c] - z * m.eM[c] * gdx;
                        const double ds = (bs - s * dz) * (pinv * s);
                        m.dzM[c] = dz; m.dsM[c] = ds;
                        pdd += ds * dz; ts = fmax(ts, -ds * (pinv * z)); tz = fmax(tz, -dz * (pinv * s));
                        if (pass == 0) m.ccM[c] = ds * dz;
                    }
                }
                for (int c = tid; c < n1p; c += cta.nt) {
                    if (ipm_has_ub(m, ctl, c)) {
                        const double s = m.sU[c], z = m.zU[c];
                        double bs = smu - s * z;
                        if (pass == 1) bs -= m.ccU[c];
                        const double pinv = 1.0 / (s * z);
                        const double dz = m.dzU[c] + z * m.eU[c] * m.dx[c];
                        const double ds = (bs - s * dz) * (pinv * s);
                        m.dzU[c] = dz; m.dsU[c] = ds;
                        pdd += ds * dz; ts = fmax(ts, -ds * (pinv * z)); tz = fmax(tz, -dz * (pinv * s));
                        if (pass == 0) m.ccU[c] = ds * dz;
                    }
                    if (ipm_has_lb(m, ctl, c)) {
                        const double s = m.sL[c], z = m.zL[c];
                        double bs = smu - s * z;
                        if (pass == 1) bs -= m.ccL[c];
                        const double pinv = 1.0 / (s * z);
                        const double dz = m.dzL[c] - z * m.eL[c] * m.dx[c];
                        const double ds = (bs - s * dz) * (pinv * s);
                        m.dzL[c] = dz; m.dsL[c] = ds;
                        pdd += ds * dz; ts = fmax(ts, -ds * (pinv * z)); tz = fmax(tz, -dz * (pinv * s));
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
            SCP_TIMER(11)
        }
        CTA_PHASE(tid)
            for (int c = tid; c < n1p; c += cta.nt) {
                if (c < n1) m.x[c] += step * m.dx[c];
                if (ipm_has_ub(m, ctl, c)) { m.sU[c] += step * m.dsU[c]; m.zU[c] += step * m.dzU[c]; }
                if (ipm_has_lb(m, ctl, c)) { m.sL[c] += step * m.dsL[c]; m.zL[c] += step * m.dzL[c]; }
            }
            for (int r = tid; r < mc; r += cta.nt) { m.sA[r] += step * m.dsA[r]; m.zA[r] += step * m.dzA[r]; }
            for (int c = tid; c < m.nr; c += cta.nt) {
                m.sP[c] += step * m.dsP[c]; m.zP[c] += step * m.dzP[c];
                m.sM[c] += step * m.dsM[c]; m.zM[c] += step * m.dzM[c];
            }
        CTA_PHASE_END
    }
    if (*fixed_p) status |= SCPB200_ST_QP_PIVOT;
    res->fval = f0; res->gap = gap; res->relgap = relgap; res->pres = pres; res->dres = dres;
    res->iters = iters; res->status = status; res->snap_saved = snap_saved;
}
