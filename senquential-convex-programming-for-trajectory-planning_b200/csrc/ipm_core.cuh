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
// blocked right-looking Cholesky whose factor is then inverted in place (chol_tiles), so that every solve is two
// triangular mat-vecs.
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
    double *S;      // [T(T+1)/2 * 64]   tile-packed lower triangle of the normal matrix / its Cholesky factor;
                    //                   chol_tiles leaves X = L^-1 here
    double *x, *q, *rx, *dx, *tn;                    // [n1p]
    double *bA, *sA, *zA, *rzA, *dsA, *dzA, *ccA, *eA;   // [mc]    collision rows (e = 1/(s + delta z))
    double *ub, *sU, *zU, *dsU, *dzU, *ccU, *eU;         // [n1p]   x <= ub rows
    double *lb, *sL, *zL, *dsL, *dzL, *ccL, *eL;         // [n1p]   x >= lb rows
    double *dinv;   // [n1p]             reciprocal pivots of the factor
    double *wbuf;   // [max(T*64, 4*n1p)] tile scratch of the inversion sweep / partial sums of the solves
    double *red;    // [SCP_RED_DOUBLES] reduction scratch (double-buffered, see scp_common.cuh)
    double *t8;     // [16] tile-solve scratch (8) + flags
};

struct IpmResult {
    double fval, gap, relgap, pres, dres;
    int iters, status;
    int snap_saved;      // an iterate was written to ctl.snap during this solve
};

SCP_HDFN size_t ipm_snap_doubles(int n1p, int mc) { return (size_t)5 * n1p + 2 * (size_t)((mc + 1) & ~1); }

// Copy the interior iterate (x, sA, zA, sU, zU, sL, zL) to / from the snapshot area (one phase).
SCP_FN void ipm_snapshot(Cta &cta, const IpmMem &m, double *snap, bool save)
{
    const int n1p = m.n1p, mc = m.mc;
    const size_t rows = (size_t)((mc + 1) & ~1);
    double *px = snap, *psA = px + n1p, *pzA = psA + rows, *psU = pzA + rows, *pzU = psU + n1p, *psL = pzU + n1p,
           *pzL = psL + n1p;
    CTA_PHASE(tid)
        if (save) {
            for (int c = tid; c < n1p; c += cta.nt) {
                px[c] = m.x[c]; psU[c] = m.sU[c]; pzU[c] = m.zU[c]; psL[c] = m.sL[c]; pzL[c] = m.zL[c];
            }
            for (int r = tid; r < mc; r += cta.nt) { psA[r] = m.sA[r]; pzA[r] = m.zA[r]; }
        } else {
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

// out = sign * sum_{j < nprod} A_j B_j   with A_j = Abase + j*astride, B_j = Bbase + j*bstride (plain products)
SCP_FN void warp_tile_gemm_sum(int lane, double *out, const double *Abase, int astride, const double *Bbase, int bstride,
                               int nprod, double sign)
{
#if SCP_DEVICE_BUILD
    // two accumulator pairs (even / odd products): the DMMA chain of one tile would otherwise be 2 * nprod deep
    double n0 = 0.0, n1 = 0.0, m0 = 0.0, m1 = 0.0;
    const int ia0 = scp_frag_rowmajor(lane, 0), ia1 = scp_frag_rowmajor(lane, 1);
    const int ib0 = scp_frag_colmajor(lane, 0), ib1 = scp_frag_colmajor(lane, 1);
    int j = 0;
    for (; j + 1 < nprod; j += 2) {
        const double *A = Abase + (size_t)j * astride, *B = Bbase + (size_t)j * bstride;
        const double a0 = A[ia0], a1 = A[ia1], b0 = B[ib0], b1 = B[ib1];
        const double e0 = A[astride + ia0], e1 = A[astride + ia1], f0 = B[bstride + ib0], f1 = B[bstride + ib1];
        scp_dmma(n0, n1, a0, b0);
        scp_dmma(m0, m1, e0, f0);
        scp_dmma(n0, n1, a1, b1);
        scp_dmma(m0, m1, e1, f1);
    }
    if (j < nprod) {
        const double *A = Abase + (size_t)j * astride, *B = Bbase + (size_t)j * bstride;
        const double a0 = A[ia0], a1 = A[ia1], b0 = B[ib0], b1 = B[ib1];
        scp_dmma(n0, n1, a0, b0);
        scp_dmma(n0, n1, a1, b1);
    }
    n0 += m0; n1 += m1;
    double2 c;
    c.x = sign * n0; c.y = sign * n1;
    *reinterpret_cast<double2 *>(out + scp_frag_c(lane)) = c;
#else
    if (lane == 0) {
        double acc[64];
        for (int e = 0; e < 64; ++e) acc[e] = 0.0;
        for (int j = 0; j < nprod; ++j) {
            const double *A = Abase + (size_t)j * astride, *B = Bbase + (size_t)j * bstride;
            for (int r = 0; r < 8; ++r)
                for (int c = 0; c < 8; ++c)
                    for (int k = 0; k < 8; ++k) acc[r * 8 + c] += A[scp_tphys(r, k)] * B[scp_tphys(k, c)];
        }
        for (int r = 0; r < 8; ++r)
            for (int c = 0; c < 8; ++c) out[scp_tphys(r, c)] = sign * acc[r * 8 + c];
    }
#endif
}

// ------------------------------------------------------------------------------------------------ Cholesky
// In-place blocked right-looking Cholesky of the tile-packed lower triangle, followed by the in-place
// inversion of the factor: on exit m.S holds X = L^-1 (lower triangular, tile-packed).
//
// Why the inverse: forward/backward substitution is a dependent chain over the n1 unknowns (two synchronised
// steps per tile column, ~13k cycles per solve at n1 = 81 on B200) and the interior-point iteration needs two
// solves per factorisation, one after the other.  With X every solve is two triangular mat-vecs
// (S^-1 b = X'(X b)); the inversion costs n1^3/6 multiply-adds once per factorisation on the tensor path.
//   per tile column K:  (a) lane 0 factors the diagonal tile, (b) panel rows are solved against it by
//   substitution (one thread per row), (c) trailing tiles -= panel panel' (one warp per tile, DMMA).
//   then:               (d) all diagonal tiles are inverted in one phase (one thread per tile column),
//                       (e) for K = T-2 .. 0:  W = L[K+1:,K] X_KK ;  X[K+1:,K] = -X[K+1:,K+1:] W   (DMMA).
// *fixed is set if any pivot had to be repaired.
SCP_FN void chol_tiles(Cta &cta, const IpmMem &m, int *fixed SCP_TIMER_ARG)
{
    const int T = m.T;
    double *S = m.S, *dinv = m.dinv, *wbuf = m.wbuf;
    // Look-ahead: the factorisation of diagonal tile K+1 (a single-lane dependency chain, ~1.7 k cycles) runs in
    // warp 0 during the trailing update of step K, right after that warp has updated the tile; the other warps
    // carry the remaining tiles of the update.
    CTA_PHASE(tid)
        if (tid == 0) tile_potrf(S + scp_tile_off(0, 0), dinv, fixed);
    CTA_PHASE_END
    SCP_TIMER(2)
    for (int K = 0; K < T - 1; ++K) {
        const double *Lkk = S + scp_tile_off(K, K);
        const int Tr = T - K - 1;
        // (b) panel rows: x L_KK' = s  ->  x[c] = (s[c] - sum_{c2<c} x[c2] L[c][c2]) dinv[c]
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
                    double acc = x[c];
#pragma unroll
                    for (int c2 = 0; c2 < 8; ++c2)
                        if (c2 < c) acc -= x[c2] * Lkk[scp_tphys(c, c2)];
                    x[c] = acc * dinv[K * 8 + c];
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
    // (d) invert every diagonal tile: thread (K, j) computes column j of inv(L_KK) in registers ...
    CTA_PHASE(tid)
        for (int t0 = 0; t0 < T * 8; t0 += cta.nt) {
            const int t = t0 + tid;
            if (t < T * 8) {
                const int K = t >> 3, j = t & 7;
                double xcol[8];
                tile_trtri_column(S + scp_tile_off(K, K), dinv + K * 8, j, xcol);
#pragma unroll
                for (int i = 0; i < 8; ++i) wbuf[K * 64 + (i << 3) + ((((j >> 2) ^ (i >> 1)) & 1) << 2) + (j & 3)] = xcol[i];
            }
        }
    CTA_PHASE_END
    // ... and the tiles are replaced once every column has been read
    CTA_PHASE(tid)
        for (int e = tid; e < T * 64; e += cta.nt) S[scp_tile_off(e >> 6, e >> 6) + (e & 63)] = wbuf[e];
    CTA_PHASE_END
    SCP_TIMER(5)
    // (e) off-diagonal tiles of the inverse, column by column from the right
    for (int K = T - 2; K >= 0; --K) {
        const int Tr = T - K - 1;
        // W_I = L_IK X_KK   (I > K), one warp per panel tile
        WARP_SECTION(w, nw)
            WARP_PHASE(lane)
                for (int ii = w; ii < Tr; ii += nw)
                    warp_tile_gemm_sum(lane, wbuf + ii * 64, S + scp_tile_off(K + 1 + ii, K), 0, S + scp_tile_off(K, K), 0, 1, 1.0);
            WARP_PHASE_END
        WARP_SECTION_END
        CTA_SYNC
        SCP_TIMER(6)
        // X_IK = -sum_{J=K+1..I} X_IJ W_J: row ii (0-based below K) is a chain of ii + 1 tile products.  Rows are dealt
        // to the warps longest first in snake order (w = 0..nw-1, nw-1..0, ...), which keeps the longest warp within
        // one row of the mean for 4 and for 8 warps.
        WARP_SECTION(w, nw)
            WARP_PHASE(lane)
                for (int pass = 0; pass * nw < Tr; ++pass) {
                    const int q = pass * nw + ((pass & 1) ? nw - 1 - w : w);      // q-th longest row
                    if (q < Tr) {
                        const int ii = Tr - 1 - q, I = K + 1 + ii;
                        warp_tile_gemm_sum(lane, S + scp_tile_off(I, K), S + scp_tile_off(I, K + 1), SCP_TILE2, wbuf, SCP_TILE2,
                                           ii + 1, -1.0);
                    }
                }
            WARP_PHASE_END
        WARP_SECTION_END
        CTA_SYNC
        SCP_TIMER(7)
    }
}

// v := S^-1 v = X'(X v) with X = L^-1 left in m.S by chol_tiles (v has length n1p; `tmp` is n1p scratch).
// One thread per row (then per column) of X, four accumulators each.
SCP_FN void chol_solve_tiles(Cta &cta, const IpmMem &m, double *v, double *tmp)
{
    const int T = m.T, n1p = m.n1p;
    const double *S = m.S;
    CTA_PHASE(tid)                                  // tmp = X v
        for (int i = tid; i < n1p; i += cta.nt) {
            const int I = i >> 3, r = i & 7, h0 = ((r >> 1) & 1) << 2, h1 = h0 ^ 4;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int J = 0; J <= I; ++J) {
                const double *row = S + scp_tile_off(I, J) + (r << 3);
                const double *vj = v + J * 8;
                a0 += row[h0] * vj[0] + row[h0 + 1] * vj[1];
                a1 += row[h0 + 2] * vj[2] + row[h0 + 3] * vj[3];
                a2 += row[h1] * vj[4] + row[h1 + 1] * vj[5];
                a3 += row[h1 + 2] * vj[6] + row[h1 + 3] * vj[7];
            }
            tmp[i] = (a0 + a1) + (a2 + a3);
        }
    CTA_PHASE_END
    CTA_PHASE(tid)                                  // v = X' tmp
        for (int j = tid; j < n1p; j += cta.nt) {
            const int J = j >> 3, c = j & 7, lo = c & 3, he = (c >> 2) << 2, ho = he ^ 4;   // rows 0,1,4,5 / rows 2,3,6,7
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int I = J; I < T; ++I) {
                const double *tl = S + scp_tile_off(I, J);
                const double *ti = tmp + I * 8;
                a0 += tl[he + lo] * ti[0] + tl[8 + he + lo] * ti[1];
                a1 += tl[16 + ho + lo] * ti[2] + tl[24 + ho + lo] * ti[3];
                a2 += tl[32 + he + lo] * ti[4] + tl[40 + he + lo] * ti[5];
                a3 += tl[48 + ho + lo] * ti[6] + tl[56 + ho + lo] * ti[7];
            }
            v[j] = (a0 + a1) + (a2 + a3);
        }
    CTA_PHASE_END
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

// On entry m.q, m.bA, m.ub, m.lb hold the problem data (padding of q/ub/lb beyond n1 is ignored).
// On exit m.x holds the solution.
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
        op.form_normal(cta, m, m.dsA, m.tn, [](int) { return 1.0; },                         // dd = 1
                       [&](int c) { return (ipm_has_ub(m, ctl, c) ? 1.0 : 0.0) + (ipm_has_lb(m, ctl, c) ? 1.0 : 0.0); }
                       SCP_TIMER_PASS);
        op.prep(cta, (const double *)0, m.bA);
        CTA_PHASE(tid)
            for (int c = tid; c < n1; c += cta.nt) {
                double rhs = op.col_dot(c) - m.q[c];
                if (ipm_has_ub(m, ctl, c)) rhs += m.ub[c];
                if (ipm_has_lb(m, ctl, c)) rhs += m.lb[c];
                m.x[c] = rhs;
            }
        CTA_PHASE_END
        SCP_TIMER(1)
        chol_tiles(cta, m, fixed_p SCP_TIMER_PASS);
        chol_solve_tiles(cta, m, m.x, m.tn);
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

        // ---- normal matrix with dd = z e and its inverted factor -----------------------------------
        op.form_normal(cta, m, m.dsA, m.tn, [&](int r) { return m.zA[r] * m.eA[r]; },            // dd (kept in dsA)
                       [&](int c) {
                           double d = 0.0;
                           if (ipm_has_ub(m, ctl, c)) d += m.zU[c] * m.eU[c];
                           if (ipm_has_lb(m, ctl, c)) d += m.zL[c] * m.eL[c];
                           return d;
                       } SCP_TIMER_PASS);
        SCP_TIMER(1)
        chol_tiles(cta, m, fixed_p SCP_TIMER_PASS);

        const double mu = gap / mrows;
        double sigma = 0.0, step = 1.0;
        for (int pass = 0; pass < 2; ++pass) {
            const double smu = sigma * mu;
            // w1 for the collision rows (input of A'w1); bs = -s z + sigma mu - [pass 1] (ds_a dz_a)
            CTA_PHASE(tid)
                for (int r = tid; r < mc; r += cta.nt) {
                    const double s = m.sA[r], z = m.zA[r];
                    double bs = smu - s * z;
                    if (pass == 1) bs -= m.ccA[r];
                    m.dzA[r] = ipm_w1(s, z, m.rzA[r], m.eA[r], bs);                       // w1 in dzA
                }
            CTA_PHASE_END
            op.prep(cta, (const double *)0, m.dzA);
            // rhs = -rx - G'w1
            CTA_PHASE(tid)
                for (int c = tid; c < n1p; c += cta.nt) {
                    double rhs = 0.0;
                    if (c < n1) rhs = -m.rx[c] - op.col_dot(c);
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
            SCP_TIMER(10)
            chol_solve_tiles(cta, m, m.dx, m.tn);
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
        CTA_PHASE_END
    }
    if (*fixed_p) status |= SCPB200_ST_QP_PIVOT;
    res->fval = f0; res->gap = gap; res->relgap = relgap; res->pres = pres; res->dres = dres;
    res->iters = iters; res->status = status; res->snap_saved = snap_saved;
}
