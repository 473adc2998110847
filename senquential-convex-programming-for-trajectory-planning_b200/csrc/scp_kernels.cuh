// scp_kernels.cuh — per-instance device code of the SCP-QP path: set-up (K1), dense assembly (K2), QCQP
// evaluation, forward prediction and the fused SCP loop (K4).  Written in the phase style of scp_common.cuh so
// that tests/emu can run the same source on the host.
#pragma once
#include "ipm_core.cuh"
#include "ops_dense.cuh"
#include "ops_pair.cuh"

// ================================================================================================ Philox
// Philox4x32-10 (Salmon et al., SC'11) + Box-Muller.  Replaces the np.random.normal draws of Model.py:84-86,
// whose global MT19937 stream cannot be reproduced in a batched kernel; keyed on (seed; instance, vehicle,
// counter) so results do not depend on how instances are sharded over GPUs.
SCP_HDFN void scp_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Every consumer of noise inside one MPC step draws from its OWN stream (the fourth counter word): the set-up (K1),
// the plant step, the delay-compensation prediction and the tick-path predictions all run with the same
// noise_counter (the MPC step index), and the reference's np.random.normal draws are independent between them.
#define SCP_NOISE_SETUP 0u      /* K1: Ec = f(x,u) + noise - ... (Model.py:58 through :84-86) */
#define SCP_NOISE_PLANT 1u      /* plant integration, main.py:185-190 */
#define SCP_NOISE_ODE 2u        /* scpb200_ode_predict: 2 + params.noise_stream (0 = delay compensation, MPC_Iter.py:25-33) */
SCP_HDFN void scp_noise_pair(uint64_t seed, uint32_t instance, uint32_t vehicle, uint32_t counter, double out[2],
                             uint32_t stream = SCP_NOISE_SETUP)
{
    uint32_t r[4];
    scp_philox4x32_10(instance, vehicle, counter, 0x5C9B200u + stream, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u1 = ((double)(((uint64_t)r[0] << 21) ^ (r[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
    const double u2 = (double)(((uint64_t)r[2] << 21) ^ (r[3] >> 11)) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1)), ang = 6.283185307179586476925286766559 * u2;
    out[0] = rad * cos(ang);
    out[1] = rad * sin(ang);
}

// ================================================================================================ K1 pieces
// Model.py:61-87
SCP_HDFN void scp_bicycle_rhs(const double x[6], double u_ref, double Lf, double Lr, double dx[6])
{
    const double L = Lf + Lr, R = Lr / L;
    const double tu = tan(x[5]);
    const double v_center = x[3] * sqrt(1.0 + (R * tu) * (R * tu));
    const double beta = atan(R * tu);
    dx[0] = v_center * cos(x[2] + beta);
    dx[1] = v_center * sin(x[2] + beta);
    dx[2] = v_center * tu * cos(beta) / L;
    dx[3] = x[4];
    dx[4] = 0.0;
    dx[5] = (u_ref - x[5]) / 0.1;
}

// 8x8 matrix product C = A B (row-major, thread-local)
SCP_HDFN void scp_mm8(const double *A, const double *B, double *C)
{
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 8; ++k) acc += A[i * 8 + k] * B[k * 8 + j];
            C[i * 8 + j] = acc;
        }
}

// E = expm(M), 8x8: degree-13 Pade approximant with scaling and squaring (what scipy.linalg.expm runs at
// MPC_Iter.py:106,111).  Returns 0, or -1 if the Pade denominator is singular.
SCP_HDFN int scp_expm8(const double *Min, double *E)
{
    const double b[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                          129060195264000.,   10559470521600.,    670442572800.,     33522128640.,
                          1323241920.,        40840800.,          960960.,           16380., 182., 1.};
    double A[64], A2[64], A4[64], A6[64], U[64], V[64], W[64];
    double nrm = 0.0;
    for (int j = 0; j < 8; ++j) {
        double c = 0.0;
        for (int i = 0; i < 8; ++i) c += fabs(Min[i * 8 + j]);
        nrm = fmax(nrm, c);
    }
    int s = 0;
    if (nrm > 5.371920351148152) {
        s = (int)ceil(log2(nrm / 5.371920351148152));
        if (s < 0) s = 0;
    }
    const double sc = ldexp(1.0, -s);
    for (int i = 0; i < 64; ++i) A[i] = Min[i] * sc;
    scp_mm8(A, A, A2);
    scp_mm8(A2, A2, A4);
    scp_mm8(A4, A2, A6);
    for (int i = 0; i < 64; ++i) W[i] = b[13] * A6[i] + b[11] * A4[i] + b[9] * A2[i];
    scp_mm8(A6, W, V);                                   // V used as scratch
    for (int i = 0; i < 64; ++i) V[i] += b[7] * A6[i] + b[5] * A4[i] + b[3] * A2[i];
    for (int i = 0; i < 8; ++i) V[i * 8 + i] += b[1];
    scp_mm8(A, V, U);
    for (int i = 0; i < 64; ++i) W[i] = b[12] * A6[i] + b[10] * A4[i] + b[8] * A2[i];
    scp_mm8(A6, W, V);
    for (int i = 0; i < 64; ++i) V[i] += b[6] * A6[i] + b[4] * A4[i] + b[2] * A2[i];
    for (int i = 0; i < 8; ++i) V[i * 8 + i] += b[0];
    // solve (V - U) X = (V + U): Gaussian elimination with partial pivoting; Q in A2, R in W
    for (int i = 0; i < 64; ++i) { A2[i] = V[i] - U[i]; W[i] = V[i] + U[i]; }
    for (int k = 0; k < 8; ++k) {
        int piv = k;
        double best = fabs(A2[k * 8 + k]);
        for (int i = k + 1; i < 8; ++i)
            if (fabs(A2[i * 8 + k]) > best) { best = fabs(A2[i * 8 + k]); piv = i; }
        if (best == 0.0) return -1;
        if (piv != k)
            for (int j = 0; j < 8; ++j) {
                double t = A2[k * 8 + j]; A2[k * 8 + j] = A2[piv * 8 + j]; A2[piv * 8 + j] = t;
                t = W[k * 8 + j]; W[k * 8 + j] = W[piv * 8 + j]; W[piv * 8 + j] = t;
            }
        for (int i = k + 1; i < 8; ++i) {
            const double l = A2[i * 8 + k] / A2[k * 8 + k];
            if (l == 0.0) continue;
            for (int j = k; j < 8; ++j) A2[i * 8 + j] -= l * A2[k * 8 + j];
            for (int j = 0; j < 8; ++j) W[i * 8 + j] -= l * W[k * 8 + j];
        }
    }
    for (int j = 0; j < 8; ++j)
        for (int i = 7; i >= 0; --i) {
            double acc = W[i * 8 + j];
            for (int k = i + 1; k < 8; ++k) acc -= A2[i * 8 + k] * W[k * 8 + j];
            W[i * 8 + j] = acc / A2[i * 8 + i];
        }
    for (int k = 0; k < s; ++k) {
        scp_mm8(W, W, V);
        for (int i = 0; i < 64; ++i) W[i] = V[i];
    }
    for (int i = 0; i < 64; ++i) E[i] = W[i];
    return 0;
}

// SampleReferTraj.py:81-122
SCP_HDFN void scp_projection2d(double x1, double y1, double x2, double y2, double x3, double y3, double *xp, double *yp,
                               double *dist, double *lambda)
{
    const double b = sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
    if (b != 0.0) {
        const double xn = (x2 - x1) / b, yn = (y2 - y1) / b, x31 = x3 - x1, y31 = y3 - y1;
        const double dot = xn * x31 + yn * y31;
        *dist = xn * y31 - yn * x31;
        *xp = x1 + dot * xn;
        *yp = y1 + dot * yn;
        *lambda = dot / b;
    } else {
        *dist = sqrt((x3 - x1) * (x3 - x1) + (y3 - y1) * (y3 - y1));
        *lambda = 0.0;
        *xp = x1;
        *yp = y1;
    }
}

// SampleReferTraj.py:8-79: projection onto the polyline (first/last pieces extended), then nSamples points at
// spacing `step`; the trajectory index is never advanced (:26-28), which reproduces the reference's
// end-of-polyline oscillation.  Returns -1 where the reference would raise IndexError (index_min == nPts).
SCP_HDFN int scp_sample_reference(int nSamples, int npts, const double *poly, double vx, double vy, double step,
                                  double *out)
{
    double cx = poly[2], cy = poly[3];
    double sd_min = sqrt((vx - poly[2]) * (vx - poly[2]) + (vy - poly[3]) * (vy - poly[3]));
    int idx = 2;
    for (int j = 1; j < npts; ++j) {
        double xp, yp, sd, lam;
        scp_projection2d(poly[(j - 1) * 2], poly[(j - 1) * 2 + 1], poly[j * 2], poly[j * 2 + 1], vx, vy, &xp, &yp, &sd,
                         &lam);
        if ((0.0 < lam || j == 1) && (lam < 1.0 || j == npts - 1)) {
            if (fabs(sd) < fabs(sd_min)) { cx = xp; cy = yp; sd_min = sd; idx = j; }
        } else {   // :69-76 (the reference's '^' at :70 is read as the square it was meant to be)
            const double ex = vx - poly[j * 2], ey = vy - poly[j * 2 + 1];
            const double d_end = sqrt(ex * ex + ey * ey);
            if (fabs(d_end) < fabs(sd_min)) {
                cx = poly[j * 2]; cy = poly[j * 2 + 1];
                sd_min = (sd > 0.0 ? 1.0 : (sd < 0.0 ? -1.0 : 0.0)) * d_end;
                idx = j;
            }
        }
    }
    if (idx >= npts) return -1;
    const double dx = poly[idx * 2] - poly[(idx - 1) * 2], dy = poly[idx * 2 + 1] - poly[(idx - 1) * 2 + 1];
    const double nrm = sqrt(dx * dx + dy * dy);
    for (int i = 0; i < nSamples; ++i) {
        const double rx = cx - poly[idx * 2], ry = cy - poly[idx * 2 + 1];
        const double remaining = sqrt(rx * rx + ry * ry);
        if (remaining > step) {
            cx = cx + step * (dx / nrm);
            cy = cy + step * (dy / nrm);
        } else {
            cx = poly[idx * 2] + (step - remaining) * (dx / nrm);
            cy = poly[idx * 2 + 1] + (step - remaining) * (dy / nrm);
        }
        out[i * 2] = cx;
        out[i * 2 + 1] = cy;
    }
    return 0;
}

// One (instance, vehicle): Model.py:45-59 (Jacobian), MPC_Iter.py:99-113 (ZOH via expm; both of the
// reference's 7x7 exponentials come out of one 8x8 exponential of dt*[[Ac,Bc,Ec],[0,0,0]]), :129-149
// (prediction recurrences), :35-43 (reference sampling).  Returns 0 or SCPB200_ST_SETUP.
SCP_HDFN int scp_setup_vehicle(int Hp, int nPts, double dt, const double *x, double u0, const double *pv,
                               const double *poly, const double *noise, double *ref, double *g, double *cterm,
                               double *abe)
{
    const double Lf = pv[0], Lr = pv[1], Ls = Lf + Lr;
    const double t5 = tan(x[5]);
    const double w = sqrt((Lr * Lr * t5 * t5) / (Ls * Ls) + 1.0);
    const double ang = x[2] + atan((Lr * t5) / Ls);
    const double sa = sin(ang), ca = cos(ang), sec2 = t5 * t5 + 1.0;
    double M[64], E[64];
    for (int i = 0; i < 64; ++i) M[i] = 0.0;
    M[0 * 8 + 2] = -x[3] * sa * w;
    M[0 * 8 + 3] = ca * w;
    M[0 * 8 + 5] = (Lr * Lr * x[3] * ca * t5 * sec2) / (w * Ls * Ls) - (Lr * x[3] * sa * sec2) / (w * Ls);
    M[1 * 8 + 2] = x[3] * ca * w;
    M[1 * 8 + 3] = sa * w;
    M[1 * 8 + 5] = (Lr * x[3] * ca * sec2) / (w * Ls) + (Lr * Lr * x[3] * sa * t5 * sec2) / (w * Ls * Ls);
    M[2 * 8 + 3] = t5 / Ls;
    M[2 * 8 + 5] = (x[3] * sec2) / Ls;
    M[3 * 8 + 4] = 1.0;
    M[5 * 8 + 5] = -10.0;
    M[5 * 8 + 6] = 10.0;                                    // Bc
    double f[6];
    scp_bicycle_rhs(x, u0, Lf, Lr, f);
    if (noise) { f[0] += noise[0]; f[1] += noise[1]; }
    for (int i = 0; i < 6; ++i) {                           // Ec = f - Ac x - Bc u  (Model.py:58)
        double acc = f[i];
        for (int j = 0; j < 6; ++j) acc -= M[i * 8 + j] * x[j];
        M[i * 8 + 7] = acc - M[i * 8 + 6] * u0;
    }
    for (int i = 0; i < 64; ++i) M[i] *= dt;
    int rc = 0;
    if (scp_expm8(M, E)) rc = SCPB200_ST_SETUP;
    double Ad[36], Bd[6], Ed[6];
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) Ad[i * 6 + j] = E[i * 8 + j];
        Bd[i] = E[i * 8 + 6];
        Ed[i] = fabs(E[i * 8 + 7]) <= 1e-30 ? 0.0 : E[i * 8 + 7];      // MPC_Iter.py:87
    }
    if (abe) {
        for (int i = 0; i < 36; ++i) abe[i] = Ad[i];
        for (int i = 0; i < 6; ++i) { abe[36 + i] = Bd[i]; abe[42 + i] = Ed[i]; }
    }
    if (scp_sample_reference(Hp, nPts, poly, x[0], x[1], x[3] * dt, ref)) rc = SCPB200_ST_SETUP;
    double CA[12], Sm[12], Tm[12];
    for (int i = 0; i < 12; ++i) { CA[i] = 0.0; Sm[i] = 0.0; }
    CA[0] = 1.0;
    CA[7] = 1.0;
    for (int k = 0; k < Hp; ++k) {
        for (int r = 0; r < 2; ++r) {
            double acc = 0.0;
            for (int j = 0; j < 6; ++j) acc += CA[r * 6 + j] * Bd[j];
            g[k * 2 + r] = acc;
        }
        for (int i = 0; i < 12; ++i) Sm[i] += CA[i];
        for (int r = 0; r < 2; ++r)
            for (int j = 0; j < 6; ++j) {
                double acc = 0.0;
                for (int l = 0; l < 6; ++l) acc += CA[r * 6 + l] * Ad[l * 6 + j];
                Tm[r * 6 + j] = acc;
            }
        for (int i = 0; i < 12; ++i) CA[i] = Tm[i];
        for (int r = 0; r < 2; ++r) {
            double acc = 0.0, acc2 = 0.0;
            for (int j = 0; j < 6; ++j) { acc += CA[r * 6 + j] * x[j]; acc2 += Sm[r * 6 + j] * Ed[j]; }
            cterm[k * 2 + r] = acc + acc2;
        }
    }
    return rc;
}

// ---- K1 on a warp (device): one warp per (instance, vehicle), 8x8 matrices row-major in shared memory, every 8x8x8
// product as two FP64 mma.m8n8k4 (the contraction north-star item (1) names), no thread-local matrix arrays.
// Lane l owns elements 2l, 2l+1 of a row-major 8x8 matrix — exactly its accumulator fragment, so element-wise steps
// follow a product without an exchange.
#define SCP_K1_WARP_DOUBLES 512      /* shared scratch per warp: 7 matrices (448) + recurrence vectors (64) */
#if SCP_DEVICE_BUILD
SCP_FN void warp_mm8(int lane, const double *X, const double *Y, double *C)      // C = X Y (C aliases neither)
{
    const int r = lane >> 2, q = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    const double x0 = X[r * 8 + q], x1 = X[r * 8 + q + 4], y0 = Y[q * 8 + r], y1 = Y[(q + 4) * 8 + r];
    scp_dmma(c0, c1, x0, y0);
    scp_dmma(c0, c1, x1, y1);
    double2 c;
    c.x = c0; c.y = c1;
    reinterpret_cast<double2 *>(C)[lane] = c;
    __syncwarp();
}

// W := expm(A) (A is scaled in place), degree-13 Pade with scaling and squaring as scp_expm8; ws = 6 matrices.
// Returns 0, or -1 if the Pade denominator is singular (warp-uniform).
SCP_FN int warp_expm8(int lane, double *A, double *ws)
{
    const double b[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                          129060195264000.,   10559470521600.,    670442572800.,     33522128640.,
                          1323241920.,        40840800.,          960960.,           16380., 182., 1.};
    double *A2 = ws, *A4 = ws + 64, *A6 = ws + 128, *U = ws + 192, *V = ws + 256, *W = ws + 320;
    const int e0 = 2 * lane, e1 = e0 + 1;
    const bool d0 = (e0 >> 3) == (e0 & 7), d1 = (e1 >> 3) == (e1 & 7);
    double cs = 0.0;
    if (lane < 8)
        for (int i = 0; i < 8; ++i) cs += fabs(A[i * 8 + lane]);
    const double nrm = scp_warp_max(cs);
    int s = 0;
    if (nrm > 5.371920351148152) {
        s = (int)ceil(log2(nrm / 5.371920351148152));
        if (s < 0) s = 0;
    }
    const double sc = ldexp(1.0, -s);
    __syncwarp();
    A[e0] *= sc; A[e1] *= sc;
    __syncwarp();
    warp_mm8(lane, A, A, A2);
    warp_mm8(lane, A2, A2, A4);
    warp_mm8(lane, A4, A2, A6);
    W[e0] = b[13] * A6[e0] + b[11] * A4[e0] + b[9] * A2[e0];
    W[e1] = b[13] * A6[e1] + b[11] * A4[e1] + b[9] * A2[e1];
    __syncwarp();
    warp_mm8(lane, A6, W, V);
    V[e0] += b[7] * A6[e0] + b[5] * A4[e0] + b[3] * A2[e0]; if (d0) V[e0] += b[1];
    V[e1] += b[7] * A6[e1] + b[5] * A4[e1] + b[3] * A2[e1]; if (d1) V[e1] += b[1];
    __syncwarp();
    warp_mm8(lane, A, V, U);
    W[e0] = b[12] * A6[e0] + b[10] * A4[e0] + b[8] * A2[e0];
    W[e1] = b[12] * A6[e1] + b[10] * A4[e1] + b[8] * A2[e1];
    __syncwarp();
    warp_mm8(lane, A6, W, V);
    V[e0] += b[6] * A6[e0] + b[4] * A4[e0] + b[2] * A2[e0]; if (d0) V[e0] += b[0];
    V[e1] += b[6] * A6[e1] + b[4] * A4[e1] + b[2] * A2[e1]; if (d1) V[e1] += b[0];
    __syncwarp();
    // solve (V - U) X = (V + U): Gaussian elimination with partial pivoting on [Q | R], Q in A2, R in W; lane c < 16 owns
    // column c of the 8 x 16 array, the pivot search is done by every lane alike
    {
        const double q0 = V[e0] - U[e0], q1 = V[e1] - U[e1], r0 = V[e0] + U[e0], r1 = V[e1] + U[e1];
        A2[e0] = q0; A2[e1] = q1; W[e0] = r0; W[e1] = r1;
    }
    __syncwarp();
    double *col = lane < 8 ? A2 + lane : W + (lane - 8);
    int fail = 0;
    for (int k = 0; k < 8; ++k) {
        int piv = k;
        double best = fabs(A2[k * 8 + k]);
        for (int i = k + 1; i < 8; ++i) {
            const double t = fabs(A2[i * 8 + k]);
            if (t > best) { best = t; piv = i; }
        }
        if (best == 0.0) { fail = 1; break; }
        double li[8];
        const double pk = A2[piv * 8 + k];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int src = i == k ? piv : (i == piv ? k : i);          // row i after the swap
            li[i] = A2[src * 8 + k] / pk;
        }
        __syncwarp();
        if (lane < 16 && (lane > k || lane >= 8)) {
            if (piv != k) { const double t = col[k * 8]; col[k * 8] = col[piv * 8]; col[piv * 8] = t; }
            const double ck = col[k * 8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i > k) col[i * 8] -= li[i] * ck;
        } else if (lane <= k && piv != k) {
            const double t = col[k * 8]; col[k * 8] = col[piv * 8]; col[piv * 8] = t;      // columns already eliminated: swap only
        }
        __syncwarp();
    }
    if (fail) return -1;
    if (lane < 8) {
        double *rc = W + lane;
        for (int i = 7; i >= 0; --i) {
            double acc = rc[i * 8];
            for (int k = i + 1; k < 8; ++k) acc -= A2[i * 8 + k] * rc[k * 8];
            rc[i * 8] = acc / A2[i * 8 + i];
        }
    }
    __syncwarp();
    for (int k = 0; k < s; ++k) {
        warp_mm8(lane, W, W, V);
        W[e0] = V[e0]; W[e1] = V[e1];
        __syncwarp();
    }
    return 0;
}

// scp_setup_vehicle by one warp.  ws: SCP_K1_WARP_DOUBLES of shared scratch.  x (6), pv (5), poly read from anywhere.
SCP_FN int warp_setup_vehicle(int lane, int Hp, int nPts, double dt, const double *x, double u0, const double *pv,
                              const double *poly, const double *noise, double *ref, double *g, double *cterm, double *abe,
                              double *ws)
{
    double *M = ws, *mats = ws + 64, *E = ws + 64 + 320, *vec = ws + 448;     // vec: CA[12] SM[12] CB[12] (double-buffered CA)
    M[2 * lane] = 0.0; M[2 * lane + 1] = 0.0;
    __syncwarp();
    if (lane == 0) {
        const double Lf = pv[0], Lr = pv[1], Ls = Lf + Lr;
        const double t5 = tan(x[5]);
        const double w = sqrt((Lr * Lr * t5 * t5) / (Ls * Ls) + 1.0);
        const double ang = x[2] + atan((Lr * t5) / Ls);
        const double sa = sin(ang), ca = cos(ang), sec2 = t5 * t5 + 1.0;
        M[0 * 8 + 2] = -x[3] * sa * w;
        M[0 * 8 + 3] = ca * w;
        M[0 * 8 + 5] = (Lr * Lr * x[3] * ca * t5 * sec2) / (w * Ls * Ls) - (Lr * x[3] * sa * sec2) / (w * Ls);
        M[1 * 8 + 2] = x[3] * ca * w;
        M[1 * 8 + 3] = sa * w;
        M[1 * 8 + 5] = (Lr * x[3] * ca * sec2) / (w * Ls) + (Lr * Lr * x[3] * sa * t5 * sec2) / (w * Ls * Ls);
        M[2 * 8 + 3] = t5 / Ls;
        M[2 * 8 + 5] = (x[3] * sec2) / Ls;
        M[3 * 8 + 4] = 1.0;
        M[5 * 8 + 5] = -10.0;
        M[5 * 8 + 6] = 10.0;                                    // Bc
        double f[6];
        scp_bicycle_rhs(x, u0, Lf, Lr, f);
        if (noise) { f[0] += noise[0]; f[1] += noise[1]; }
        for (int i = 0; i < 6; ++i) {                           // Ec = f - Ac x - Bc u  (Model.py:58)
            double acc = f[i];
            for (int j = 0; j < 6; ++j) acc -= M[i * 8 + j] * x[j];
            M[i * 8 + 7] = acc - M[i * 8 + 6] * u0;
        }
    }
    int rc = 0;
    // the sampler does not depend on the exponential: one lane of the otherwise idle half
    if (lane == 31 && scp_sample_reference(Hp, nPts, poly, x[0], x[1], x[3] * dt, ref)) rc = SCPB200_ST_SETUP;
    __syncwarp();
    M[2 * lane] *= dt; M[2 * lane + 1] *= dt;
    __syncwarp();
    if (warp_expm8(lane, M, mats)) rc = SCPB200_ST_SETUP;
    // E (row-major 8x8): Ad = E[0:6,0:6], Bd = E[0:6,6], Ed = E[0:6,7] with |Ed| <= 1e-30 -> 0 (MPC_Iter.py:87)
    if (lane < 6) {
        const double ed = E[lane * 8 + 7];
        if (fabs(ed) <= 1e-30) E[lane * 8 + 7] = 0.0;
    }
    __syncwarp();
    if (abe) {
        for (int i = lane; i < 36; i += 32) abe[i] = E[(i / 6) * 8 + (i % 6)];
        if (lane < 6) { abe[36 + lane] = E[lane * 8 + 6]; abe[42 + lane] = E[lane * 8 + 7]; }
    }
    // recurrences (MPC_Iter.py:129-149): CA_{i+1} = CA_i Ad, g_i = CA_i Bd, c(k) = CA_{k+1} x0 + (sum_{l<=k} CA_l) Ed; lanes
    // 0..11 own the entries of CA / its running sum, lanes 12, 13 the two components of g, lanes 14, 15 those of c
    double *CA = vec, *SM = vec + 12, *CB = vec + 24;
    const int r = lane < 12 ? lane / 6 : (lane & 1), j = lane < 12 ? lane - 6 * (lane / 6) : 0;
    double ca = 0.0, sm = 0.0;
    if (lane < 12) { ca = (lane == 0 || lane == 7) ? 1.0 : 0.0; CA[lane] = ca; }
    __syncwarp();
    for (int k = 0; k < Hp; ++k) {
        double *cur = (k & 1) ? CB : CA, *nxt = (k & 1) ? CA : CB;
        if (lane < 12) {
            sm += ca;
            SM[lane] = sm;
            double acc = 0.0;
            for (int l = 0; l < 6; ++l) acc += cur[r * 6 + l] * E[l * 8 + j];
            ca = acc;
            nxt[lane] = ca;
        } else if (lane < 14) {
            double acc = 0.0;
            for (int jj = 0; jj < 6; ++jj) acc += cur[r * 6 + jj] * E[jj * 8 + 6];
            g[k * 2 + r] = acc;
        }
        __syncwarp();
        if (lane == 14 || lane == 15) {
            double acc = 0.0, acc2 = 0.0;
            for (int jj = 0; jj < 6; ++jj) { acc += nxt[r * 6 + jj] * x[jj]; acc2 += SM[r * 6 + jj] * E[jj * 8 + 7]; }
            cterm[k * 2 + r] = acc + acc2;
        }
        __syncwarp();
    }
    return __any_sync(0xffffffffu, rc != 0) ? SCPB200_ST_SETUP : 0;
}
#endif

// K1, one CTA per instance: phase 1 one thread per vehicle (serial expm/recurrences), phase 2 all threads on
// the cost matrices of MPC_Iter.py:116-127:  H = B'QB + R,  qv = -2 B'Q(Ref - c),  gamma0 = sum_v Err'Q Err.
// x0b / u0b: the instance's own states ([nVeh][6], [nVeh]); every other array is a batch base indexed with b.
// coh: the instance's arrays may have been written by another CTA during this launch (rollout entry): read them past L1.
SCP_FN void scp_setup_instance_at(Cta &cta, const scpb200_dims &d, const scpb200_params &p, int b, const double *x0b,
                                  const double *u0b, const double *veh, const double *poly, double *ref, double *g,
                                  double *cterm, double *H, double *qv, double *gamma0, double *abe,
                                  int32_t *setup_status, double *red, int *flag, bool coh, double *k1ws, int k1_warps = 1 << 30)
{
    const int nVeh = d.nVeh, Hp = d.Hp, nPts = d.nPts;
    CTA_PHASE(tid)
        if (tid == 0) *flag = 0;
    CTA_PHASE_END
#if SCP_DEVICE_BUILD
    // one warp per vehicle (warp_setup_vehicle), scratch SCP_K1_WARP_DOUBLES per warp
    {
        const int w = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31, nw = scp_imin(cta.nt >> 5, k1_warps);
        for (int v = w; v < nVeh && w < nw; v += nw) {
            const size_t iv = (size_t)b * nVeh + v;
            double nz[2];
            const double *noise = 0;
            if (p.noise_sigma > 0.0) {
                scp_noise_pair(p.seed, p.instance0 + (uint32_t)b, (uint32_t)v, p.noise_counter, nz);
                nz[0] *= p.noise_sigma;
                nz[1] *= p.noise_sigma;
                noise = nz;
            }
            const int rc = warp_setup_vehicle(lane, Hp, nPts, p.dt, x0b + v * 6, u0b[v], veh + iv * 5, poly + iv * nPts * 2, noise,
                                              ref + iv * Hp * 2, g + iv * Hp * 2, cterm + iv * Hp * 2, abe ? abe + iv * 48 : 0,
                                              k1ws + (size_t)w * SCP_K1_WARP_DOUBLES);
            if (rc && lane == 0) *flag = rc;
        }
    }
    __syncthreads();
#else
    (void)k1ws;
    CTA_PHASE(tid)
        for (int v = tid; v < nVeh; v += cta.nt) {
            const size_t iv = (size_t)b * nVeh + v;
            double nz[2];
            const double *noise = 0;
            if (p.noise_sigma > 0.0) {
                scp_noise_pair(p.seed, p.instance0 + (uint32_t)b, (uint32_t)v, p.noise_counter, nz);
                nz[0] *= p.noise_sigma;
                nz[1] *= p.noise_sigma;
                noise = nz;
            }
            const int rc = scp_setup_vehicle(Hp, nPts, p.dt, x0b + v * 6, u0b[v], veh + iv * 5, poly + iv * nPts * 2,
                                             noise, ref + iv * Hp * 2, g + iv * Hp * 2, cterm + iv * Hp * 2,
                                             abe ? abe + iv * 48 : 0);
            if (rc) *flag = rc;
        }
    CTA_PHASE_END
#endif
    CTA_RED_BEGIN(cta, 1)
    CTA_PHASE(tid)
        double gam = 0.0;
        for (int e = tid; e < nVeh * Hp * Hp; e += cta.nt) {
            const int v = e / (Hp * Hp), a = (e / Hp) % Hp, bb = e % Hp;
            const size_t iv = (size_t)b * nVeh + v;
            const double *gv = g + iv * Hp * 2;
            const double Q = veh[iv * 5 + 2], Qf = veh[iv * 5 + 3], R = veh[iv * 5 + 4];
            double acc = 0.0;
            for (int i = (a > bb ? a : bb); i < Hp; ++i) {
                const double wq = (i == Hp - 1) ? Qf : Q;
                acc += wq * (scp_ldc(gv + (i - a) * 2, coh) * scp_ldc(gv + (i - bb) * 2, coh) +
                             scp_ldc(gv + (i - a) * 2 + 1, coh) * scp_ldc(gv + (i - bb) * 2 + 1, coh));
            }
            H[iv * Hp * Hp + a * Hp + bb] = acc + (a == bb ? R : 0.0);
        }
        for (int c = tid; c < nVeh * Hp; c += cta.nt) {
            const int v = c / Hp, a = c - v * Hp;
            const size_t iv = (size_t)b * nVeh + v;
            const double *gv = g + iv * Hp * 2, *cv = cterm + iv * Hp * 2, *rv = ref + iv * Hp * 2;
            const double Q = veh[iv * 5 + 2], Qf = veh[iv * 5 + 3];
            double acc = 0.0;
            for (int i = a; i < Hp; ++i) {
                const double wq = (i == Hp - 1) ? Qf : Q;
                acc += wq * (scp_ldc(gv + (i - a) * 2, coh) * (scp_ldc(rv + i * 2, coh) - scp_ldc(cv + i * 2, coh)) +
                             scp_ldc(gv + (i - a) * 2 + 1, coh) * (scp_ldc(rv + i * 2 + 1, coh) - scp_ldc(cv + i * 2 + 1, coh)));
            }
            qv[iv * Hp + a] = -2.0 * acc;
            const double wq = (a == Hp - 1) ? Qf : Q;
            const double ex = scp_ldc(rv + a * 2, coh) - scp_ldc(cv + a * 2, coh), ey = scp_ldc(rv + a * 2 + 1, coh) - scp_ldc(cv + a * 2 + 1, coh);
            gam += wq * (ex * ex + ey * ey);
        }
        CTA_RED_SUM(cta, red, 0, tid, gam)
    CTA_PHASE_END_RED(cta, red, 1)
    const double gam = cta_red_sum(cta, red, 0);
    CTA_PHASE(tid)
        if (tid == 0) {
            gamma0[b] = gam;
            if (setup_status) setup_status[b] = *flag;
        }
    CTA_PHASE_END
}

SCP_FN void scp_setup_instance(Cta &cta, const scpb200_dims &d, const scpb200_params &p, int b, const double *x0,
                               const double *u0, const double *veh, const double *poly, double *ref, double *g,
                               double *cterm, double *H, double *qv, double *gamma0, double *abe,
                               int32_t *setup_status, double *red, int *flag, double *k1ws = 0)
{
    scp_setup_instance_at(cta, d, p, b, x0 + (size_t)b * d.nVeh * 6, u0 + (size_t)b * d.nVeh, veh, poly, ref, g, cterm, H, qv,
                          gamma0, abe, setup_status, red, flag, false, k1ws);
}

// ================================================================================================ ODE prediction
// Classical RK4 of Model.py:61-87 with constant steering reference over `T` seconds, `nsub` substeps per output
// interval, writing `steps` samples (t = 0 .. T inclusive) — the delay-compensation prediction of IterClass
// (MPC_Iter.py:25-33: odeint over linspace(0, delay_x+dt+delay_u, 10)).  With noise_sigma > 0 every RHS
// evaluation adds N(0, sigma) to (dx, dy) (Model.py:84-86) from the keyed Philox stream
// (counter = noise_counter * 65536 + stage index, stream = the consumer's tag, see scp_noise_pair).
SCP_HDFN void scp_ode_predict_vehicle(const double *x_in, double u_ref, double Lf, double Lr, double T, int steps,
                                      int nsub, double noise_sigma, uint64_t seed, uint32_t instance, uint32_t vehicle,
                                      uint32_t noise_counter, double *out /*[steps][6]*/, uint32_t noise_stream = SCP_NOISE_ODE)
{
    double x[6], k1[6], k2[6], k3[6], k4[6], xt[6];
    for (int i = 0; i < 6; ++i) { x[i] = x_in[i]; out[i] = x_in[i]; }
    const double h = T / (double)((steps - 1) * nsub);
    uint32_t stage = 0;
    for (int s = 1; s < steps; ++s) {
        for (int j = 0; j < nsub; ++j) {
            double nz[2];
#define SCP_RHS(xx, kk)                                                                         \
    scp_bicycle_rhs(xx, u_ref, Lf, Lr, kk);                                                     \
    if (noise_sigma > 0.0) {                                                                    \
        scp_noise_pair(seed, instance, vehicle, noise_counter * 65536u + (stage++), nz, noise_stream); \
        kk[0] += noise_sigma * nz[0];                                                           \
        kk[1] += noise_sigma * nz[1];                                                           \
    }
            SCP_RHS(x, k1)
            for (int i = 0; i < 6; ++i) xt[i] = x[i] + 0.5 * h * k1[i];
            SCP_RHS(xt, k2)
            for (int i = 0; i < 6; ++i) xt[i] = x[i] + 0.5 * h * k2[i];
            SCP_RHS(xt, k3)
            for (int i = 0; i < 6; ++i) xt[i] = x[i] + h * k3[i];
            SCP_RHS(xt, k4)
#undef SCP_RHS
            for (int i = 0; i < 6; ++i) x[i] += h / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
        }
        for (int i = 0; i < 6; ++i) out[s * 6 + i] = x[i];
    }
}

// ================================================================================================ plant step
// The caller's half of one MPC step (main.py), for one vehicle:
//   :104-109  uMax = min(mechanicalSteeringLimit, atan(lateralAccelerationLimit (Lf+Lr) / speed^2)) at the measured state
//   :164-174  clamp of the controller output: U[0] to +-uMax and u0 +- duLim, U[j] to +-uMax and U[j-1] +- duLim
//   :176-191  plant integration over one sample time.  The command computed at step i reaches the actuator
//             ticks_per_sim + ticks_delay_u ticks later, so during step i the plant runs with the PREVIOUS command
//             (controlPathFullRes at the step's last tick = u_path[:, -1] = Iter.u0); the state main.py measures at
//             the next step is the integral over dt with that command held constant.
// The reference integrates with dopri5 (rtol = atol = 1e-8); here classical RK4 with nsub substeps.
// x (6 states) and u_act (the command being actuated) are updated in place; U points at U[b, 0, v], stride nVeh.
SCP_HDFN void scp_plant_step_vehicle(double *x, double *u_act, const double *U, double *Uc, int Hp, int nVeh, double Lf,
                                     double Lr, double mech_limit, double lat_acc_limit, double duLim, double T, int nsub,
                                     double noise_sigma, uint64_t seed, uint32_t instance, uint32_t vehicle,
                                     uint32_t noise_counter, double *uMax_out)
{
    const double speed = x[3];
    const double uMax = fmin(mech_limit, atan(lat_acc_limit * (Lf + Lr) / (speed * speed)));
    const double u0 = *u_act;
    double prev = u0, first = 0.0;
    for (int j = 0; j < (Uc ? Hp : 1); ++j) {
        double uj = U[(size_t)j * nVeh];
        uj = fmin(uj, uMax); uj = fmax(uj, -uMax);
        uj = fmin(uj, prev + duLim); uj = fmax(uj, prev - duLim);
        if (Uc) Uc[(size_t)j * nVeh] = uj;
        if (j == 0) first = uj;
        prev = uj;
    }
    double out[12];
    scp_ode_predict_vehicle(x, u0, Lf, Lr, T, 2, nsub, noise_sigma, seed, instance, vehicle, noise_counter, out, SCP_NOISE_PLANT);
    for (int i = 0; i < 6; ++i) x[i] = out[6 + i];
    *u_act = first;
    if (uMax_out) *uMax_out = uMax;
}

// ================================================================================================ linear advance
// x <- Ad x + Bd u_applied + Ed, u0 <- u_applied with u_applied = U[0, v] clamped as main.py:164-168 does
// (|u| <= uMax, |u - u0| <= duLim).  This is the linearised plant the controller itself predicts with
// (MPC_Iter.py:94-97); the synthetic benchmark closes the loop with it so that the timed region isolates the
// controller stage.  The non-linear plant step of main.py:176-191 is scpb200_plant_step.
SCP_HDFN void scp_advance_vehicle(const double *abe, double u_cmd, double uMax, double duLim, double *x, double *u0)
{
    double ua = u_cmd;
    ua = fmin(ua, uMax); ua = fmax(ua, -uMax);
    ua = fmin(ua, *u0 + duLim); ua = fmax(ua, *u0 - duLim);
    double xn[6];
    for (int i = 0; i < 6; ++i) {
        double acc = abe[42 + i] + abe[36 + i] * ua;
        for (int j = 0; j < 6; ++j) acc += abe[i * 6 + j] * x[j];
        xn[i] = acc;
    }
    for (int i = 0; i < 6; ++i) x[i] = xn[i];
    *u0 = ua;
}

// ================================================================================================ shared pieces
// pos[(v,k)] = cterm_v(k) + sum_{a<=k} g_v[k-a] u_v[a]      (forward_U, SCP_controller.py:199-213)
SCP_FN void scp_positions(Cta &cta, int nVeh, int Hp, const double *g, const double *cterm, const double *u, double *pos,
                          bool coh = false)
{
    CTA_PHASE(tid)
        for (int c = tid; c < nVeh * Hp; c += cta.nt) {
            const int v = c / Hp, k = c - v * Hp;
            const double *gv = g + (size_t)v * Hp * 2;
            double px = scp_ldc(cterm + c * 2, coh), py = scp_ldc(cterm + c * 2 + 1, coh);
            for (int a = 0; a <= k; ++a) {
                px += gv[(k - a) * 2] * u[v * Hp + a];
                py += gv[(k - a) * 2 + 1] * u[v * Hp + a];
            }
            pos[c * 2] = px;
            pos[c * 2 + 1] = py;
        }
    CTA_PHASE_END
}

SCP_FN void scp_row_decode(int nVeh, int Hp, int nObst, int mcv, int r, int *i, int *j, int *o, int *k)
{
    if (r < mcv) {
        const int p = r / Hp;
        *k = r - p * Hp;
        int ii = 0, rem = p;
        while (rem >= nVeh - 1 - ii) { rem -= nVeh - 1 - ii; ++ii; }
        *i = ii; *j = ii + 1 + rem; *o = -1;
    } else {
        const int q = r - mcv;
        *i = q / (nObst * Hp); *o = (q / Hp) % nObst; *k = q % Hp; *j = -1;
    }
}

struct ScpEval {
    double obj, max_violation, sum_violations;
    int feasible;
};

// QCQP_evaluate (SCP_controller.py:215-265) for one instance.  `pos` is [n][2] scratch, `red` reduction scratch.
// obstacle_mode 1 reproduces the reference's nesting of the obstacle loop inside the v2 loop (:249-263).
SCP_FN void scp_evaluate(Cta &cta, int nVeh, int Hp, int nObst, const double *g, const double *cterm, const double *H,
                         const double *qv, double gamma0, const double *u, const double *dsafe,
                         const double *dsafe_obst, const double *obst, double dsafeExtra, double tol,
                         int obstacle_mode, double *pos, double *red, ScpEval *out, double *ci_out, double *cio_out,
                         bool coh = false)
{
    const int n = nVeh * Hp, mcv = Hp * (nVeh * (nVeh - 1) / 2), mc = mcv + Hp * nVeh * nObst;
    scp_positions(cta, nVeh, Hp, g, cterm, u, pos, coh);
    CTA_RED_BEGIN(cta, 3)
    CTA_PHASE(tid)
        double po = 0.0, ps = 0.0, pm = 0.0;
        for (int c = tid; c < n; c += cta.nt) {
            const int v = c / Hp;
            const double *Hr = H + (size_t)c * Hp;
            double acc = 0.0;
            for (int bb = 0; bb < Hp; ++bb) acc += scp_ldc(Hr + bb, coh) * u[v * Hp + bb];
            po += u[c] * (acc + scp_ldc(qv + c, coh));
        }
        if (ci_out)
            for (int e = tid; e < nVeh * nVeh * Hp; e += cta.nt) {
                const int i = e / (nVeh * Hp), j = (e / Hp) % nVeh;
                if (i == j) ci_out[e] = -INFINITY;
            }
        for (int r = tid; r < mc; r += cta.nt) {
            int i, j, o, k;
            scp_row_decode(nVeh, Hp, nObst, mcv, r, &i, &j, &o, &k);
            double dx, dy, sbar;
            int mult = 1;
            if (o < 0) {
                dx = pos[(i * Hp + k) * 2] - pos[(j * Hp + k) * 2];
                dy = pos[(i * Hp + k) * 2 + 1] - pos[(j * Hp + k) * 2 + 1];
                sbar = dsafe[i * nVeh + j] + dsafeExtra;
            } else {
                dx = pos[(i * Hp + k) * 2] - obst[(o * Hp + k) * 2];
                dy = pos[(i * Hp + k) * 2 + 1] - obst[(o * Hp + k) * 2 + 1];
                sbar = dsafe_obst[i * nObst + o] + dsafeExtra;
                if (obstacle_mode == 1) mult = nVeh - 1 - i;
            }
            const double ci = sbar * sbar - (dx * dx + dy * dy);
            if (o < 0) {
                if (ci_out) { ci_out[(i * nVeh + j) * Hp + k] = ci; ci_out[(j * nVeh + i) * Hp + k] = ci; }
            } else if (cio_out) {
                cio_out[(i * nObst + o) * Hp + k] = (mult > 0) ? ci : -INFINITY;
            }
            if (ci > tol && mult > 0) { ps += mult * ci; pm = fmax(pm, ci); }
        }
        CTA_RED_SUM(cta, red, 0, tid, po)
        CTA_RED_SUM(cta, red, 1, tid, ps)
        CTA_RED_MAX(cta, red, 2, tid, pm)
    CTA_PHASE_END_RED(cta, red, 3)
    out->obj = cta_red_sum(cta, red, 0) + gamma0;
    out->sum_violations = cta_red_sum(cta, red, 1);
    out->max_violation = cta_red_max(cta, red, 2);
    out->feasible = out->max_violation > 0.0 ? 0 : 1;
}

// Linearisation about ubar (SCP_controller.py:97-114 through the structured identity of ops_pair.cuh):
// dbar[r] = relative position at ubar, bA[r] = -sbar^2 - |dbar|^2 + 2 dbar.(c_i(k) - c_j(k)).
SCP_FN void scp_linearise(Cta &cta, int nVeh, int Hp, int nObst, const double *g, const double *cterm,
                          const double *ubar, const double *dsafe, const double *dsafe_obst, const double *obst,
                          double dsafeExtra, double *pos, double *dbar, double *bA, bool coh = false)
{
    const int mcv = Hp * (nVeh * (nVeh - 1) / 2), mc = mcv + Hp * nVeh * nObst;
    scp_positions(cta, nVeh, Hp, g, cterm, ubar, pos, coh);
    CTA_PHASE(tid)
        for (int r = tid; r < mc; r += cta.nt) {
            int i, j, o, k;
            scp_row_decode(nVeh, Hp, nObst, mcv, r, &i, &j, &o, &k);
            double dx, dy, bx, by, sbar;
            if (o < 0) {
                dx = pos[(i * Hp + k) * 2] - pos[(j * Hp + k) * 2];
                dy = pos[(i * Hp + k) * 2 + 1] - pos[(j * Hp + k) * 2 + 1];
                bx = scp_ldc(cterm + (i * Hp + k) * 2, coh) - scp_ldc(cterm + (j * Hp + k) * 2, coh);
                by = scp_ldc(cterm + (i * Hp + k) * 2 + 1, coh) - scp_ldc(cterm + (j * Hp + k) * 2 + 1, coh);
                sbar = dsafe[i * nVeh + j] + dsafeExtra;
            } else {
                const double ox = obst[(o * Hp + k) * 2], oy = obst[(o * Hp + k) * 2 + 1];
                dx = pos[(i * Hp + k) * 2] - ox;
                dy = pos[(i * Hp + k) * 2 + 1] - oy;
                bx = scp_ldc(cterm + (i * Hp + k) * 2, coh) - ox;
                by = scp_ldc(cterm + (i * Hp + k) * 2 + 1, coh) - oy;
                sbar = dsafe_obst[i * nObst + o] + dsafeExtra;
            }
            dbar[r * 2] = dx;
            dbar[r * 2 + 1] = dy;
            bA[r] = -sbar * sbar - (dx * dx + dy * dy) + 2.0 * (dx * bx + dy * by);
        }
    CTA_PHASE_END
}

// ================================================================================================ working set
// Two-pool bump allocator for a CTA's working set: shared memory first; whatever does not fit under `sh_lim`
// overflows into the CTA's slice of the global workspace (L2-resident).  The same function runs on the host
// (null bases) to size both pools, so the layout is a pure function of the problem dimensions.
struct ScpBump {
    double *sh;
    size_t sh_off, sh_lim;      // doubles
    double *gl;
    size_t gl_off;
    bool all_shared;            // compile-time true in the kernels instantiated for fully shared-resident sets
    bool last_shared;
    SCP_HDMFN double *take(size_t nd)
    {
        nd = (nd + 1) & ~(size_t)1;      // keep 16-byte alignment
        if (all_shared || sh_off + nd <= sh_lim) {
#ifdef __CUDA_ARCH__
            double *p = sh + sh_off;             // never null on the device (a select here is paid at every use)
#else
            double *p = sh ? sh + sh_off : 0;
#endif
            sh_off += nd;
            last_shared = true;
            return p;
        }
        double *p = gl ? gl + gl_off : 0;
        gl_off += nd;
        last_shared = false;
        return p;
    }
};

SCP_HDFN ScpBump scp_bump(double *sh, size_t sh_lim, double *gl, bool all_shared)
{
    ScpBump bp;
    bp.sh = sh; bp.sh_off = 0; bp.sh_lim = sh_lim; bp.gl = gl; bp.gl_off = 0; bp.all_shared = all_shared; bp.last_shared = true;
    return bp;
}

// Vectors of the interior-point working set (small and hot first); the tile scratch and the normal matrix are
// carved last by ipm_carve_big so that they are the first to overflow.
// red_doubles: size of the per-warp reduction scratch of the kernel that will run on this layout (SCP_RED_DOUBLES of
// ITS translation unit: units compiled for wider CTAs carry more warps)
// nr: difference (steering-rate) rows, 0 = none (the layout without them is unchanged)
SCP_HDFN void ipm_carve(ScpBump &bp, IpmMem &m, int n1, int mc, int red_doubles = SCP_RED_DOUBLES, int nr = 0, int rper = 1)
{
    m.n1 = n1; m.n1p = ipm_padded(n1); m.T = m.n1p / SCP_TILE; m.mc = mc;
    m.nr = nr; m.rper = rper > 0 ? rper : 1;
    m.red = bp.take((size_t)red_doubles);
    m.t8 = bp.take(16);
    m.x = bp.take(m.n1p); m.q = bp.take(m.n1p); m.rx = bp.take(m.n1p); m.dx = bp.take(m.n1p); m.tn = bp.take(m.n1p);
    m.dinv = bp.take(m.n1p);
    m.ub = bp.take(m.n1p); m.sU = bp.take(m.n1p); m.zU = bp.take(m.n1p); m.dsU = bp.take(m.n1p);
    m.dzU = bp.take(m.n1p); m.ccU = bp.take(m.n1p); m.eU = bp.take(m.n1p);
    m.lb = bp.take(m.n1p); m.sL = bp.take(m.n1p); m.zL = bp.take(m.n1p); m.dsL = bp.take(m.n1p);
    m.dzL = bp.take(m.n1p); m.ccL = bp.take(m.n1p); m.eL = bp.take(m.n1p);
    m.bA = bp.take(mc); m.sA = bp.take(mc); m.zA = bp.take(mc); m.rzA = bp.take(mc);
    {
        const size_t rows = (size_t)((mc + 1) & ~1);          // dsA, dzA, ccA back to back (see ipm_carve_big)
        double *blk = bp.take(3 * rows);
#ifdef __CUDA_ARCH__
        m.dsA = blk; m.dzA = blk + rows; m.ccA = blk + 2 * rows;
#else
        m.dsA = blk; m.dzA = blk ? blk + rows : 0; m.ccA = blk ? blk + 2 * rows : 0;
#endif
    }
    m.eA = bp.take(mc);
    if (nr > 0) {
        m.hP = bp.take(nr); m.sP = bp.take(nr); m.zP = bp.take(nr); m.dsP = bp.take(nr); m.dzP = bp.take(nr); m.ccP = bp.take(nr); m.eP = bp.take(nr);
        m.hM = bp.take(nr); m.sM = bp.take(nr); m.zM = bp.take(nr); m.dsM = bp.take(nr); m.dzM = bp.take(nr); m.ccM = bp.take(nr); m.eM = bp.take(nr);
        m.rsub = bp.take(nr);
    } else {
        m.hP = m.sP = m.zP = m.dsP = m.dzP = m.ccP = m.eP = 0;
        m.hM = m.sM = m.zM = m.dsM = m.dzM = m.ccM = m.eM = 0;
        m.rsub = 0;
    }
}

SCP_HDFN void ipm_carve_big(ScpBump &bp, IpmMem &m)
{
    m.S = bp.take((size_t)(m.T * (m.T + 1) / 2) * SCP_TILE2);
    m.S_far = !bp.last_shared;
}

struct ScpMem {
    IpmMem ipm;
    double *g, *dbar, *resp, *frc, *ucur, *Msm, *Hs;
    int *rowtab;       // [mc] per constraint row: (i*Hp + k) | (j*Hp + k) << 16  (obstacle rows: vehicle index only)
    int alpha_slots;
    bool H_local;      // Hs is shared-resident: the instance's cost blocks are copied there once per instance
};

// alpha_slots: pair-block mode of the normal matrix (> 0: tensor path, 0: entry by entry; see PairOp)
// want_H: keep a shared-memory copy of the instance's cost blocks (read every iteration); otherwise they are read
// from global memory.
SCP_HDFN void scp_carve(ScpBump &bp, ScpMem &s, int nVeh, int Hp, int nObst, int alpha_slots, int want_H,
                        int red_doubles = SCP_RED_DOUBLES, int rate_rows = 0)
{
    const int n = nVeh * Hp, mc = Hp * (nVeh * (nVeh - 1) / 2 + nVeh * nObst);
    ipm_carve(bp, s.ipm, n + 1, mc, red_doubles, rate_rows ? n : 0, Hp);
    s.g = bp.take((size_t)n * 2);
    s.dbar = bp.take((size_t)mc * 2);
    s.resp = bp.take((size_t)n * 2);
    s.frc = bp.take((size_t)n * 2);
    s.rowtab = (int *)bp.take((size_t)(mc + 1) / 2);
    // ucur (the SCP iterate) is dead while the interior-point method runs and rx is dead outside it: one array.
    s.ucur = s.ipm.rx;
    // Msm (per (vehicle, step) 2x2 aggregates) is only live inside form_normal, where ccA is dead.
    if ((size_t)n * 3 <= (size_t)mc) s.Msm = s.ipm.ccA;
    else s.Msm = bp.take((size_t)n * 3);
    s.alpha_slots = alpha_slots;
    s.Hs = want_H ? bp.take((size_t)n * Hp) : 0;
    s.H_local = want_H && bp.last_shared;
    ipm_carve_big(bp, s.ipm);
}

// total doubles of the working set / split under a shared-memory limit (host-side planning)
SCP_HDFN void scp_footprint(int nVeh, int Hp, int nObst, int alpha_slots, int want_H, size_t sh_lim, size_t *sh_used,
                            size_t *gl_used, int red_doubles = SCP_RED_DOUBLES, int rate_rows = 0)
{
    ScpBump bp = scp_bump(0, sh_lim, 0, false);
    ScpMem s;
    scp_carve(bp, s, nVeh, Hp, nObst, alpha_slots, want_H, red_doubles, rate_rows);
    *sh_used = bp.sh_off;
    *gl_used = bp.gl_off;
}

SCP_HDFN void ipm_footprint(int n1, int mc, size_t sh_lim, size_t *sh_used, size_t *gl_used)
{
    ScpBump bp = scp_bump(0, sh_lim, 0, false);
    IpmMem m;
    ipm_carve(bp, m, n1, mc);
    ipm_carve_big(bp, m);
    *sh_used = bp.sh_off;
    *gl_used = bp.gl_off;
}

// ================================================================================================ K4: the SCP loop
struct ScpIO {
    const double *g, *cterm, *H, *qv, *gamma0, *dsafe, *dsafe_obst, *obst;   // batch base pointers
    double *u, *traj, *U, *log, *obj, *max_violation;
    int32_t *scp_iters, *ipm_iters, *status;
    // Preemption (work-queue scheduling at QP granularity): with `state` non-null an invocation runs at most `quantum`
    // SCP iterations of the instance, then parks it (u in io.u, the loop scalars in state[b]) for a later invocation
    // by any CTA.  The arithmetic is that of the uninterrupted loop, so results are bit-identical.
    double *state;      // [B][SCP_STATE_W] or null (run to completion)
    int quantum;
    double *snap;       // [B][ipm_snap_doubles] interior-point warm-start iterates, or null (every QP starts cold)
    int coherent;       // the set-up outputs (g, cterm, H, qv, gamma0) were written during this launch: read them past L1
    const double *u_prev;   // [B][nVeh] command being actuated (Iter.u0): anchors the first steering-rate row of every
                            // vehicle; read only when params.enable_rate_rows (then required)
};
#define SCP_STATE_W 8   /* obj0, mv0, it, ipm_total, status bits, pinned (never parked), snapshot valid, (spare) */

// SCP_optimizer (SCP_controller.py:74-197) + the result shaping of SCP_controller (:68-70) for instance b.
// Returns true when the instance is finished (results written), false when it was parked.
// pin: -1 = the instance's own flag (state[5], set when the queue is filled); 0 / 1 = the caller's decision for this
// invocation (the rollout entry pins the instances that are furthest behind).  Must be uniform over the CTA.
SCP_FN bool scp_solve_instance(Cta &cta, const scpb200_dims &d, const scpb200_params &p, int b, const ScpIO &io,
                               ScpMem &s, int pin = -1)
{
    const int nVeh = d.nVeh, Hp = d.Hp, nObst = d.nObst, n = nVeh * Hp;
    const int mcv = Hp * (nVeh * (nVeh - 1) / 2), mc = mcv + Hp * nVeh * nObst;
    IpmMem &m = s.ipm;
    const double *gB = io.g + (size_t)b * n * 2, *cB = io.cterm + (size_t)b * n * 2;
    const double *HB = io.H + (size_t)b * n * Hp, *qB = io.qv + (size_t)b * n;
    const double *dsB = io.dsafe + (size_t)b * nVeh * nVeh;
    const double *dsoB = nObst ? io.dsafe_obst + (size_t)b * nVeh * nObst : 0;
    const double *obB = nObst ? io.obst + (size_t)b * nObst * Hp * 2 : 0;
    const bool coh = io.coherent != 0;
    const double gamma0 = scp_ldc(io.gamma0 + b, coh);
    double *uB = io.u + (size_t)b * n;

    IpmCtl ctl;
    ctl.abstol = p.qp_abstol; ctl.reltol = p.qp_reltol; ctl.feastol = p.qp_feastol;
    ctl.dual_reg = p.qp_dual_reg; ctl.inf_bound = p.inf_bound; ctl.max_iter = p.ipm_max_iter;
    ctl.dres_floor = p.qp_dres_floor_factor * p.qp_feastol;
    ctl.snap = (io.snap && p.qp_warm_start) ? io.snap + (size_t)b * ipm_snap_doubles(m.n1p, mc, m.nr) : 0;
    ctl.snap_relgap = p.qp_warm_relgap;
    ctl.warm = 0;
    ctl.snap_min_iter = p.qp_warm_min_iter;
    int snap_valid = 0;

    PairOp op;
    op.nVeh = nVeh; op.Hp = Hp; op.n = n; op.nObst = nObst; op.mcv = mcv; op.mc = mc;
    op.g = s.g; op.H = s.H_local ? s.Hs : HB; op.dbar = s.dbar; op.resp = s.resp; op.frc = s.frc; op.red = m.red;
    op.xom = 0.0; op.wsum = 0.0;
    op.Msm = s.Msm; op.alpha_slots = s.alpha_slots; op.rowtab = s.rowtab;
    op.coh = coh && !s.H_local;      // the operator reads H only: past L1 when it is the global array another CTA wrote, plainly from the shared copy

    double *stB = io.state ? io.state + (size_t)b * SCP_STATE_W : 0;
    const int it_resume = stB ? (int)SCP_LD_COHERENT(stB + 2) : 0;       // > 0: a parked instance
    const int snap_carried = stB ? (int)SCP_LD_COHERENT(stB + 6) : 0;

    // instance data -> shared; warm start (SCP_controller.py:42-43) with the eps tweak of :75-76
    CTA_PHASE(tid)
        for (int e = tid; e < n * 2; e += cta.nt) s.g[e] = scp_ldc(gB + e, coh);
        for (int r = tid; r < mc; r += cta.nt) {
            int i, j, o, k;
            scp_row_decode(nVeh, Hp, nObst, mcv, r, &i, &j, &o, &k);
            s.rowtab[r] = (i * Hp + k) | ((j >= 0 ? j * Hp + k : 0xffff) << 16);
        }
        if (s.H_local)
            for (int e = tid; e < n * Hp; e += cta.nt) s.Hs[e] = scp_ldc(HB + e, coh);
        for (int c = tid; c < m.n1p; c += cta.nt) {
            double uv = 0.0;
            if (c < n) {
                uv = SCP_LD_COHERENT(uB + c);
                if (c == 0 && it_resume == 0 && fabs(uv) < 2.220446049250313e-16) uv = 2.220446049250313e-16;
            }
            s.ucur[c] = uv;
            m.q[c] = (c < n) ? scp_ldc(qB + c, coh) : (c == n ? p.omega_weight : 0.0);
        }
    CTA_PHASE_END

    ScpEval ev;
    double obj0, mv0;
    int it = 0, ipm_total = 0, st = 0, stopped = 0;
    if (it_resume == 0) {
        scp_evaluate(cta, nVeh, Hp, nObst, s.g, cB, HB, qB, gamma0, s.ucur, dsB, dsoB, obB, p.dsafeExtra, p.constraint_tol,
                     p.obstacle_eval_mode, s.resp, m.red, &ev, 0, 0, coh);
        obj0 = ev.obj; mv0 = ev.max_violation;
        snap_valid = snap_carried;
    } else {
        obj0 = SCP_LD_COHERENT(stB + 0); mv0 = SCP_LD_COHERENT(stB + 1);
        it = it_resume; ipm_total = (int)SCP_LD_COHERENT(stB + 3); st = (int)SCP_LD_COHERENT(stB + 4);
        snap_valid = (int)SCP_LD_COHERENT(stB + 6);
        ev.obj = obj0; ev.max_violation = mv0; ev.sum_violations = 0.0; ev.feasible = mv0 > 0.0 ? 0 : 1;
    }
    const bool parks = stB && (pin < 0 ? stB[5] == 0.0 : pin == 0);
    const int it_park = parks ? it + (io.quantum > 0 ? io.quantum : 1) : p.max_scp_iter;
    for (; it < p.max_scp_iter; ++it) {
        if (it >= it_park) {
            // park: u and the loop scalars go back to global memory; another invocation continues from here
            CTA_PHASE(tid)
                for (int c = tid; c < n; c += cta.nt) uB[c] = s.ucur[c];
                if (tid == 0) {
                    stB[0] = obj0; stB[1] = mv0; stB[2] = (double)it; stB[3] = (double)ipm_total; stB[4] = (double)st;
                    stB[6] = (double)snap_valid;
                }
            CTA_PHASE_END
            return false;
        }
        scp_linearise(cta, nVeh, Hp, nObst, s.g, cB, s.ucur, dsB, dsoB, obB, p.dsafeExtra, s.resp, s.dbar, m.bA, coh);
        CTA_PHASE(tid)
            for (int c = tid; c < m.n1p; c += cta.nt) {
                double lo = -p.uLim, hi = p.uLim;
                if (c < n) {
                    if (p.trust_radius < 1e300) {
                        lo = fmax(lo, s.ucur[c] - p.trust_radius);
                        hi = fmin(hi, s.ucur[c] + p.trust_radius);
                    }
                } else if (c == n) { lo = 0.0; hi = p.omega_ub; }
                m.lb[c] = lo;
                m.ub[c] = hi;
            }
            // steering-rate rows (extension; the reference clamps after the solve, main.py:164-174):
            // |u_v[k] - u_v[k-1]| <= duLim with u_v[-1] = the command being actuated
            for (int c = tid; c < m.nr; c += cta.nt) {
                const int v = c / Hp;
                const double up = (c - v * Hp == 0) ? SCP_LD_COHERENT(io.u_prev + (size_t)b * nVeh + v) : 0.0;
                m.hP[c] = p.duLim + up;
                m.hM[c] = p.duLim - up;
            }
        CTA_PHASE_END
        IpmResult res;
        int warm_iters = 0;                                 // iterations of an abandoned warm attempt (logged with the QP)
        ctl.warm = ctl.snap && snap_valid;
        {
            // ONE call site of the solver for the warm attempt, its cold restart and the plain cold start: inlined three times
            // the interior-point method was two thirds of the kernel's 390 KB of code, and the instruction cache of an SM
            // that runs three CTAs in different phases missed on the critical lanes (ncu: stall_no_inst next to the
            // shared-memory stalls in the substitution).
            // A warm start that does not converge quickly (the active set moved too far) is abandoned for a cold one.
            const int cap = ctl.max_iter;
#pragma unroll 1
            for (int attempt = 0; attempt < 2; ++attempt) {
                ctl.max_iter = (ctl.warm && p.qp_warm_max_iter > 0) ? p.qp_warm_max_iter : cap;
                ipm_solve(cta, op, m, ctl, &res);
                if (!(ctl.warm && (res.status & SCPB200_ST_QP_MAXITER))) break;
                warm_iters = res.iters; ipm_total += res.iters; ctl.warm = 0;
            }
            ctl.max_iter = cap;
        }
        // a warm-started QP that converged before it reached the iteration a new iterate is taken from leaves the old
        // one in place: it was a good start for this QP and the next one is closer still (SCP is converging)
        snap_valid = res.snap_saved || (ctl.warm && !(res.status & SCPB200_ST_QP_MAXITER) && snap_valid);
        ipm_total += res.iters;
        if (res.status & SCPB200_ST_QP_MAXITER) st |= SCPB200_ST_QP_MAXITER;
        if (res.status & SCPB200_ST_QP_PIVOT) st |= SCPB200_ST_QP_PIVOT;
        if (res.status & SCPB200_ST_QP_DRES_FLOOR) st |= SCPB200_ST_QP_DRES_FLOOR;
        const double slack = m.x[n];
        CTA_PHASE(tid)
            for (int c = tid; c < n; c += cta.nt) s.ucur[c] = m.x[c];
        CTA_PHASE_END
        scp_evaluate(cta, nVeh, Hp, nObst, s.g, cB, HB, qB, gamma0, s.ucur, dsB, dsoB, obB, p.dsafeExtra,
                     p.constraint_tol, p.obstacle_eval_mode, s.resp, m.red, &ev, 0, 0, coh);
        const double fval = res.fval + gamma0;
        const double merit0 = obj0 + p.omega_weight * mv0;
        const double delta_hat = merit0 - fval;
        const double delta = merit0 - (ev.obj + p.omega_weight * ev.max_violation);
        obj0 = ev.obj;
        mv0 = ev.max_violation;
        if (io.log) {
            CTA_PHASE(tid)
                if (tid == 0) {
                    const int log_rows = p.log_capacity > 0 ? p.log_capacity : p.max_scp_iter;     // rows per instance
                    double *L = io.log + ((size_t)b * log_rows + it) * SCPB200_LOG_W;
                    L[0] = slack; L[1] = fval; L[2] = ev.obj; L[3] = delta_hat; L[4] = delta; L[5] = ev.feasible;
                    L[6] = ev.max_violation; L[7] = ev.sum_violations;
                    L[8] = res.iters + warm_iters;          // sum over the log = ipm_iters[b]
                    L[9] = res.status | (warm_iters ? SCPB200_ST_QP_WARM_RESTART : 0);
#ifdef SCP_DIAG_LOG     /* tuning builds: where did the interior-point method stop */
                    L[5] = res.dres; L[6] = res.pres; L[7] = res.relgap; L[3] = res.gap;
#endif
                }
            CTA_PHASE_END
        }
        if (nVeh == 1 && fabs(delta) < p.delta_tol && ev.max_violation > p.constraint_tol) { ++it; stopped = 1; break; }
        if (fabs(delta) < p.delta_tol && ev.max_violation <= p.constraint_tol) { ++it; stopped = 1; break; }
    }
    if (!stopped) st |= SCPB200_ST_SCP_MAXITER;
    if (!ev.feasible) st |= SCPB200_ST_INFEASIBLE;
    // results: u, forward_U shapes traj[Hp][2][nVeh], U[Hp][nVeh]
    scp_positions(cta, nVeh, Hp, s.g, cB, s.ucur, s.resp, coh);
    CTA_PHASE(tid)
        for (int c = tid; c < n; c += cta.nt) {
            const int v = c / Hp, k = c - v * Hp;
            uB[c] = s.ucur[c];
            if (io.U) io.U[((size_t)b * Hp + k) * nVeh + v] = s.ucur[c];
            if (io.traj) {
                io.traj[(((size_t)b * Hp + k) * 2 + 0) * nVeh + v] = s.resp[c * 2];
                io.traj[(((size_t)b * Hp + k) * 2 + 1) * nVeh + v] = s.resp[c * 2 + 1];
            }
        }
        if (tid == 0) {
            if (io.scp_iters) io.scp_iters[b] = it;
            if (io.ipm_iters) io.ipm_iters[b] = ipm_total;
            if (io.status) io.status[b] = st;
            if (io.obj) io.obj[b] = ev.obj;
            if (io.max_violation) io.max_violation[b] = ev.max_violation;
            if (stB) stB[6] = (double)snap_valid;          // the iterate a following call may start from (qp_warm_carry)
        }
    CTA_PHASE_END
    return true;
}

// ================================================================================================ rollout
// Closed-loop MPC steps of ONE instance inside the solve kernel (scpb200_mpc_rollout): instances of a Monte-Carlo batch
// are independent ACROSS MPC steps too, so a CTA that finishes an instance's SCP loop closes the loop for it (clamp +
// plant / linear advance), runs the next step's set-up (K1) and hands the instance back to the work queue — no per-step
// kernel boundary, hence no step that ends on its longest chain of QPs while most CTAs idle (measured: 20 % of the warp
// samples of the per-step kernel wait at the queue).  The arithmetic per step is that of scpb200_mpc_setup ->
// scpb200_scp_solve -> scpb200_advance_linear (mode 0) or of BatchSCP.mpc_step (mode 1): results are bit-identical.
struct ScpRollout {
    int nsteps;                // 0: the kernel runs in per-step mode
    int mode;                  // 0: advance on the controller's linear model; 1: delay compensation + plant integration
    uint32_t counter0;         // noise counter of step 0 (step s draws with counter0 + s)
    const double *veh, *poly;  // [B,nVeh,5], [B,nVeh,nPts,2]
    double *x0, *u0;           // [B,nVeh,6], [B,nVeh]  set-up inputs (mode 0: advanced in place; mode 1: written per step)
    double *x_meas, *u_act;    // mode 1: measured state / command being actuated, advanced in place
    double *ref, *g, *cterm, *H, *qv, *gamma0, *abe;   // set-up outputs (per-instance slices of the batch arrays)
    int32_t *setup_status;
    double uMax, duLim, mech_limit, lat_acc_limit, delay;
    int nsub_delay, nsub_plant;
    int32_t *qp_total, *ipm_total, *status_or;      // [B] accumulated over the steps
    int32_t *scp_iters_hist, *status_hist;          // [B,nsteps] or null
    double *U_hist, *x_hist;                        // [B,nsteps,Hp,nVeh], [B,nsteps+1,nVeh,6] or null
};

// Set-up of instance b for its MPC step `step`.  `stage` is CTA scratch of >= 7 nVeh doubles.
SCP_FN void scp_rollout_setup(Cta &cta, const scpb200_dims &d, const scpb200_params &p, const ScpRollout &ro, int b, int step,
                              double *stage, double *red, int *flag, double *k1ws, int k1_warps)
{
    const int nVeh = d.nVeh;
    scpb200_params ps = p;
    ps.noise_counter = ro.counter0 + (uint32_t)step;
    double *xs = stage, *us = stage + (size_t)nVeh * 6;
    CTA_PHASE(tid)
        for (int v = tid; v < nVeh; v += cta.nt) {
            const size_t iv = (size_t)b * nVeh + v;
            double x[6];
            if (ro.mode == 1) {
                // IterClass (MPC_Iter.py:25-33): the state the command will meet, delay_x + dt + delay_u ahead
                double xm[6], out[12];
                for (int i = 0; i < 6; ++i) xm[i] = scp_ldc(ro.x_meas + iv * 6 + i, true);
                const double ua = scp_ldc(ro.u_act + iv, true);
                scp_ode_predict_vehicle(xm, ua, ro.veh[iv * 5], ro.veh[iv * 5 + 1], ro.delay, 2, ro.nsub_delay, ps.noise_sigma, ps.seed,
                                        ps.instance0 + (uint32_t)b, (uint32_t)v, ps.noise_counter, out, SCP_NOISE_ODE);
                for (int i = 0; i < 6; ++i) { x[i] = out[6 + i]; ro.x0[iv * 6 + i] = x[i]; }
                ro.u0[iv] = ua;
                us[v] = ua;
            } else {
                for (int i = 0; i < 6; ++i) x[i] = scp_ldc(ro.x0 + iv * 6 + i, true);
                us[v] = scp_ldc(ro.u0 + iv, true);
            }
            for (int i = 0; i < 6; ++i) xs[v * 6 + i] = x[i];
            if (ro.x_hist && step == 0)
                for (int i = 0; i < 6; ++i)
                    ro.x_hist[(((size_t)b * (ro.nsteps + 1)) * nVeh + v) * 6 + i] = ro.mode == 1 ? scp_ldc(ro.x_meas + iv * 6 + i, true) : x[i];
        }
    CTA_PHASE_END
    scp_setup_instance_at(cta, d, ps, b, xs, us, ro.veh, ro.poly, ro.ref, ro.g, ro.cterm, ro.H, ro.qv, ro.gamma0, ro.abe,
                          ro.setup_status, red, flag, true, k1ws, k1_warps);
}

// After the SCP loop of step `step` has finished (results of the step in io.U / io.scp_iters / ...): record, close the
// loop, account.  Returns nothing; the caller advances the step counter.
SCP_FN void scp_rollout_advance(Cta &cta, const scpb200_dims &d, const scpb200_params &p, const ScpRollout &ro, const ScpIO &io,
                                int b, int step)
{
    const int nVeh = d.nVeh, Hp = d.Hp;
    CTA_PHASE(tid)
        if (ro.U_hist)
            for (int e = tid; e < Hp * nVeh; e += cta.nt)
                ro.U_hist[((size_t)b * ro.nsteps + step) * Hp * nVeh + e] = io.U[(size_t)b * Hp * nVeh + e];
        for (int v = tid; v < nVeh; v += cta.nt) {
            const size_t iv = (size_t)b * nVeh + v;
            double x[6];
            if (ro.mode == 1) {
                for (int i = 0; i < 6; ++i) x[i] = scp_ldc(ro.x_meas + iv * 6 + i, true);
                double ua = scp_ldc(ro.u_act + iv, true);
                scp_plant_step_vehicle(x, &ua, io.U + (size_t)b * Hp * nVeh + v, (double *)0, Hp, nVeh, ro.veh[iv * 5], ro.veh[iv * 5 + 1],
                                       ro.mech_limit, ro.lat_acc_limit, ro.duLim, p.dt, ro.nsub_plant, p.noise_sigma, p.seed,
                                       p.instance0 + (uint32_t)b, (uint32_t)v, ro.counter0 + (uint32_t)step, (double *)0);
                for (int i = 0; i < 6; ++i) ro.x_meas[iv * 6 + i] = x[i];
                ro.u_act[iv] = ua;
            } else {
                double abe[48];
                for (int i = 0; i < 48; ++i) abe[i] = scp_ldc(ro.abe + iv * 48 + i, true);
                for (int i = 0; i < 6; ++i) x[i] = scp_ldc(ro.x0 + iv * 6 + i, true);
                double u0 = scp_ldc(ro.u0 + iv, true);
                scp_advance_vehicle(abe, io.U[(size_t)b * Hp * nVeh + v], ro.uMax, ro.duLim, x, &u0);
                for (int i = 0; i < 6; ++i) ro.x0[iv * 6 + i] = x[i];
                ro.u0[iv] = u0;
            }
            if (ro.x_hist)
                for (int i = 0; i < 6; ++i) ro.x_hist[(((size_t)b * (ro.nsteps + 1) + step + 1) * nVeh + v) * 6 + i] = x[i];
        }
        if (tid == 0) {
            const int its = io.scp_iters[b], st = io.status[b];
            if (step == 0) { ro.qp_total[b] = 0; ro.ipm_total[b] = 0; ro.status_or[b] = 0; }
            ro.qp_total[b] += its;
            ro.ipm_total[b] += io.ipm_iters[b];
            ro.status_or[b] |= st;
            if (ro.scp_iters_hist) ro.scp_iters_hist[(size_t)b * ro.nsteps + step] = its;
            if (ro.status_hist) ro.status_hist[(size_t)b * ro.nsteps + step] = st;
        }
    CTA_PHASE_END
}

// ================================================================================================ K3: dense QP
struct QpIO {
    const double *P, *q, *A, *b, *lb, *ub;
    double *x, *fval, *zA;
    int32_t *iters, *status;
};

SCP_FN void qp_solve_instance(Cta &cta, const scpb200_params &p, int n1, int mc, int b, const QpIO &io, IpmMem &m)
{
    IpmCtl ctl;
    ctl.abstol = p.qp_abstol; ctl.reltol = p.qp_reltol; ctl.feastol = p.qp_feastol;
    ctl.dual_reg = p.qp_dual_reg; ctl.inf_bound = p.inf_bound; ctl.max_iter = p.ipm_max_iter;
    ctl.dres_floor = p.qp_dres_floor_factor * p.qp_feastol;
    ctl.snap = 0; ctl.snap_relgap = 0.0; ctl.warm = 0; ctl.snap_min_iter = 0;
    DenseOp op;
    op.n1 = n1; op.mc = mc;
    op.P = io.P + (size_t)b * n1 * n1;
    op.A = io.A + (size_t)b * mc * n1;
    op.xp = 0; op.wp = 0;
    CTA_PHASE(tid)
        for (int c = tid; c < m.n1p; c += cta.nt) {
            m.q[c] = c < n1 ? io.q[(size_t)b * n1 + c] : 0.0;
            m.ub[c] = c < n1 ? io.ub[(size_t)b * n1 + c] : 1e300;
            m.lb[c] = c < n1 ? io.lb[(size_t)b * n1 + c] : -1e300;
        }
        for (int r = tid; r < mc; r += cta.nt) m.bA[r] = io.b[(size_t)b * mc + r];
    CTA_PHASE_END
    IpmResult res;
    ipm_solve(cta, op, m, ctl, &res);
    CTA_PHASE(tid)
        for (int c = tid; c < n1; c += cta.nt) io.x[(size_t)b * n1 + c] = m.x[c];
        if (io.zA)
            for (int r = tid; r < mc; r += cta.nt) io.zA[(size_t)b * mc + r] = m.zA[r];
        if (tid == 0) {
            io.fval[b] = res.fval;
            if (io.iters) io.iters[b] = res.iters;
            if (io.status) io.status[b] = res.status;
        }
    CTA_PHASE_END
}

// ================================================================================================ K2: dense assembly
// One CTA per work item (instance, row range): linearise about ubar, then write the dense QP of
// SCP_controller.py:93-128 to HBM in the reference's layout.  The kernel is a pure streaming write (238 KB per
// instance at Hp = 10 against 2 KB of input) and 86 % of what it writes is structural zeros (causal rows,
// block-diagonal P), so it is organised around the store stream:
//   * the item's P and A ranges are cut into 4 KiB chunks of the flat arrays; every warp owns two chunk buffers in
//     shared memory and works through its chunks alone (no CTA barrier after the linearisation);
//   * per chunk: zero the buffer (8 STS.128 per lane), write the chunk's few non-zeros (cost blocks, causal row
//     segments, the -1 column) into it, fence towards the async proxy, and hand the finished 4 KiB to the TMA
//     (cp.async.bulk global <- shared, SASS UBLKCP); the other buffer is composed while this one drains.
// Every byte of the output is therefore written exactly once by a full-width bulk store: no second pass over the
// non-zeros and no sector read-modify-write in L2 (the previous version zero-filled with bulk stores and then
// scattered 8-byte non-zeros over the zeros: 30.6 MB read back per 246 MB written).
// Measured on B200 (246 MB per launch, B = 1024, Hp = 10): per-element 16-byte streaming stores 92 us; CTA-wide
// staging with barriers 107 us; bulk zero-fill + scattered non-zeros 67.5 us; this version: see profiles/; a plain
// memset of the same bytes 41 us.
#ifndef SCP_ASM_CHUNK
#define SCP_ASM_CHUNK 512          /* doubles per chunk buffer (4 KiB) */
#endif
#ifndef SCP_ASM_NBUF
#define SCP_ASM_NBUF 2             /* chunk buffers per warp */
#endif

// the chunk [lo, lo + len) of instance-local flat A (row-major [mc][n1]) into buf (already zero).
// Lane mapping: (row of the chunk = lane >> 2, +8, ...) x (a = lane & 3, +4, ... <= Hp; a == Hp is the -1 column):
// no divisions in the loops.
SCP_FN void scp_compose_A(int lane, int nVeh, int Hp, int n1, const int *rowinfo, const double *gs, const double *dbar,
                          int lo, int len, double *buf)
{
    (void)nVeh;
    const int n = n1 - 1;
    const int r0 = lo / n1, r1 = (lo + len - 1) / n1;
    for (int row = r0 + (lane >> 2); row <= r1; row += 8) {
        const int info = rowinfo[row], i = info & 0xff, j = (info >> 8) & 0xff, kk = info >> 16;   // j = 0xff: obstacle row
        const int base = row * n1 - lo;
        const double dx = dbar[row * 2], dy = dbar[row * 2 + 1];
        for (int a = lane & 3; a <= Hp; a += 4) {
            if (a == Hp) {                                                                         // SCP_controller.py:125
                const int idx = base + n;
                if (idx >= 0 && idx < len) buf[idx] = -1.0;
            } else if (a <= kk) {
                double vi = -2.0 * (dx * gs[(i * Hp + kk - a) * 2] + dy * gs[(i * Hp + kk - a) * 2 + 1]);
                if (fabs(vi) <= 1e-20) vi = 0.0;                                                   // SCP_controller.py:128
                const int ii = base + i * Hp + a;
                if (ii >= 0 && ii < len) buf[ii] = vi;
                if (j != 0xff) {
                    double vj = 2.0 * (dx * gs[(j * Hp + kk - a) * 2] + dy * gs[(j * Hp + kk - a) * 2 + 1]);
                    if (fabs(vj) <= 1e-20) vj = 0.0;
                    const int jj = base + j * Hp + a;
                    if (jj >= 0 && jj < len) buf[jj] = vj;
                }
            }
        }
    }
}

// the chunk [lo, lo + len) of instance-local flat P = blkdiag(2 H, 0) (row-major [n1][n1]) into buf (already zero)
SCP_FN void scp_compose_P(int lane, int Hp, int n1, const double *HB, int lo, int len, double *buf)
{
    const int n = n1 - 1;
    const int r0 = lo / n1;
    int r1 = (lo + len - 1) / n1;
    if (r1 >= n) r1 = n - 1;                                  // the omega row is zero (SCP_controller.py:124)
    for (int row = r0 + (lane >> 2); row <= r1; row += 8) {
        const int v = row / Hp, base = row * n1 - lo + v * Hp;
        for (int bb = lane & 3; bb < Hp; bb += 4) {
            const int idx = base + bb;
            if (idx >= 0 && idx < len) buf[idx] = 2.0 * HB[(size_t)row * Hp + bb];                // SCP_controller.py:120
        }
    }
}

// rowinfo[r] = i | j << 8 | k << 16 (j = 0xff for obstacle rows); built once per kernel, dims only
SCP_FN void scp_asm_rowinfo(Cta &cta, int nVeh, int Hp, int nObst, int *rowinfo)
{
    const int mcv = Hp * (nVeh * (nVeh - 1) / 2), mc = mcv + Hp * nVeh * nObst;
    CTA_PHASE(tid)
        for (int r = tid; r < mc; r += cta.nt) {
            int i, j, o, k;
            scp_row_decode(nVeh, Hp, nObst, mcv, r, &i, &j, &o, &k);
            rowinfo[r] = i | ((o < 0 ? j : 0xff) << 8) | (k << 16);
        }
    CTA_PHASE_END
}

struct ScpAsmMem {
    double *pos, *dbar, *gs, *us, *cs, *ds, *dso, *obs, *bufs;      // staged inputs: g, ubar, cterm, dsafe, (dsafe_obst, obst)
    int *rowinfo;
    int nstage;             // doubles of the staged-input block gs .. obs (contiguous)
};
// doubles of shared memory for nw warps
SCP_HDFN size_t scp_asm_carve(ScpAsmMem &a, double *sh, int nVeh, int Hp, int nObst, int nw)
{
    const int n = nVeh * Hp, mc = Hp * (nVeh * (nVeh - 1) / 2 + nVeh * nObst);
    ScpBump bp = scp_bump(sh, (size_t)1 << 40, (double *)0, true);
    a.bufs = bp.take((size_t)nw * SCP_ASM_NBUF * SCP_ASM_CHUNK);          // first: 16-byte (in fact 128-byte) aligned
    a.pos = bp.take((size_t)n * 2); a.dbar = bp.take((size_t)mc * 2);
    // one contiguous block, every piece an even number of doubles (take() rounds up)
    const size_t s0 = bp.sh_off;
    a.gs = bp.take((size_t)n * 2); a.us = bp.take(n); a.cs = bp.take((size_t)n * 2); a.ds = bp.take((size_t)nVeh * nVeh);
    a.dso = bp.take((size_t)nVeh * nObst); a.obs = bp.take((size_t)nObst * Hp * 2);
    a.nstage = (int)(bp.sh_off - s0);
    a.rowinfo = (int *)bp.take((size_t)(mc + 1) / 2);
    return bp.sh_off;
}

// element e of the staged-input block of instance b, read from global memory
SCP_FN double scp_asm_stage_src(const scpb200_dims &d, int b, int e, const double *g, const double *cterm,
                                const double *ubar, const double *dsafe, const double *dsafe_obst, const double *obst)
{
    const int n = d.nVeh * d.Hp;
    int o = 0, len;
    len = n * 2;            if (e < o + ((len + 1) & ~1)) return e - o < len ? g[(size_t)b * len + (e - o)] : 0.0;       o += (len + 1) & ~1;
    len = n;                if (e < o + ((len + 1) & ~1)) return e - o < len ? ubar[(size_t)b * len + (e - o)] : 0.0;    o += (len + 1) & ~1;
    len = n * 2;            if (e < o + ((len + 1) & ~1)) return e - o < len ? cterm[(size_t)b * len + (e - o)] : 0.0;   o += (len + 1) & ~1;
    len = d.nVeh * d.nVeh;  if (e < o + ((len + 1) & ~1)) return e - o < len ? dsafe[(size_t)b * len + (e - o)] : 0.0;   o += (len + 1) & ~1;
    len = d.nVeh * d.nObst; if (e < o + ((len + 1) & ~1)) return e - o < len ? dsafe_obst[(size_t)b * len + (e - o)] : 0.0; o += (len + 1) & ~1;
    len = d.nObst * d.Hp * 2;
    return e - o < len ? obst[(size_t)b * len + (e - o)] : 0.0;
}

#define SCP_ASM_PF 4        /* staged-input doubles a thread can hold in registers across the store phase */

// Work items first, first + stride, ... < nitems, item = (instance, part): an instance is split into `nparts` items
// over contiguous row ranges of A and P (part 0 also writes q, lb, ub), so that a batch that is not a multiple of the
// resident CTA count still balances.  Per item:
//   phase 1  predicted positions at ubar from the staged inputs            (forward_U, SCP_controller.py:199-213)
//   phase 2  linearisation rows dbar, b (SCP_controller.py:97-114) and the small vectors q, lb, ub, b -> HBM;
//            then every thread fetches its share of the NEXT item's inputs into registers (the loads are in flight
//            while the warps stream this item's chunks, so no item waits on global-memory latency)
//   chunks   per warp, no CTA barrier (see above)
//   phase 3  the prefetched inputs -> the staged block
SCP_FN void scp_assemble_items(Cta &cta, const scpb200_dims &d, const scpb200_params &p, long first, long stride, long nitems,
                               int nparts, const double *g, const double *cterm, const double *H, const double *qv,
                               const double *ubar, const double *dsafe, const double *dsafe_obst, const double *obst,
                               double *P, double *q, double *A, double *bvec, double *lb, double *ub, const ScpAsmMem &sm)
{
    const int nVeh = d.nVeh, Hp = d.Hp, nObst = d.nObst, n = nVeh * Hp, n1 = n + 1;
    const int mcv = Hp * (nVeh * (nVeh - 1) / 2), mc = mcv + Hp * nVeh * nObst;
    double *pos = sm.pos, *dbar = sm.dbar, *gs = sm.gs, *us = sm.us;
    const long NP = (long)n1 * n1, NA = (long)mc * n1;
#if SCP_DEVICE_BUILD
    const bool prefetch = sm.nstage <= SCP_ASM_PF * cta.nt;
    double pf[SCP_ASM_PF];
    unsigned gen = 0;
#else
    const bool prefetch = false;
    unsigned gen = 0;
#endif
    if (first < nitems) {
        CTA_PHASE(tid)
            for (int e = tid; e < sm.nstage; e += cta.nt)
                gs[e] = scp_asm_stage_src(d, (int)(first / nparts), e, g, cterm, ubar, dsafe, dsafe_obst, obst);
        CTA_PHASE_END
    }
    for (long item = first; item < nitems; item += stride) {
        const int b = (int)(item / nparts), part = (int)(item - (long)b * nparts);
        const long nxt = item + stride;
        const int bn = nxt < nitems ? (int)(nxt / nparts) : -1;           // instance of the next item (-1: none)
        const double *HB = H + (size_t)b * n * Hp;
        const int ar0 = (int)((long)mc * part / nparts), ar1 = (int)((long)mc * (part + 1) / nparts);      // rows of A
        const int pr0 = (int)((long)n1 * part / nparts), pr1 = (int)((long)n1 * (part + 1) / nparts);      // rows of P
        scp_positions(cta, nVeh, Hp, gs, sm.cs, us, pos);
        CTA_PHASE(tid)
            for (int r = tid; r < mc; r += cta.nt) {
                const int info = sm.rowinfo[r], i = info & 0xff, j = (info >> 8) & 0xff, k = info >> 16;
                double dx, dy, bx, by, sbar;
                if (j != 0xff) {
                    dx = pos[(i * Hp + k) * 2] - pos[(j * Hp + k) * 2];
                    dy = pos[(i * Hp + k) * 2 + 1] - pos[(j * Hp + k) * 2 + 1];
                    bx = sm.cs[(i * Hp + k) * 2] - sm.cs[(j * Hp + k) * 2];
                    by = sm.cs[(i * Hp + k) * 2 + 1] - sm.cs[(j * Hp + k) * 2 + 1];
                    sbar = sm.ds[i * nVeh + j] + p.dsafeExtra;
                } else {
                    const int o = ((r - mcv) / Hp) % nObst;
                    const double ox = sm.obs[(o * Hp + k) * 2], oy = sm.obs[(o * Hp + k) * 2 + 1];
                    dx = pos[(i * Hp + k) * 2] - ox;
                    dy = pos[(i * Hp + k) * 2 + 1] - oy;
                    bx = sm.cs[(i * Hp + k) * 2] - ox;
                    by = sm.cs[(i * Hp + k) * 2 + 1] - oy;
                    sbar = sm.dso[i * nObst + o] + p.dsafeExtra;
                }
                dbar[r * 2] = dx;
                dbar[r * 2 + 1] = dy;
                if (r >= ar0 && r < ar1)
                    bvec[(size_t)b * mc + r] = -sbar * sbar - (dx * dx + dy * dy) + 2.0 * (dx * bx + dy * by);
            }
            if (part == 0) {
                double *qB = q + (size_t)b * n1, *lbB = lb + (size_t)b * n1, *ubB = ub + (size_t)b * n1;
                for (int c = tid; c < n1; c += cta.nt) {
                    double lo = -p.uLim, hi = p.uLim;
                    if (c < n) {
                        if (p.trust_radius < 1e300) { lo = fmax(lo, us[c] - p.trust_radius); hi = fmin(hi, us[c] + p.trust_radius); }
                        qB[c] = qv[(size_t)b * n + c];
                    } else { lo = 0.0; hi = p.omega_ub; qB[c] = p.omega_weight; }
                    lbB[c] = lo;
                    ubB[c] = hi;
                }
            }
#if SCP_DEVICE_BUILD
            if (prefetch && bn >= 0 && bn != b) {
#pragma unroll
                for (int s = 0; s < SCP_ASM_PF; ++s) {
                    const int e = tid + s * cta.nt;
                    pf[s] = e < sm.nstage ? scp_asm_stage_src(d, bn, e, g, cterm, ubar, dsafe, dsafe_obst, obst) : 0.0;
                }
            }
#endif
        CTA_PHASE_END
        // the two flat ranges of this item, in elements of the whole P / A arrays (16-byte alignment is a property of
        // the global element index: odd leading / trailing elements go out as plain stores)
        const long lo0 = (long)b * NP + (long)pr0 * n1, hi0 = (long)b * NP + (long)pr1 * n1;
        const long lo1 = (long)b * NA + (long)ar0 * n1, hi1 = (long)b * NA + (long)ar1 * n1;
        const int h0 = (int)(lo0 & 1), h1 = (int)(lo1 & 1);
        const int nc0 = hi0 > lo0 ? h0 + (int)((hi0 - lo0 - h0 + SCP_ASM_CHUNK - 1) / SCP_ASM_CHUNK) : 0;
        const int nc1 = hi1 > lo1 ? h1 + (int)((hi1 - lo1 - h1 + SCP_ASM_CHUNK - 1) / SCP_ASM_CHUNK) : 0;
        WARP_SECTION(w, nw)
            for (int t = w; t < nc0 + nc1; t += nw) {
                const bool isA = t >= nc0;
                const int tt = isA ? t - nc0 : t, h = isA ? h1 : h0;
                const long lo = isA ? lo1 : lo0, hi = isA ? hi1 : hi0;
                long c0;
                int len;
                if (h && tt == 0) { c0 = lo; len = 1; }
                else {
                    c0 = lo + h + (long)(tt - h) * SCP_ASM_CHUNK;
                    len = (int)(hi - c0 < SCP_ASM_CHUNK ? hi - c0 : SCP_ASM_CHUNK);
                }
                double *buf = sm.bufs + ((size_t)w * SCP_ASM_NBUF + (gen % SCP_ASM_NBUF)) * SCP_ASM_CHUNK;
                double *dst = (isA ? A : P) + c0;
                const int local = (int)(c0 - (long)b * (isA ? NA : NP));
                WARP_PHASE(lane)
#if SCP_DEVICE_BUILD
                    // the bulk store issued from this buffer SCP_ASM_NBUF chunks ago has finished reading it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(SCP_ASM_NBUF - 1) : "memory");
#endif
                WARP_PHASE_END
                WARP_PHASE(lane)
#if SCP_DEVICE_BUILD
                    double2 *b2 = reinterpret_cast<double2 *>(buf);
                    for (int e = lane; e < ((len + 1) >> 1); e += 32) b2[e] = make_double2(0.0, 0.0);
#else
                    for (int e = lane; e < len; e += 32) buf[e] = 0.0;
#endif
                WARP_PHASE_END
                WARP_PHASE(lane)
                    if (isA) scp_compose_A(lane, nVeh, Hp, n1, sm.rowinfo, gs, dbar, local, len, buf);
                    else scp_compose_P(lane, Hp, n1, HB, local, len, buf);
#if SCP_DEVICE_BUILD
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy
#endif
                WARP_PHASE_END
                WARP_PHASE(lane)
#if SCP_DEVICE_BUILD
                    if (lane == 0) {
                        const int even = len & ~1;
                        if (even)
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                                         "r"((unsigned)__cvta_generic_to_shared(buf)), "r"((unsigned)even * 8u)
                                         : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");     // one group per chunk, empty or not
                        if (len & 1) dst[len - 1] = buf[len - 1];
                    }
#else
                    for (int e = lane; e < len; e += 32) dst[e] = buf[e];
#endif
                WARP_PHASE_END
                ++gen;
            }
        WARP_SECTION_END
        CTA_SYNC                                // every warp is done composing from gs / dbar
        if (bn >= 0 && bn != b) {
            CTA_PHASE(tid)
#if SCP_DEVICE_BUILD
                if (prefetch) {
#pragma unroll
                    for (int s = 0; s < SCP_ASM_PF; ++s) {
                        const int e = tid + s * cta.nt;
                        if (e < sm.nstage) gs[e] = pf[s];
                    }
                } else
#endif
                for (int e = tid; e < sm.nstage; e += cta.nt)
                    gs[e] = scp_asm_stage_src(d, bn, e, g, cterm, ubar, dsafe, dsafe_obst, obst);
            CTA_PHASE_END
        }
    }
}
