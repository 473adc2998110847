// ops_dense.cuh — constraint operator over the reference's own dense QP data (P, Aineq in HBM), used by the
// CVXOPT/Gurobi-replacement entry scpb200_qp_solve_dense (SCP_controller.py:135-150 on the inputs :93-128 builds).
// Straightforward global-memory loops: this entry exists for drop-in use on arbitrary dense QPs and for
// QP-level parity tests; the fused SCP kernel uses ops_pair.cuh instead.
#pragma once
#include "scp_common.cuh"

struct DenseOp {
    int n1, mc;
    const double *P;   // [n1][n1]  (global)
    const double *A;   // [mc][n1]  (global)

    SCP_MFN void mul_P(Cta &cta, const double *x, double *y) const
    {
        CTA_PHASE(tid)
            for (int c = tid; c < n1; c += cta.nt) {
                double acc = 0.0;
                for (int j = 0; j < n1; ++j) acc += P[(size_t)j * n1 + c] * x[j];   // P symmetric: column walk is coalesced
                y[c] = acc;
            }
        CTA_PHASE_END
    }

    SCP_MFN void add_P(Cta &cta, double *S) const
    {
        CTA_PHASE(tid)
            const int tot = n1 * (n1 + 1) >> 1;
            for (int e = tid; e < tot; e += cta.nt) {
                int ci, cj;
                scp_tri_decode(e, &ci, &cj);
                S[scp_sidx(ci, cj)] += 0.5 * (P[(size_t)ci * n1 + cj] + P[(size_t)cj * n1 + ci]);
            }
        CTA_PHASE_END
    }

    SCP_MFN void mul_A(Cta &cta, const double *x, double *y) const
    {
        CTA_PHASE(tid)
            for (int r = tid; r < mc; r += cta.nt) {
                const double *Ar = A + (size_t)r * n1;
                double acc = 0.0;
                for (int c = 0; c < n1; ++c) acc += Ar[c] * x[c];
                y[r] = acc;
            }
        CTA_PHASE_END
    }

    SCP_MFN void add_At(Cta &cta, const double *w, double *vout) const
    {
        CTA_PHASE(tid)
            for (int c = tid; c < n1; c += cta.nt) {
                double acc = 0.0;
                for (int r = 0; r < mc; ++r) acc += A[(size_t)r * n1 + c] * w[r];
                vout[c] += acc;
            }
        CTA_PHASE_END
    }

    SCP_MFN void add_AtDA(Cta &cta, const double *dd, double *S) const
    {
        CTA_PHASE(tid)
            const int tot = n1 * (n1 + 1) >> 1;
            for (int e = tid; e < tot; e += cta.nt) {
                int ci, cj;
                scp_tri_decode(e, &ci, &cj);
                double acc = 0.0;
                for (int r = 0; r < mc; ++r) {
                    const double a = A[(size_t)r * n1 + ci];
                    if (a != 0.0) acc += dd[r] * a * A[(size_t)r * n1 + cj];
                }
                S[scp_sidx(ci, cj)] += acc;
            }
        CTA_PHASE_END
    }
};
