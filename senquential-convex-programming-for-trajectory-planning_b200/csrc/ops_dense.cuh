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

    const double *xp, *wp;   // the x / w of the last prep

    SCP_MFN double P_col(int c, const double *x) const
    {
        double acc = 0.0;
        for (int j = 0; j < n1; ++j) acc += P[(size_t)j * n1 + c] * x[j];          // P symmetric: column walk is coalesced
        return acc;
    }
    SCP_MFN void prep(Cta &cta, const double *x, const double *w)
    {
        (void)cta;
        xp = x;
        wp = w;
    }
    SCP_MFN double row_dot(int r) const
    {
        const double *Ar = A + (size_t)r * n1;
        double acc = 0.0;
        for (int c = 0; c < n1; ++c) acc += Ar[c] * xp[c];
        return acc;
    }
    SCP_MFN double col_dot(int c) const
    {
        double acc = 0.0;
        for (int r = 0; r < mc; ++r) acc += A[(size_t)r * n1 + c] * wp[r];
        return acc;
    }

    // S(lower) = P + A' diag(dd) A + diag(dg): clear, then accumulate
    template <class Mem, class DD, class DG>
    SCP_MFN void form_normal(Cta &cta, const Mem &m, double *dd, double *dg, const double *rhs_row, DD ddf, DG dgf SCP_TIMER_ARG)
    {
        CTA_PHASE(tid)
            const int tot = (m.T * (m.T + 1) >> 1) * SCP_TILE2;
            for (int e = tid; e < tot; e += cta.nt) m.S[e] = 0.0;
            for (int r = tid; r < mc; r += cta.nt) dd[r] = ddf(r);
            for (int c = tid; c < m.n1p; c += cta.nt) dg[c] = dgf(c);
        CTA_PHASE_END
        CTA_PHASE(tid)
            for (int c = tid; c < m.n1p; c += cta.nt) m.S[scp_sidx(c, c)] = c < n1 ? dg[c] : (c == m.n1p - 1 ? 1e300 : 1.0);
            // the right-hand side rides as the last row (see chol_factor)
            if (rhs_row)
                for (int j = tid; j < n1; j += cta.nt) m.S[scp_sidx(m.n1p - 1, j)] = rhs_row[j];
        CTA_PHASE_END
        add_P(cta, m.S);
        add_AtDA(cta, dd, m.S);
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
