// tests/emu/emu.cpp — kernel-logic emulator.  TEST INFRASTRUCTURE ONLY.
//
// Compiles the SAME device source the CUDA library is built from (csrc/*.cuh) with g++: a CTA becomes a loop
// over thread ids per phase (ascending or descending — differing results expose intra-phase races), shared
// memory becomes a heap buffer.  It lets the CPU test-suite (-m "not gpu") exercise the indexing, phase
// structure and numerics of the kernels before they are run on a B200.  It is built into tests/emu/ by the
// tests themselves, is never installed, and nothing in the product package can load it: the product's only
// compute path is libscpb200.so on a CUDA device.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../senquential-convex-programming-for-trajectory-planning_b200/csrc/scp_kernels.cuh"

static int g_nt = 128, g_reverse = 0, g_force_global_S = 0, g_alpha_slots = -1;

extern "C" void emu_config(int nt, int reverse, int force_global_S, int alpha_slots)
{
    g_nt = nt;
    g_reverse = reverse;
    g_force_global_S = force_global_S;
    g_alpha_slots = alpha_slots;          // -1: one scratch slot per warp
}

static Cta *new_cta()
{
    Cta *c = (Cta *)calloc(1, sizeof(Cta));
    c->nt = g_nt;
    c->reverse = g_reverse;
    return c;
}

extern "C" int emu_mpc_setup(const scpb200_dims *d, const scpb200_params *p, const double *x0, const double *u0,
                             const double *veh, const double *poly, double *ref, double *g, double *cterm, double *H,
                             double *qv, double *gamma0, double *abe, int32_t *setup_status)
{
    Cta *cta = new_cta();
    std::vector<double> red(SCP_RED_DOUBLES);
    int flag = 0;
    for (int b = 0; b < d->B; ++b)
        scp_setup_instance(*cta, *d, *p, b, x0, u0, veh, poly, ref, g, cterm, H, qv, gamma0, abe, setup_status,
                           red.data(), &flag);
    free(cta);
    return 0;
}

extern "C" int emu_assemble_dense(const scpb200_dims *d, const scpb200_params *p, const double *g, const double *cterm,
                                  const double *H, const double *qv, const double *ubar, const double *dsafe,
                                  const double *dsafe_obst, const double *obst, double *P, double *q, double *A,
                                  double *b, double *lb, double *ub)
{
    Cta *cta = new_cta();
    const int n = d->nVeh * d->Hp, mc = d->Hp * (d->nVeh * (d->nVeh - 1) / 2 + d->nVeh * d->nObst);
    (void)n;
    ScpAsmMem am;
    const size_t nd = scp_asm_carve(am, (double *)0, d->nVeh, d->Hp, d->nObst, cta->nt / 32);
    std::vector<double> sh(nd + 32, 1e300);                 // poisoned: every element written must come from the composer
    scp_asm_carve(am, sh.data(), d->nVeh, d->Hp, d->nObst, cta->nt / 32);
    scp_asm_rowinfo(*cta, d->nVeh, d->Hp, d->nObst, am.rowinfo);
    const int nparts = mc >= 16 ? 2 : 1;
    // two interleaved "CTAs" (stride 2) so that items of different instances follow each other, as on the device
    for (int first = 0; first < 2; ++first)
        scp_assemble_items(*cta, *d, *p, first, 2, (long)d->B * nparts, nparts, g, cterm, H, qv, ubar, dsafe, dsafe_obst, obst,
                           P, q, A, b, lb, ub, am);
    free(cta);
    return 0;
}

extern "C" int emu_qcqp_evaluate(const scpb200_dims *d, const scpb200_params *p, const double *g, const double *cterm,
                                 const double *H, const double *qv, const double *gamma0, const double *u,
                                 const double *dsafe, const double *dsafe_obst, const double *obst, double *obj,
                                 double *max_violation, double *sum_violations, int32_t *feasible, double *ci,
                                 double *ci_obst)
{
    Cta *cta = new_cta();
    const int nVeh = d->nVeh, Hp = d->Hp, nObst = d->nObst, n = nVeh * Hp;
    std::vector<double> pos((size_t)n * 2), red(SCP_RED_DOUBLES);
    for (int b = 0; b < d->B; ++b) {
        ScpEval ev;
        scp_evaluate(*cta, nVeh, Hp, nObst, g + (size_t)b * n * 2, cterm + (size_t)b * n * 2, H + (size_t)b * n * Hp,
                     qv + (size_t)b * n, gamma0[b], u + (size_t)b * n, dsafe + (size_t)b * nVeh * nVeh,
                     nObst ? dsafe_obst + (size_t)b * nVeh * nObst : 0, nObst ? obst + (size_t)b * nObst * Hp * 2 : 0,
                     p->dsafeExtra, p->constraint_tol, p->obstacle_eval_mode, pos.data(), red.data(), &ev,
                     ci ? ci + (size_t)b * nVeh * nVeh * Hp : 0, ci_obst ? ci_obst + (size_t)b * nVeh * nObst * Hp : 0);
        obj[b] = ev.obj;
        max_violation[b] = ev.max_violation;
        sum_violations[b] = ev.sum_violations;
        feasible[b] = ev.feasible;
    }
    free(cta);
    return 0;
}

extern "C" int emu_qp_solve_dense(const scpb200_dims *d, const scpb200_params *p, int32_t n1, int32_t mc,
                                  const double *P, const double *q, const double *A, const double *b, const double *lb,
                                  const double *ub, double *x, double *fval, int32_t *iters, int32_t *status, double *zA)
{
    Cta *cta = new_cta();
    size_t shu, glu;
    const size_t lim = g_force_global_S ? 2000 : ((size_t)1 << 40);
    ipm_footprint(n1, mc, lim, &shu, &glu);
    std::vector<double> smem(shu + 2), gmem(glu + 2);
    ScpBump bp = scp_bump(smem.data(), lim, gmem.data(), false);
    IpmMem m;
    ipm_carve(bp, m, n1, mc);
    ipm_carve_big(bp, m);
    QpIO io = {P, q, A, b, lb, ub, x, fval, zA, iters, status};
    for (int bi = 0; bi < d->B; ++bi) qp_solve_instance(*cta, *p, n1, mc, bi, io, m);
    free(cta);
    return 0;
}

extern "C" int emu_scp_solve_rate(const scpb200_dims *d, const scpb200_params *p, const double *g, const double *cterm,
                                  const double *H, const double *qv, const double *gamma0, const double *dsafe,
                                  const double *dsafe_obst, const double *obst, double *u_inout, double *traj, double *U,
                                  double *log, int32_t *scp_iters, int32_t *ipm_iters, int32_t *status, double *obj,
                                  double *max_violation, const double *u_prev);

extern "C" int emu_scp_solve(const scpb200_dims *d, const scpb200_params *p, const double *g, const double *cterm,
                             const double *H, const double *qv, const double *gamma0, const double *dsafe,
                             const double *dsafe_obst, const double *obst, double *u_inout, double *traj, double *U,
                             double *log, int32_t *scp_iters, int32_t *ipm_iters, int32_t *status, double *obj,
                             double *max_violation)
{
    return emu_scp_solve_rate(d, p, g, cterm, H, qv, gamma0, dsafe, dsafe_obst, obst, u_inout, traj, U, log, scp_iters, ipm_iters,
                              status, obj, max_violation, (const double *)0);
}

// as scpb200_scp_solve_rate: with p->enable_rate_rows the steering-rate rows anchored at u_prev[B][nVeh]
extern "C" int emu_scp_solve_rate(const scpb200_dims *d, const scpb200_params *p, const double *g, const double *cterm,
                                  const double *H, const double *qv, const double *gamma0, const double *dsafe,
                                  const double *dsafe_obst, const double *obst, double *u_inout, double *traj, double *U,
                                  double *log, int32_t *scp_iters, int32_t *ipm_iters, int32_t *status, double *obj,
                                  double *max_violation, const double *u_prev)
{
    const int rate = p->enable_rate_rows != 0;
    if (rate && !u_prev) return -1;
    Cta *cta = new_cta();
    const int slots = g_alpha_slots < 0 ? g_nt / 32 : g_alpha_slots;
    size_t shu, glu;
    const size_t lim = g_force_global_S ? 3000 : ((size_t)1 << 40);
    scp_footprint(d->nVeh, d->Hp, d->nObst, slots, 1, lim, &shu, &glu, SCP_RED_DOUBLES, rate);
    std::vector<double> smem(shu + 2), gmem(glu + 2);
    ScpBump bp = scp_bump(smem.data(), lim, gmem.data(), false);
    ScpMem s;
    scp_carve(bp, s, d->nVeh, d->Hp, d->nObst, slots, 1, SCP_RED_DOUBLES, rate);
    ScpIO io = {g, cterm, H, qv, gamma0, dsafe, dsafe_obst, obst, u_inout, traj, U, log, obj, max_violation,
                scp_iters, ipm_iters, status};
    io.u_prev = u_prev;
    const size_t snapw = ipm_snap_doubles(s.ipm.n1p, s.ipm.mc, s.ipm.nr);
    std::vector<double> snapbuf((size_t)d->B * snapw, 0.0);
    io.snap = snapbuf.data();
    // SCPB200_EMU_QUANTUM=q: exercise the park / resume path of the work-queue scheduler (q SCP iterations per
    // invocation, round-robin over the live instances as the device FIFO does)
    const char *qe = getenv("SCPB200_EMU_QUANTUM");
    if (qe && atoi(qe) > 0) {
        io.quantum = atoi(qe);
        io.state = (double *)calloc((size_t)d->B * SCP_STATE_W, sizeof(double));
        char *done = (char *)calloc(d->B, 1);
        // SCPB200_EMU_POISON=1: every invocation starts from a working set full of NaNs, as a CTA that last worked on
        // another instance would (nothing an invocation reads may be left over from the one before)
        const char *pe = getenv("SCPB200_EMU_POISON");
        const bool poison = pe && atoi(pe) > 0;
        for (int live = d->B; live > 0;)
            for (int b = 0; b < d->B; ++b) {
                if (done[b]) continue;
                if (poison) {
                    for (auto &v : smem) v = NAN;
                    for (auto &v : gmem) v = NAN;
                }
                if (scp_solve_instance(*cta, *d, *p, b, io, s)) { done[b] = 1; --live; }
            }
        free(done);
        free(io.state);
    } else {
        for (int b = 0; b < d->B; ++b) scp_solve_instance(*cta, *d, *p, b, io, s);
    }
    free(cta);
    return 0;
}

// scpb200_mpc_rollout on the host: per instance, step after step, the same device functions the kernel calls
// (scp_rollout_setup -> scp_solve_instance -> scp_rollout_advance)
extern "C" int emu_mpc_rollout(const scpb200_dims *d, const scpb200_params *p, int nsteps, int mode, const double *veh,
                               const double *poly, const double *dsafe, double *x0, double *u0, double *x_meas, double *u_act,
                               double *ref, double *g, double *cterm, double *H, double *qv, double *gamma0, double *abe,
                               double *u, double *traj, double *U, int32_t *scp_iters, int32_t *ipm_iters, int32_t *status,
                               double uMax, double duLim, double mech_limit, double lat_acc_limit, double delay, int nsub_delay,
                               int nsub_plant, int32_t *qp_total, int32_t *ipm_total, int32_t *status_or, int32_t *scp_iters_hist,
                               int32_t *status_hist, double *U_hist, double *x_hist)
{
    Cta *cta = new_cta();
    const int slots = g_alpha_slots < 0 ? g_nt / 32 : g_alpha_slots;
    size_t shu, glu;
    const size_t lim = (size_t)1 << 40;
    scp_footprint(d->nVeh, d->Hp, d->nObst, slots, 1, lim, &shu, &glu);
    std::vector<double> smem(shu + 2), gmem(glu + 2), obj(d->B), mv(d->B);
    ScpBump bp = scp_bump(smem.data(), lim, gmem.data(), false);
    ScpMem s;
    scp_carve(bp, s, d->nVeh, d->Hp, d->nObst, slots, 1);
    ScpIO io = {g, cterm, H, qv, gamma0, dsafe, (const double *)0, (const double *)0, u, traj, U, (double *)0, obj.data(), mv.data(),
                scp_iters, ipm_iters, status};
    const size_t snapw = ipm_snap_doubles(s.ipm.n1p, s.ipm.mc);
    std::vector<double> snapbuf((size_t)d->B * snapw, 0.0);
    io.snap = snapbuf.data();
    ScpRollout ro;
    memset(&ro, 0, sizeof ro);
    ro.nsteps = nsteps; ro.mode = mode; ro.counter0 = p->noise_counter; ro.veh = veh; ro.poly = poly; ro.x0 = x0; ro.u0 = u0;
    ro.x_meas = x_meas; ro.u_act = u_act; ro.ref = ref; ro.g = g; ro.cterm = cterm; ro.H = H; ro.qv = qv; ro.gamma0 = gamma0;
    ro.abe = abe; ro.uMax = uMax; ro.duLim = duLim; ro.mech_limit = mech_limit; ro.lat_acc_limit = lat_acc_limit; ro.delay = delay;
    ro.nsub_delay = nsub_delay; ro.nsub_plant = nsub_plant; ro.qp_total = qp_total; ro.ipm_total = ipm_total; ro.status_or = status_or;
    ro.scp_iters_hist = scp_iters_hist; ro.status_hist = status_hist; ro.U_hist = U_hist; ro.x_hist = x_hist;
    int flag = 0;
    for (int b = 0; b < d->B; ++b)
        for (int step = 0; step < nsteps; ++step) {
            scp_rollout_setup(*cta, *d, *p, ro, b, step, s.resp, s.ipm.red, &flag, (double *)0, 1);
            scp_solve_instance(*cta, *d, *p, b, io, s);
            scp_rollout_advance(*cta, *d, *p, ro, io, b, step);
        }
    free(cta);
    return 0;
}

extern "C" int emu_ode_predict(const scpb200_dims *d, const scpb200_params *p, const double *x, const double *u_ref,
                               const double *veh, double T, int32_t steps, int32_t nsub, double *out)
{
    for (int e = 0; e < d->B * d->nVeh; ++e)
        scp_ode_predict_vehicle(x + (size_t)e * 6, u_ref[e], veh[(size_t)e * 5], veh[(size_t)e * 5 + 1], T, steps, nsub,
                                p->noise_sigma, p->seed, p->instance0 + (uint32_t)(e / d->nVeh), (uint32_t)(e % d->nVeh),
                                p->noise_counter, out + (size_t)e * steps * 6);
    return 0;
}

extern "C" int emu_plant_step(const scpb200_dims *d, const scpb200_params *p, const double *veh, const double *U,
                              double mech_limit, double lat_acc_limit, double duLim, double T, int32_t nsub, double *x_meas,
                              double *u_act, double *u_max_out, double *U_clamped)
{
    for (int e = 0; e < d->B * d->nVeh; ++e) {
        const int b = e / d->nVeh, v = e % d->nVeh;
        scp_plant_step_vehicle(x_meas + (size_t)e * 6, u_act + e, U + (size_t)b * d->Hp * d->nVeh + v,
                               U_clamped ? U_clamped + (size_t)b * d->Hp * d->nVeh + v : (double *)0, d->Hp, d->nVeh,
                               veh[(size_t)e * 5], veh[(size_t)e * 5 + 1], mech_limit, lat_acc_limit, duLim, T, nsub,
                               p->noise_sigma, p->seed, p->instance0 + (uint32_t)b, (uint32_t)v, p->noise_counter,
                               u_max_out ? u_max_out + e : (double *)0);
    }
    return 0;
}

extern "C" size_t emu_scp_shared_bytes(int nVeh, int Hp, int nObst, int S_in_shared)
{
    size_t shu, glu;
    scp_footprint(nVeh, Hp, nObst, g_alpha_slots < 0 ? g_nt / 32 : g_alpha_slots, 1, S_in_shared ? ((size_t)1 << 40) : 3000, &shu, &glu);
    return shu * 8;
}
