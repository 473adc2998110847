/*
 * scpb200.h — C ABI of libscpb200.so: the B200-native (sm_100a) batched replacement for the SCP-QP hot path
 * of Zhang-Xiaoxue/Senquential-Convex-Programming-for-Trajectory-Planning.
 *
 * The reference has no FFI: its boundary for this path is three Python classes (IterClass / MPCclass in
 * MPC_Iter.py:13-149, SCPcontroller in SCP_controller.py:18-400) called from main.py:123-134.  The Python
 * facade of this repo keeps those classes and binds the entry points below with ctypes (INTEGRATION.md shows
 * the stub).  Every entry point cites the reference code it replaces.
 *
 * Conventions
 *   - Every data pointer is a DEVICE pointer to caller-owned, contiguous memory (FP64 unless stated int32),
 *     batch outermost, row-major.  The library never allocates caller-visible memory.
 *   - All work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*); nothing
 *     synchronises.  `ws` is a caller-provided device workspace of scpb200_workspace_bytes() bytes.
 *   - Return value: 0 = OK, <0 = argument / launch error (scpb200_last_error() gives the text).
 *     Per-instance numerical outcomes are DATA (status arrays), not errors.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with SCPB200_ERR_CUDA.
 *
 * Index conventions (SCP_controller.py:295,202): n = nVeh*Hp, u[v*Hp + k] (vehicle-major), n1 = n+1 with the
 * slack omega last (SCP_controller.py:123-127), collision rows ordered (i, j>i, k) then obstacle rows
 * (v, o, k) (SCP_controller.py:97-114), mc = Hp*(nVeh(nVeh-1)/2 + nVeh*nObst).
 */
#ifndef SCPB200_H
#define SCPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCPB200_VERSION 100

#define SCPB200_OK 0
#define SCPB200_ERR_ARG (-1)      /* bad dims / null pointer */
#define SCPB200_ERR_CUDA (-2)     /* CUDA runtime error (no device, launch failure, ...) */
#define SCPB200_ERR_SIZE (-3)     /* problem does not fit this build's limits */

/* per-instance status bits (scp_solve / qp_solve_dense `status` outputs) */
#define SCPB200_ST_QP_MAXITER 1   /* some QP hit ipm_max_iter before meeting the tolerances */
#define SCPB200_ST_QP_PIVOT 2     /* a Cholesky pivot had to be repaired (ill-conditioned normal matrix) */
#define SCPB200_ST_SCP_MAXITER 4  /* SCP loop ended on max_scp_iter, not on the stop test */
#define SCPB200_ST_INFEASIBLE 8   /* final iterate violates a collision constraint by more than constraint_tol */
#define SCPB200_ST_SETUP 16       /* set-up failed for this instance (expm / sampler index) */
#define SCPB200_ST_QP_DRES_FLOOR 32 /* some QP was accepted at the precision floor of its dual residual: gap and primal
                                     * residual within tolerance, dual residual within 100 qp_feastol and no longer
                                     * decreasing (further iterations only degrade it) */

#define SCPB200_ST_QP_WARM_RESTART 64 /* log rows only: the QP's warm start did not converge within qp_warm_max_iter and it
                                     * was solved again from the cold starting point (the row's iteration count is the sum) */

typedef struct scpb200_dims {
    int32_t B;      /* instances (independent scenarios / noise samples) */
    int32_t nVeh;   /* vehicles per instance (Scenarios.py:59) */
    int32_t Hp;     /* prediction = control horizon (Scenarios.py:50-51; the reference asserts Hu <= Hp, uses Hu = Hp) */
    int32_t nObst;  /* obstacles per instance (Scenarios.py:220) */
    int32_t nPts;   /* points per reference polyline (2 in every shipped scenario) */
} scpb200_dims;

/* Constants of the path.  scpb200_default_params() fills in the reference's values. */
typedef struct scpb200_params {
    double dt;               /* MPC sample time, Scenarios.py:49 (0.4) */
    double uLim;             /* steering bound, SCP_controller.py:34 (undefined in the reference; = mechanicalSteeringLimit) */
    double dsafeExtra;       /* Scenarios.py:57 (1.0) */
    double delta_tol;        /* SCP_controller.py:83 (1e-3) */
    double omega_weight;     /* SCP_controller.py:84 (1e5) */
    double omega_ub;         /* SCP_controller.py:85 (1e25; bounds >= inf_bound are treated as absent) */
    double constraint_tol;   /* Config.py:18 (2*2.1*1e-3) */
    int32_t max_scp_iter;    /* SCP_controller.py:86 (20) */
    int32_t obstacle_eval_mode; /* 0: every (v,o,k) once; 1: the reference's nesting (SCP_controller.py:249-263) */
    /* interior-point controls (the reference delegates these to its third-party solver) */
    double qp_abstol, qp_reltol, qp_feastol; /* CVXOPT-style stopping rule; defaults 1e-7, 1e-13, 1e-9: the absolute gap
                              * certifies |u - u*| <= sqrt(2 gap / lambda_min) = 5e-6 (lambda_min = 2 R = 8000) */
    double qp_dual_reg;      /* proximal regularisation of s/z in the normal matrix (default 1e-11: no pivot breakdown up to Hp = 50, fewer iterations than 1e-12, and u within 2e-8 of the extended-precision minimiser where 1e-10 leaves 1e-6; swept 1e-12 ... 1e-9 on B200) */
    double inf_bound;        /* |bound| >= inf_bound means "no bound" (default 1e20, Gurobi's convention) */
    int32_t ipm_max_iter;    /* default 60 */
    int32_t qp_warm_start;   /* scpb200_scp_solve only: start the QP of SCP iteration k+1 from an interior iterate of QP k
                              * (same unique minimiser, fewer interior-point iterations); 0 = every QP from the coneqp
                              * starting point.  Default 1. */
    /* extensions, default = reference behaviour */
    double trust_radius;     /* |u - ubar|_inf <= rho folded into lb/ub; +inf (>= 1e300) = off */
    double noise_sigma;      /* std-dev of the process noise added to f(x,u)[0:2] (Model.py:84-86: 3e-6); 0 = off */
    uint64_t seed;           /* Philox4x32-10 key */
    uint32_t instance0;      /* global index of instance 0 of this call (sharding keeps streams G-independent) */
    uint32_t noise_counter;  /* draw counter (e.g. the MPC step index) */
    double qp_warm_relgap;   /* the iterate kept for the next QP is the first with relative gap <= this (default 1.0: an early, well-centred iterate; measured best on the 1024-instance benchmark) */
    int32_t qp_warm_max_iter; /* a warm-started QP not converged after this many iterations restarts cold (default 30) */
    int32_t qp_warm_min_iter; /* ... and not before this iteration of the QP it is taken from (default 5; swept 1..6 with relgap 0.1..100 on the 1024-instance benchmark, profiles/r01_sweep_warm_start.txt) */
    int32_t qp_warm_carry;   /* 1: the first QP of a call starts from the iterate the previous call on the same workspace
                              * left for that instance (consecutive MPC steps of the same scenarios; the workspace must be
                              * zero-initialised before its first use).  Default 0: every call starts cold. */
    int32_t qp_dres_floor_factor; /* a QP whose gap and primal residual have converged is accepted when its dual residual is
                              * within this factor of qp_feastol and no longer decreasing (SCPB200_ST_QP_DRES_FLOOR);
                              * 0 = never (iterate to ipm_max_iter).  Default 100. */
    /* extension: steering-rate rows inside the QP (the reference only clamps AFTER the solve, main.py:144-174).
     * With enable_rate_rows = 1 every QP carries, per vehicle v and step k, the two rows
     *     u_v[k] - u_v[k-1] <= duLim,   u_v[k-1] - u_v[k] <= duLim        (u_v[-1] = u_prev[v], the command being actuated)
     * (row order of the equivalent dense QP: after the collision / obstacle rows, (v, k, +), (v, k, -); 2 nVeh Hp rows).
     * Inside K4 they are never materialised: two non-zeros per row make them a tridiagonal term of every vehicle block of
     * the normal matrix.  u_prev comes in through scpb200_scp_solve_rate (per-step entry) or is the set-up input u0
     * (scpb200_mpc_rollout).  Default 0 reproduces the reference bit for bit (same kernels, same layout). */
    int32_t enable_rate_rows;
    int32_t log_capacity;    /* rows per instance allocated in the `log` array of scpb200_scp_solve (0: max_scp_iter rows);
                              * a call with max_scp_iter > log_capacity > 0 is rejected instead of writing past the rows */
    double duLim;            /* steering-rate bound per MPC step, Scenarios.py:54 (6 deg); read when enable_rate_rows = 1 */
    uint32_t noise_stream;   /* scpb200_ode_predict only: which of its callers draws (0 = the delay-compensation prediction,
                              * 1 = tick-path predictions, ...).  The set-up, the plant step and every ode_predict caller use
                              * disjoint Philox streams, so that the draws of one MPC step are independent as the reference's
                              * np.random.normal draws are (Model.py:84-86). */
    uint32_t reserved0;
} scpb200_params;

/* width of one row of the per-iteration log of scpb200_scp_solve (SCP_controller.py:169-189, scalar fields) */
#define SCPB200_LOG_W 10
/* log row: {slack, SCP_ObjVal (fval+gamma0), QCQP_ObjVal, delta_hat, delta, feasible, max_violation,
 *           sum_violations, ipm_iterations, qp_status} */

int scpb200_version(void);
const char *scpb200_last_error(void);
void scpb200_default_params(scpb200_params *p);

/* number of CUDA devices visible (0 if none); does not create a context on failure */
int scpb200_device_count(void);

/* bytes of device workspace the solver entry points need for these dims on the current device */
int scpb200_workspace_bytes(const scpb200_dims *d, size_t *bytes);
/* same for scpb200_qp_solve_dense on an arbitrary (n1, mc) */
int scpb200_qp_workspace_bytes(int32_t n1, int32_t mc, size_t *bytes);
/* launch geometry scpb200_scp_solve would use (diagnostics): out[6] = {grid, threads, dynamic shared bytes,
 * normal matrix in shared memory (1/0), SM count, pair-block scratch slots} */
int scpb200_scp_plan(const scpb200_dims *d, int64_t *out);

/*
 * K1 — replaces MPCclass.__init__ (MPC_Iter.py:57-97): comp_jacobian (Model.py:45-59), discretize
 * (MPC_Iter.py:99-113), prediction_matrices (:129-149), const_term (:90), mpc_cost_function_matrices
 * (:116-127), and the reference sampling of IterClass (MPC_Iter.py:35-43 -> SampleReferTraj.py:8-122).
 *   in : x0[B,nVeh,6] u0[B,nVeh] veh[B,nVeh,5]=(Lf,Lr,Q,Q_final,R) poly[B,nVeh,nPts,2]
 *   out: ref[B,nVeh,Hp,2]   sampled reference points (ReferenceTrajectoryPoints[k,:,v])
 *        g[B,nVeh,Hp,2]     impulse response C A^l B      (Mathcal_B[2i:2i+2, j] = g[i-j])
 *        cterm[B,nVeh,Hp,2] free response (const_term)
 *        H[B,nVeh,Hp,Hp]    Phi_0      qv[B,nVeh,Hp] Psi_0      gamma0[B] sum_v gamma_0
 *        abe[B,nVeh,48]     Ad(36) Bd(6) Ed(6), may be NULL   setup_status[B] int32 (0 or SCPB200_ST_SETUP), may be NULL
 */
int scpb200_mpc_setup(const scpb200_dims *d, const scpb200_params *p, const double *x0, const double *u0,
                      const double *veh, const double *poly, double *ref, double *g, double *cterm, double *H,
                      double *qv, double *gamma0, double *abe, int32_t *setup_status, void *stream);

/*
 * K2 (materialised) — replaces QCQP_formulate (SCP_controller.py:278-341) + the row assembly of SCP_optimizer
 * (:93-128): the dense QP the reference hands to its solver, in the reference's layout.
 *   in : g, cterm, H, qv (K1 outputs), ubar[B,n] linearisation point, dsafe[B,nVeh,nVeh],
 *        dsafe_obst[B,nVeh,nObst] and obst[B,nObst,Hp,2] (NULL when nObst == 0)
 *   out: P[B,n1,n1] q[B,n1] A[B,mc,n1] b[B,mc] lb[B,n1] ub[B,n1]
 */
int scpb200_assemble_dense(const scpb200_dims *d, const scpb200_params *p, const double *g, const double *cterm,
                           const double *H, const double *qv, const double *ubar, const double *dsafe,
                           const double *dsafe_obst, const double *obst, double *P, double *q, double *A,
                           double *b, double *lb, double *ub, void *stream);

/*
 * a9 — replaces QCQP_evaluate (SCP_controller.py:215-265) (items 3-4 of its return tuple are unused by every
 * caller and are not produced).   ci[B,nVeh,nVeh,Hp] (symmetric fill, -inf elsewhere) and
 * ci_obst[B,nVeh,nObst,Hp] may be NULL.
 */
int scpb200_qcqp_evaluate(const scpb200_dims *d, const scpb200_params *p, const double *g, const double *cterm,
                          const double *H, const double *qv, const double *gamma0, const double *u,
                          const double *dsafe, const double *dsafe_obst, const double *obst, double *obj,
                          double *max_violation, double *sum_violations, int32_t *feasible, double *ci,
                          double *ci_obst, void *stream);

/*
 * a11 — replaces forward_U (SCP_controller.py:199-213).  traj[B,Hp,2,nVeh], U[B,Hp,nVeh].
 */
int scpb200_forward_u(const scpb200_dims *d, const double *g, const double *cterm, const double *u, double *traj,
                      double *U, void *stream);

/*
 * Delay-compensation prediction of IterClass (MPC_Iter.py:25-33: odeint of Model.py:61-87 over
 * linspace(0, delay_x+dt+delay_u, steps) with the steering reference held at u_path[:, -1]).
 * Classical RK4, nsub substeps per output interval (16 gives < 1e-9 m against the reference's LSODA).
 *   in : x[B,nVeh,6] measured states, u_ref[B,nVeh], veh[B,nVeh,5] (Lf, Lr used), T seconds, steps >= 2
 *   out: out[B,nVeh,steps,6]   (MPC_delay_compensation_trajectory[s,:,v]; the last sample is Iter.x0)
 */
int scpb200_ode_predict(const scpb200_dims *d, const scpb200_params *p, const double *x, const double *u_ref,
                        const double *veh, double T, int32_t steps, int32_t nsub, double *out, void *stream);

/*
 * Test / validation entry: the raw N(0,1) pairs the kernels consume (Philox4x32-10 keyed by seed; counter words
 * instance0 + b, vehicle, counter0 + c, stream tag; Box-Muller).  out[B,nVeh,ncount,2].  Replaces nothing in the
 * reference (np.random.normal, Model.py:84-86, cannot be reproduced in a batched kernel); it exists so that the
 * generator can be checked against Random123's known-answer vectors and statistically (hard part 7).
 */
int scpb200_noise_draws(const scpb200_dims *d, const scpb200_params *p, uint32_t noise_stream, uint32_t counter0,
                        int32_t ncount, double *out, void *stream);

/*
 * The caller's half of one MPC step — replaces main.py:104-109 (dynamic steering limit),
 * :164-174 (clamp of the controller output to +-uMax and the rate limit duLim) and :176-191 (plant integration over
 * one sample time, Model.py:89-115, with the actuation delay of main.py:176-181: during step i the plant runs with
 * the command of step i-1).  The reference integrates with dopri5 at 1e-8; here RK4 with nsub substeps (64 gives
 * < 1e-9 m).  Process noise as in scpb200_ode_predict (p->noise_sigma, Philox keyed by instance / vehicle / counter).
 *   in    : veh[B,nVeh,5] (Lf, Lr used), U[B,Hp,nVeh] controller output (K4), mech_limit, lat_acc_limit, duLim, T = dt
 *   in/out: x_meas[B,nVeh,6] state at the step's first tick -> at its last tick (the next x_measured);
 *           u_act[B,nVeh] command being actuated (u_path[:, -1] = Iter.u0) -> clamped U[0] (the next Iter.u0)
 *   out   : u_max_out[B,nVeh] (Iter.uMax; may be NULL), U_clamped[B,Hp,nVeh] (controlPredictions; may be NULL)
 */
int scpb200_plant_step(const scpb200_dims *d, const scpb200_params *p, const double *veh, const double *U,
                       double mech_limit, double lat_acc_limit, double duLim, double T, int32_t nsub, double *x_meas,
                       double *u_act, double *u_max_out, double *U_clamped, void *stream);

/*
 * Closed-loop advance on the controller's own linear model (MPC_Iter.py:94-97): x0 <- Ad x0 + Bd u + Ed,
 * u0 <- u, with u = U[b,0,v] clamped to |u| <= uMax and |u - u0| <= duLim as main.py:164-168 does.
 * Used by the synthetic benchmark to close the loop between MPC steps without leaving the device.
 *   in : abe[B,nVeh,48] (K1), U[B,Hp,nVeh] (K4)      in/out: x0[B,nVeh,6], u0[B,nVeh]
 */
int scpb200_advance_linear(const scpb200_dims *d, const double *abe, const double *U, double uMax, double duLim,
                           double *x0, double *u0, void *stream);

/*
 * K3 — replaces the third-party QP solve of SCP_controller.py:135-150 on the reference's own dense inputs
 * (the CVXOPT/Gurobi-replacement entry):   min 1/2 x'Px + q'x  s.t.  A x <= b,  lb <= x <= ub.
 *   in : n1, mc and P[B,n1,n1] q[B,n1] A[B,mc,n1] b[B,mc] lb[B,n1] ub[B,n1]   (only d->B is read from dims)
 *   out: x[B,n1] fval[B] (= prob.value) iters[B] int32, status[B] int32, zA[B,mc] multipliers (may be NULL)
 */
int scpb200_qp_solve_dense(const scpb200_dims *d, const scpb200_params *p, int32_t n1, int32_t mc,
                           const double *P, const double *q, const double *A, const double *b, const double *lb,
                           const double *ub, double *x, double *fval, int32_t *iters, int32_t *status, double *zA,
                           void *ws, void *stream);

/*
 * K4 (fused) — replaces SCPcontroller.SCP_controller / SCP_optimizer (SCP_controller.py:40-197): the whole SCP
 * loop on device, one CTA per instance pulled from a work queue: linearise about ubar, solve the QP with the
 * CTA-resident interior-point method on the structured rows, evaluate the true QCQP, merit / stop test.
 *   in : K1 outputs, dsafe, (dsafe_obst, obst), u_inout[B,n] warm start (SCP_controller.py:42-43)
 *   out: u_inout (solution), traj[B,Hp,2,nVeh], U[B,Hp,nVeh] (forward_U shapes),
 *        log[B,max_scp_iter,SCPB200_LOG_W] (may be NULL), scp_iters[B] int32 (= QPs solved), ipm_iters[B] int32,
 *        status[B] int32, obj[B] final QCQP objective, max_violation[B]
 */
int scpb200_scp_solve(const scpb200_dims *d, const scpb200_params *p, const double *g, const double *cterm,
                      const double *H, const double *qv, const double *gamma0, const double *dsafe,
                      const double *dsafe_obst, const double *obst, double *u_inout, double *traj, double *U,
                      double *log, int32_t *scp_iters, int32_t *ipm_iters, int32_t *status, double *obj,
                      double *max_violation, void *ws, void *stream);

/*
 * Scheduling extension (no reference counterpart: the reference solves one scenario at a time).  The SCP / interior-
 * point iteration counts vary by 10x between instances, so the order in which the persistent CTAs pull instances
 * decides how long the batch waits for its last straggler.
 *   scpb200_work_order       order[B] int32 = instance indices sorted by descending work[B] int32 (e.g. the previous MPC
 *                            step's ipm_iters of the same instances); device arrays.
 *   scpb200_scp_solve_ordered  as scpb200_scp_solve, pulling instances in `order` (NULL = natural order).
 * Results are independent of the order (instances are independent).
 */
int scpb200_work_order(int32_t B, const int32_t *work, int32_t *order, void *stream);
int scpb200_scp_solve_ordered(const scpb200_dims *d, const scpb200_params *p, const double *g, const double *cterm,
                              const double *H, const double *qv, const double *gamma0, const double *dsafe,
                              const double *dsafe_obst, const double *obst, double *u_inout, double *traj, double *U,
                              double *log, int32_t *scp_iters, int32_t *ipm_iters, int32_t *status, double *obj,
                              double *max_violation, const int32_t *order, void *ws, void *stream);

/*
 * K4 with steering-rate rows (north-star item 2, "input/rate-bound constraint rows"; no reference counterpart: the
 * reference clamps the solved command afterwards, main.py:164-174, so its QP can ask for a steering step the actuator
 * will not deliver).  As scpb200_scp_solve_ordered, plus
 *   u_prev[B,nVeh]   the command being actuated (Iter.u0, MPC_Iter.py:19): anchors the first rate row of every vehicle.
 * Required when p->enable_rate_rows != 0 (the other two entries then fail with SCPB200_ERR_ARG); ignored otherwise.
 */
int scpb200_scp_solve_rate(const scpb200_dims *d, const scpb200_params *p, const double *g, const double *cterm,
                           const double *H, const double *qv, const double *gamma0, const double *dsafe,
                           const double *dsafe_obst, const double *obst, double *u_inout, double *traj, double *U,
                           double *log, int32_t *scp_iters, int32_t *ipm_iters, int32_t *status, double *obj,
                           double *max_violation, const int32_t *order, const double *u_prev, void *ws, void *stream);

/*
 * Rollout entry (north-star item 4: "the SCP ... loop kept on-device with noise injection for Monte-Carlo rollouts") —
 * replaces the caller's loop main.py:98-191 for a whole batch: `nsteps` closed-loop MPC steps per instance in ONE launch.
 * Scenarios / noise samples are independent across MPC steps as well, so a CTA that finishes an instance's SCP loop
 * closes the loop for it and re-queues it for its next step; no step of the batch waits for its slowest instance.  Per
 * instance and step the arithmetic is exactly that of
 *     mode 0:  scpb200_mpc_setup -> scpb200_scp_solve -> scpb200_advance_linear          (the benchmark's loop closure)
 *     mode 1:  scpb200_ode_predict (delay compensation) -> scpb200_mpc_setup -> scpb200_scp_solve -> scpb200_plant_step
 * with noise_counter = p->noise_counter + step; results are bit-identical to those call sequences.
 * All pointers are device pointers with the layouts of the per-step entry points; arrays marked (work) are per-instance
 * scratch that holds the last step's values on return.
 */
typedef struct scpb200_rollout {
    int32_t nsteps;          /* MPC steps per instance (>= 1) */
    int32_t mode;            /* 0 or 1, see above */
    const double *veh, *poly, *dsafe, *dsafe_obst, *obst;       /* scenario (read-only) */
    double *x0, *u0;         /* [B,nVeh,6], [B,nVeh]: mode 0: state, advanced in place; mode 1: (work) delay-compensated state */
    double *x_meas, *u_act;  /* mode 1: measured state / command being actuated, advanced in place (NULL in mode 0) */
    double *ref, *g, *cterm, *H, *qv, *gamma0, *abe;            /* (work) set-up outputs */
    int32_t *setup_status;   /* (work) may be NULL */
    double *u;               /* [B,n] warm start of step 0 in, last solution out */
    double *traj, *U, *obj, *max_violation;                      /* (work) last step's results; traj/obj/max_violation may be NULL */
    int32_t *scp_iters, *ipm_iters, *status;                     /* (work) last step's counters */
    double uMax, duLim;      /* clamp of the applied command (mode 0: |u| <= uMax; both modes: |u - u_prev| <= duLim) */
    double mech_limit, lat_acc_limit, delay;                     /* mode 1: as scpb200_plant_step / the horizon of scpb200_ode_predict */
    int32_t nsub_delay, nsub_plant;                              /* mode 1: RK4 substeps of the two integrations */
    int32_t *qp_total, *ipm_total, *status_or;                   /* [B] QPs solved / interior-point iterations / OR of the status over the steps */
    int32_t *scp_iters_hist, *status_hist;                       /* [B,nsteps] per-step records, may be NULL */
    double *U_hist;          /* [B,nsteps,Hp,nVeh] controller outputs per step, may be NULL */
    double *x_hist;          /* [B,nsteps+1,nVeh,6] state before step 0 and after every step, may be NULL */
} scpb200_rollout;

int scpb200_mpc_rollout(const scpb200_dims *d, const scpb200_params *p, const scpb200_rollout *r, void *ws, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SCPB200_H */
