// scp_solve_kernel.cuh — the persistent SCP kernel (K4) as a template over the residency of the working set and,
// optionally, literal problem dimensions and CTA width.  The library instantiates it in separate translation units
// (scp_solve_generic.cu: run-time dimensions; scp_solve_fixed.cu: the shapes BASELINE.json names, once per CTA
// width) which nvcc compiles in parallel; each unit exports a ScpKernelEntry through which the host code in
// scpb200.cu queries occupancy and launches, so no device symbol crosses a unit boundary.
#pragma once
#include "scp_kernels.cuh"

#ifndef SCP_FIXED_ALPHA
#define SCP_FIXED_ALPHA 1      /* pair-block mode of the fixed-shape instantiations (tuning: 0 = entry by entry) */
#endif
#ifndef SCP_MIN_CTAS
#define SCP_MIN_CTAS 2
#endif

// ---- work queue of the SCP kernel -------------------------------------------------------------------------------
// The unit of scheduling is ONE QP (one SCP iteration of one instance), not one instance: instances need between 1 and
// max_scp_iter QPs, so with instance-granular scheduling a 1024-instance step waits for stragglers that started late.
// A bounded FIFO ring in the workspace holds the instances that still have work; a CTA pops one, runs `quantum` SCP
// iterations, and either finishes it or parks it (u + five scalars) and pushes it back at the tail.  Every live
// instance therefore advances round-robin and the CTAs stay busy until the last QP of the step.
//   hdr[0] = head ticket, hdr[1] = tail ticket, hdr[2] = instances not yet finished; slots[cap], cap = power of two
//   >= 2B, empty = -1.  Pop ticket h is served by push ticket h (FIFO); a popper whose ticket is never served leaves
//   when hdr[2] reaches 0.
// Rollout entry only: hdr[3] = the lowest MPC step any instance is still at, hdr[SCP_Q_CNT + s] = instances at step s
// (s < SCP_Q_STEPS; longer rollouts run without the laggard rule below).
#define SCP_Q_CNT 64
#define SCP_Q_STEPS 4096
struct WorkQueue {
    int *hdr, *slots;
    int cap;
};

// everything a launch of the kernel needs, passed by value
struct ScpKernelArgs {
    scpb200_dims d;
    scpb200_params p;
    ScpIO io;
    WorkQueue q;
    double *gws;
    size_t gl_stride, sh_lim;
    int alpha_slots, want_H;
    ScpRollout ro;             // ro.nsteps > 0: the rollout entry (scpb200_mpc_rollout)
};

// what a translation unit exports per kernel instantiation
struct ScpKernelEntry {
    // sets the dynamic shared-memory limit of the kernel and returns its resident CTAs per SM for this launch shape
    int (*prepare)(int threads, size_t smem_bytes, int *ctas_per_sm);
    int (*launch)(int grid, int threads, size_t smem_bytes, void *stream, const ScpKernelArgs *a);
    // tuning builds (-DSCP_PHASE_TIMERS): read and clear this unit's per-region cycle counters; null otherwise
    int (*read_timers)(unsigned long long *out32);
    // CTA-width bound this unit was compiled for (SCP_MAX_THREADS of the unit): sizes the reduction scratch in the layout
    int max_threads;
};

#if SCP_DEVICE_BUILD
__device__ __forceinline__ void queue_push(const WorkQueue &q, int b)
{
    const int t = atomicAdd(q.hdr + 1, 1);
    int *p = q.slots + (t & (q.cap - 1));
    while (atomicCAS(p, -1, b) != -1) __nanosleep(64);
}

// called by thread 0; returns an instance index, or -1 when every instance has finished
__device__ __forceinline__ int queue_pop(const WorkQueue &q)
{
    const int h = atomicAdd(q.hdr, 1);
    int *p = q.slots + (h & (q.cap - 1));
    unsigned ns = 32;
    for (;;) {
        const int v = atomicExch(p, -1);
        if (v >= 0) return v;
        if (*(volatile int *)(q.hdr + 2) <= 0) return -1;
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
    }
}

// One invocation of the rollout entry on instance b: (set-up if the instance is at the start of an MPC step) -> one
// quantum of its SCP loop -> (if the loop finished: record, close the loop, next step).  Returns true when the instance
// has run all its MPC steps.  k1ws / k1_warps: shared scratch for the warp-level set-up (the normal-matrix area, dead
// between QPs).
__device__ __forceinline__ bool rollout_invocation(Cta &cta, const scpb200_dims &d, const ScpKernelArgs &a, int b, ScpMem &s,
                                                   double *k1ws, int k1_warps, int *setup_flag, int pin)
{
    double *stB = a.io.state + (size_t)b * SCP_STATE_W;
    const int it_resume = (int)SCP_LD_COHERENT(stB + 2), step = (int)SCP_LD_COHERENT(stB + 7);
    if (it_resume == 0) {
        scp_rollout_setup(cta, d, a.p, a.ro, b, step, s.resp, s.ipm.red, setup_flag, k1ws, k1_warps);
        __syncthreads();
    }
    if (!scp_solve_instance(cta, d, a.p, b, a.io, s, pin)) return false;
    __syncthreads();
    scp_rollout_advance(cta, d, a.p, a.ro, a.io, b, step);
    if (threadIdx.x == 0) {
        if (a.ro.nsteps <= SCP_Q_STEPS) {
            // the instance moves from step to step + 1 (added before it is removed: the counts never undercount), and the
            // lowest live step moves up past every step nobody is at any more
            int *cnt = a.q.hdr + SCP_Q_CNT;
            if (step + 1 < a.ro.nsteps) atomicAdd(cnt + step + 1, 1);
            atomicSub(cnt + step, 1);
            int m = *(volatile int *)(a.q.hdr + 3);
            while (m < a.ro.nsteps - 1 && *(volatile int *)(cnt + m) == 0) ++m;
            atomicMax(a.q.hdr + 3, m);
        }
        stB[7] = (double)(step + 1);
        stB[2] = 0.0;                                          // the next invocation starts a fresh SCP loop ...
        if (!(a.p.qp_warm_start && a.p.qp_warm_carry)) stB[6] = 0.0;   // ... cold, as a new scpb200_scp_solve call would
    }
    return step + 1 >= a.ro.nsteps;
}

// NVEH / HP / NT > 0: literal dimensions and CTA width.  After inlining the compiler folds every index computation
// (n, n1, tile counts, divisions by Hp, strided loops over the CTA) into constants and unrolls the short loops; in
// its generic form the kernel executes ~20 instructions of addressing and loop control per FP64 operation.
// MT = the unit's CTA-width bound: part of the kernel's NAME, so that the instantiations of units compiled for
// different bounds (scp_solve_generic.cu, twice) are different symbols for the linker and the CUDA runtime.
// ROLLOUT: the instantiation behind scpb200_mpc_rollout (set-up and loop closure inside the kernel).  A separate
// instantiation, not a run-time branch: compiled into the per-step kernel the extra code cost every phase of the solver
// 1.5x (measured), although it only runs between MPC steps.
template <bool ALL_SHARED, int NVEH, int HP, int NT, int MT = SCP_MAX_THREADS, bool ROLLOUT = false>
__global__ void __launch_bounds__(NT > 0 ? NT : SCP_MAX_THREADS, (NT > 0 && NT <= 128) ? 3 : (HP > 10 ? 1 : SCP_MIN_CTAS))
k_scp_solve(const __grid_constant__ ScpKernelArgs a)
{
    extern __shared__ double sh[];
    __shared__ int slot;
    __shared__ int setup_flag;
    __shared__ int pin_flag;
    int carry = -1;                                // thread 0: the instance this CTA keeps instead of re-queueing it
    scpb200_dims d = a.d;
    if (NVEH > 0) { d.nVeh = NVEH; d.Hp = HP; d.nObst = 0; }
    const int alpha_slots = NVEH > 0 ? SCP_FIXED_ALPHA : a.alpha_slots, want_H = (NVEH > 0 && HP <= 10) ? 0 : a.want_H;
    Cta cta = {NT > 0 ? NT : (int)blockDim.x};
    ScpBump bp = scp_bump(sh, a.sh_lim, ALL_SHARED ? (double *)0 : a.gws + (size_t)blockIdx.x * a.gl_stride, ALL_SHARED);
    ScpMem s;
    // steering-rate rows: run-time-dimension instantiations only (a literal 0 folds every rate-row loop out of the others)
    const int rate_rows = NVEH > 0 ? 0 : (a.p.enable_rate_rows != 0);
    scp_carve(bp, s, d.nVeh, d.Hp, d.nObst, alpha_slots, want_H, SCP_RED_DOUBLES, rate_rows);
    for (;;) {
        if (threadIdx.x == 0) {
            const int b = carry >= 0 ? carry : queue_pop(a.q);
            carry = -1;
            __threadfence();                       // acquire: the parked state written by the CTA that pushed b
            slot = b;
            if (ROLLOUT) {
                // Laggard rule.  Across MPC steps the batch ends on the instance with the most QPs IN TOTAL, and a FIFO ring
                // advances every live instance by one QP per round: the heavy instances fall behind in steps and run
                // alone at the end (measured: the rollout was slower than the per-step calls it replaces).  An instance
                // that is at the lowest live step is therefore never parked and never re-queued: its chain of QPs runs
                // without waiting, everything else fills the other CTAs round-robin.  In the last step every instance is
                // a laggard and the QP-granular round-robin of the per-step kernel applies again.  Scheduling only:
                // results do not depend on it (park / resume is bit-identical).
                int pin = 0;
                if (b >= 0 && a.ro.nsteps <= SCP_Q_STEPS) {
                    const int step = (int)SCP_LD_COHERENT(a.io.state + (size_t)b * SCP_STATE_W + 7);
                    pin = step <= *(volatile int *)(a.q.hdr + 3) && step < a.ro.nsteps - 1;
                }
                pin_flag = pin;
            }
        }
        __syncthreads();
        const int b = slot;
        if (b < 0) break;
        bool done;
        if (ROLLOUT) {
            // scratch of the warp-level set-up: the normal matrix when it is shared-resident, else the row vectors
            const int ntile = s.ipm.T * (s.ipm.T + 1) / 2;
            double *k1ws = ALL_SHARED ? s.ipm.S : s.ipm.bA;
            const int k1_doubles = ALL_SHARED ? ntile * SCP_TILE2 : 8 * s.ipm.mc;
            done = rollout_invocation(cta, d, a, b, s, k1ws, scp_imax(1, k1_doubles / SCP_K1_WARP_DOUBLES), &setup_flag, pin_flag);
        } else {
            (void)setup_flag;
            done = scp_solve_instance(cta, d, a.p, b, a.io, s);
        }
        __threadfence();                           // release: every thread's writes of this invocation ...
        __syncthreads();                           // ... are ordered before thread 0 hands the instance on
        if (threadIdx.x == 0) {
            if (done) atomicSub(a.q.hdr + 2, 1);
            else if (ROLLOUT && a.ro.nsteps <= SCP_Q_STEPS &&
                     (int)a.io.state[(size_t)b * SCP_STATE_W + 7] <= *(volatile int *)(a.q.hdr + 3) &&
                     (int)a.io.state[(size_t)b * SCP_STATE_W + 7] < a.ro.nsteps - 1) carry = b;   // still furthest behind: keep it
            else queue_push(a.q, b);
        }
    }
}

#ifdef SCP_PHASE_TIMERS
#define SCP_ENTRY_TIMERS(NAME)                                                                         \
    static int NAME##_read_timers(unsigned long long *out32)                                           \
    {                                                                                                  \
        if (cudaDeviceSynchronize() != cudaSuccess) return -1;                                         \
        if (cudaMemcpyFromSymbol(out32, g_scp_prof, sizeof(unsigned long long) * 32) != cudaSuccess) return -1; \
        unsigned long long z[32] = {0};                                                                \
        return cudaMemcpyToSymbol(g_scp_prof, z, sizeof z) == cudaSuccess ? 0 : -1;                    \
    }
#define SCP_ENTRY_TIMERS_PTR(NAME) NAME##_read_timers
#else
#define SCP_ENTRY_TIMERS(NAME)
#define SCP_ENTRY_TIMERS_PTR(NAME) 0
#endif

// Defines `extern "C" const ScpKernelEntry *NAME(void)` for one instantiation of the kernel.
#define SCP_DEFINE_KERNEL_ENTRY(NAME, ...)                                                             \
    static int NAME##_prepare(int threads, size_t smem_bytes, int *ctas_per_sm)                        \
    {                                                                                                  \
        if (cudaFuncSetAttribute(k_scp_solve<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)smem_bytes) != cudaSuccess) return -1;                           \
        if (cudaFuncSetAttribute(k_scp_solve<__VA_ARGS__>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                                 cudaSharedmemCarveoutMaxShared) != cudaSuccess) return -1;            \
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_scp_solve<__VA_ARGS__>, threads, \
                                                             smem_bytes) == cudaSuccess ? 0 : -1;      \
    }                                                                                                  \
    static int NAME##_launch(int grid, int threads, size_t smem_bytes, void *stream, const ScpKernelArgs *a) \
    {                                                                                                  \
        k_scp_solve<__VA_ARGS__><<<grid, threads, smem_bytes, (cudaStream_t)stream>>>(*a);             \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                             \
    }                                                                                                  \
    SCP_ENTRY_TIMERS(NAME)                                                                             \
    extern "C" const ScpKernelEntry *NAME(void)                                                        \
    {                                                                                                  \
        static const ScpKernelEntry e = {NAME##_prepare, NAME##_launch, SCP_ENTRY_TIMERS_PTR(NAME), SCP_MAX_THREADS};    \
        return &e;                                                                                     \
    }
#endif   // SCP_DEVICE_BUILD

// entries of the library (defined in scp_solve_generic.cu / scp_solve_fixed.cu)
extern "C" const ScpKernelEntry *scp_entry_generic_shared(void);
extern "C" const ScpKernelEntry *scp_entry_generic_global(void);
extern "C" const ScpKernelEntry *scp_entry_v8h10_t256(void);       // BASELINE.json configs[1]: 8 vehicles, Hp = 10
extern "C" const ScpKernelEntry *scp_entry_v8h10_t128(void);
extern "C" const ScpKernelEntry *scp_entry_v8h20_t256(void);       // BASELINE.json configs[2]: 8 vehicles, Hp = 20
extern "C" const ScpKernelEntry *scp_entry_v8h20_t512(void);
extern "C" const ScpKernelEntry *scp_entry_generic_shared_wide(void);   // run-time dimensions, CTAs of up to 512 threads
extern "C" const ScpKernelEntry *scp_entry_generic_global_wide(void);
// the rollout instantiations (scpb200_mpc_rollout): run-time dimensions at both CTA-width bounds, and the headline shape
extern "C" const ScpKernelEntry *scp_entry_ro_generic_shared(void);
extern "C" const ScpKernelEntry *scp_entry_ro_generic_global(void);
extern "C" const ScpKernelEntry *scp_entry_ro_generic_shared_wide(void);
extern "C" const ScpKernelEntry *scp_entry_ro_generic_global_wide(void);
extern "C" const ScpKernelEntry *scp_entry_ro_v8h10_t128(void);
