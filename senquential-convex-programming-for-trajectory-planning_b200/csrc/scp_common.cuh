// scp_common.cuh — execution-model glue shared by every kernel of libscpb200.
//
// All CTA-cooperative code in this library is written as a sequence of PHASES: inside a phase each thread
// works on its own slice of shared memory; phases are separated by a CTA barrier.  The same source is
// compiled two ways:
//   * by nvcc for sm_100a — a phase is `{ tid = threadIdx.x; ... } __syncthreads();`
//   * by g++ for the kernel-logic emulator under tests/emu (CPU tests only, never loaded by the product) —
//     a phase is a loop over tid, in ascending or descending order (a cheap intra-phase race detector).
// Warp-level leaf routines (shuffle reductions, the 8x8 tile factorisation) have a device body and a plain
// sequential host body.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/scpb200.h"

#if defined(__CUDACC__)
#define SCP_FN __device__ __forceinline__
#define SCP_NOINLINE_FN static __device__ __noinline__      /* leaf routines with their own register allocation */
#define SCP_MFN __device__ __forceinline__
#define SCP_HDFN __host__ __device__ __forceinline__
#define SCP_HDMFN __host__ __device__ __forceinline__
#define SCP_DEVICE_BUILD 1
#else
#define SCP_FN static inline
#define SCP_NOINLINE_FN static inline
#define SCP_MFN inline
#define SCP_HDFN static inline
#define SCP_HDMFN inline
#define SCP_DEVICE_BUILD 0
#endif

#ifndef SCP_MAX_THREADS
#define SCP_MAX_THREADS 256      // upper bound on threads per CTA of every kernel in this library
#endif
#define SCP_MAX_WARPS (SCP_MAX_THREADS / 32)
#define SCP_TILE 8
#define SCP_TILE2 64

// ------------------------------------------------------------------------------------------------ phases
// Reductions: a reducing phase leaves one partial per warp in `red`; after the phase's barrier every thread folds them.
// Nothing separates those reads from the NEXT reducing phase's writes but that phase's own work, so a warp held up
// right after the barrier could read partials of the following reduction (observed on B200 as rounding-level,
// schedule-dependent differences in ~1 of 1000 instances).  The scratch is therefore double-buffered: consecutive
// reductions alternate between two sets of SCP_RED_SLOTS slots (`Cta::flip`), and a late reader is always at most one
// reduction behind its writers' barrier.
#define SCP_RED_SLOTS 8
#define SCP_RED_DOUBLES (2 * SCP_RED_SLOTS * SCP_MAX_WARPS)
#if SCP_DEVICE_BUILD
struct Cta {
    int nt;
    int flip;      // 0 or SCP_RED_SLOTS: which half of the reduction scratch the current reduction uses
};
#define CTA_PHASE(tid) { const int tid = (int)threadIdx.x;
#define CTA_PHASE_END } __syncthreads();
#define CTA_RED_BEGIN(cta, nslots) (cta).flip ^= SCP_RED_SLOTS;
#define CTA_PHASE_END_RED(cta, red, nslots) } __syncthreads();
#else
#define SCP_EMU_MAXNT 1024
struct Cta {
    int nt;
    int flip;                              // as on the device
    int reverse;                           // emulate threads in descending order
    double part[8 * SCP_EMU_MAXNT];        // per-thread partials of the reduction slots
    int part_kind[8];                      // 0 = sum, 1 = max
};
static inline int scp_emu_tid(const Cta &c, int i) { return c.reverse ? c.nt - 1 - i : i; }
#define CTA_PHASE(tid) for (int tid##_i = 0; tid##_i < cta.nt; ++tid##_i) { const int tid = scp_emu_tid(cta, tid##_i);
#define CTA_PHASE_END }
#define CTA_RED_BEGIN(cta, nslots)                                                     \
    (cta).flip ^= SCP_RED_SLOTS;                                                       \
    for (int s_ = 0; s_ < (nslots); ++s_)                                              \
        for (int t_ = 0; t_ < (cta).nt; ++t_) (cta).part[s_ * SCP_EMU_MAXNT + t_] = 0.0;
// butterfly in the device's order: v[l] (+|max)= v[l ^ off], off = 16..1; lane 0 of each warp publishes
#define CTA_PHASE_END_RED(cta, red, nslots)                                            \
    }                                                                                  \
    for (int s_ = 0; s_ < (nslots); ++s_)                                              \
        for (int w_ = 0; w_ < (cta).nt / 32; ++w_) {                                   \
            double v_[32], n_[32];                                                     \
            for (int l_ = 0; l_ < 32; ++l_) v_[l_] = (cta).part[s_ * SCP_EMU_MAXNT + w_ * 32 + l_]; \
            for (int off_ = 16; off_ >= 1; off_ >>= 1) {                               \
                for (int l_ = 0; l_ < 32; ++l_)                                        \
                    n_[l_] = (cta).part_kind[s_] ? fmax(v_[l_], v_[l_ ^ off_]) : v_[l_] + v_[l_ ^ off_]; \
                for (int l_ = 0; l_ < 32; ++l_) v_[l_] = n_[l_];                       \
            }                                                                          \
            (red)[(s_ + (cta).flip) * SCP_MAX_WARPS + w_] = v_[0];                     \
        }
#endif

// Warp sections: inside WARP_SECTION each warp works on its own data; WARP_PHASE ... WARP_PHASE_END separates
// steps that exchange data between the lanes of one warp through shared memory (__syncwarp).  A CTA_SYNC must
// follow a warp section before other warps consume its results.
#if SCP_DEVICE_BUILD
#define WARP_SECTION(w, nw) { const int w = (int)threadIdx.x >> 5; const int nw = cta.nt >> 5;
#define WARP_SECTION_END }
#define WARP_PHASE(lane) { const int lane = (int)threadIdx.x & 31;
#define WARP_PHASE_END } __syncwarp();
#define CTA_SYNC __syncthreads();
#else
#define WARP_SECTION(w, nw)                                                                       \
    for (int w##_i = 0; w##_i < (cta.nt >> 5); ++w##_i) {                                         \
        const int w = cta.reverse ? (cta.nt >> 5) - 1 - w##_i : w##_i;                            \
        const int nw = cta.nt >> 5;
#define WARP_SECTION_END }
#define WARP_PHASE(lane) for (int lane##_i = 0; lane##_i < 32; ++lane##_i) { const int lane = cta.reverse ? 31 - lane##_i : lane##_i;
#define WARP_PHASE_END }
#define CTA_SYNC
#endif

// Inside a reducing phase every thread calls these exactly once per slot (uniform control flow).
#if SCP_DEVICE_BUILD
SCP_FN double scp_warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
SCP_FN double scp_warp_max(double v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
#define CTA_RED_SUM(cta, red, slot, tid, val)                                 \
    {                                                                         \
        double r_ = scp_warp_sum(val);                                        \
        if (((tid) & 31) == 0) (red)[((slot) + (cta).flip) * SCP_MAX_WARPS + ((tid) >> 5)] = r_; \
    }
#define CTA_RED_MAX(cta, red, slot, tid, val)                                 \
    {                                                                         \
        double r_ = scp_warp_max(val);                                        \
        if (((tid) & 31) == 0) (red)[((slot) + (cta).flip) * SCP_MAX_WARPS + ((tid) >> 5)] = r_; \
    }
#else
#define CTA_RED_SUM(cta, red, slot, tid, val) { (cta).part_kind[slot] = 0; (cta).part[(slot) * SCP_EMU_MAXNT + (tid)] = (val); }
#define CTA_RED_MAX(cta, red, slot, tid, val) { (cta).part_kind[slot] = 1; (cta).part[(slot) * SCP_EMU_MAXNT + (tid)] = (val); }
#endif

// after the reducing phase: every thread folds the per-warp results (uniform value)
SCP_FN double cta_red_sum(const Cta &cta, const double *red, int slot)
{
    double t = 0.0;
    const int nw = cta.nt >> 5;
    for (int w = 0; w < nw; ++w) t += red[(slot + cta.flip) * SCP_MAX_WARPS + w];
    return t;
}
SCP_FN double cta_red_max(const Cta &cta, const double *red, int slot)
{
    double t = red[(slot + cta.flip) * SCP_MAX_WARPS];
    const int nw = cta.nt >> 5;
    for (int w = 1; w < nw; ++w) t = fmax(t, red[(slot + cta.flip) * SCP_MAX_WARPS + w]);
    return t;
}

// ------------------------------------------------------------------------------------------------ phase timers
// Tuning builds (-DSCP_PHASE_TIMERS) accumulate, per code region, the cycles thread 0 of every CTA spends between
// consecutive marks (regions end at CTA barriers, so thread 0's clock is the CTA's).  Off in the product build.
#if SCP_DEVICE_BUILD && defined(SCP_PHASE_TIMERS)
static __device__ unsigned long long g_scp_prof[32];   // one copy per translation unit
#define SCP_TIMER_DECL long long scp_t_last = clock64();
#define SCP_TIMER(id)                                                                         \
    if (threadIdx.x == 0) {                                                                   \
        const long long t_ = clock64();                                                       \
        atomicAdd(&g_scp_prof[id], (unsigned long long)(t_ - scp_t_last));                    \
        scp_t_last = t_;                                                                      \
    }
#define SCP_TIMER_ARG , long long &scp_t_last
#define SCP_TIMER_PASS , scp_t_last
#else
#define SCP_TIMER_DECL
#define SCP_TIMER(id)
#define SCP_TIMER_ARG
#define SCP_TIMER_PASS
#endif

// loads of data another CTA may have written during this launch (parked instances) must bypass L1
#if SCP_DEVICE_BUILD
#define SCP_LD_COHERENT(ptr) __ldcg(ptr)
#else
#define SCP_LD_COHERENT(ptr) (*(ptr))
#endif

// Data that another CTA may have written during this launch and that this SM may hold in L1 from before (the rollout
// entry re-runs the set-up of an instance on whichever CTA pops it): `coh` selects loads that bypass L1.
SCP_FN double scp_ldc(const double *ptr, bool coh)
{
#if SCP_DEVICE_BUILD
    return coh ? __ldcg(ptr) : *ptr;
#else
    (void)coh;
    return *ptr;
#endif
}

// ------------------------------------------------------------------------------------------------ misc
SCP_HDFN int scp_imin(int a, int b) { return a < b ? a : b; }
SCP_HDFN int scp_imax(int a, int b) { return a > b ? a : b; }
SCP_HDFN int scp_round_up(int a, int m) { return (a + m - 1) / m * m; }

// index of lower-triangular tile (I, J), J <= I, in the tile-packed normal matrix
SCP_HDFN int scp_tile_off(int I, int J) { return ((I * (I + 1) >> 1) + J) * SCP_TILE2; }
// Position of element (r, c) inside an 8x8 tile.  Rows are 64 bytes; in rows 2,3,6,7 the two 32-byte halves
// are swapped.  With plain row-major tiles rows r and r+2 share shared-memory banks, so the column-shaped
// accesses of the FP64 mma fragments (8 rows x 32 bytes) would be 4-way bank-conflicted; with the swap rows
// 0..3 cover all 32 banks and a fragment load costs the minimum of two wavefronts.
SCP_HDFN constexpr int scp_tphys(int r, int c) { return (r << 3) + ((((c >> 2) ^ (r >> 1)) & 1) << 2) + (c & 3); }
// element (ci, cj), ci >= cj
SCP_HDFN int scp_sidx(int ci, int cj)
{
    return scp_tile_off(ci >> 3, cj >> 3) + scp_tphys(ci & 7, cj & 7);
}
// t -> (ii, jj) with ii >= jj, t = ii(ii+1)/2 + jj
SCP_HDFN void scp_tri_decode(int t, int *ii, int *jj);
// The same through a table for the tile loops of the factorisation, whose t is warp-uniform (constant-cache broadcast,
// two loads instead of a search).
#define SCP_TRI_TAB_ROWS 52
#define SCP_TRI_TAB (SCP_TRI_TAB_ROWS * (SCP_TRI_TAB_ROWS + 1) / 2)
struct ScpTriTab {
    unsigned char ii[SCP_TRI_TAB], jj[SCP_TRI_TAB];
    constexpr ScpTriTab() : ii(), jj()
    {
        int t = 0;
        for (int i = 0; i < SCP_TRI_TAB_ROWS; ++i)
            for (int j = 0; j <= i; ++j) { ii[t] = (unsigned char)i; jj[t] = (unsigned char)j; ++t; }
    }
};
#if SCP_DEVICE_BUILD
static __constant__ ScpTriTab c_scp_tri = ScpTriTab();
#else
static const ScpTriTab c_scp_tri = ScpTriTab();
#endif
SCP_FN void scp_tri_lookup(int t, int *ii, int *jj)
{
    if (t < SCP_TRI_TAB) { *ii = c_scp_tri.ii[t]; *jj = c_scp_tri.jj[t]; }
    else scp_tri_decode(t, ii, jj);
}
SCP_HDFN void scp_tri_decode(int t, int *ii, int *jj)
{
    int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);      // estimate, corrected below
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    *ii = i;
    *jj = t - i * (i + 1) / 2;
}
