// scp_rollout_fixed.cu — the rollout instantiation of K4 for BASELINE.json configs[1] (8 vehicles, Hp = 10, 128-thread
// CTAs): the Monte-Carlo workload the rollout entry exists for.
#include "scp_solve_kernel.cuh"

SCP_DEFINE_KERNEL_ENTRY(scp_entry_ro_v8h10_t128, true, 8, 10, 128, SCP_MAX_THREADS, true)
