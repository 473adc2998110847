// scp_rollout_generic.cu — the rollout instantiations of K4 (scpb200_mpc_rollout) with run-time dimensions; compiled
// twice like scp_solve_generic.cu (-DSCP_GENERIC_WIDE: CTAs of up to 512 threads).
#ifdef SCP_GENERIC_WIDE
#define SCP_MAX_THREADS 512
#define SCP_MIN_CTAS 1
#endif
#include "scp_solve_kernel.cuh"

#ifdef SCP_GENERIC_WIDE
SCP_DEFINE_KERNEL_ENTRY(scp_entry_ro_generic_shared_wide, true, 0, 0, 0, SCP_MAX_THREADS, true)
SCP_DEFINE_KERNEL_ENTRY(scp_entry_ro_generic_global_wide, false, 0, 0, 0, SCP_MAX_THREADS, true)
#else
SCP_DEFINE_KERNEL_ENTRY(scp_entry_ro_generic_shared, true, 0, 0, 0, SCP_MAX_THREADS, true)
SCP_DEFINE_KERNEL_ENTRY(scp_entry_ro_generic_global, false, 0, 0, 0, SCP_MAX_THREADS, true)
#endif
