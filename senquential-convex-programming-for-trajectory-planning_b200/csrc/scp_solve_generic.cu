// scp_solve_generic.cu — K4 with run-time dimensions: working set shared-resident, or with its tail in the global
// workspace (long horizons).
// -DSCP_GENERIC_WIDE: the same two kernels for CTAs of up to 512 threads (one CTA per SM, 128 registers per thread):
// long horizons run one instance per SM and are bound by the latency of their phases, which more warps hide.
#ifdef SCP_GENERIC_WIDE
#define SCP_MAX_THREADS 512
#define SCP_MIN_CTAS 1
#endif
#include "scp_solve_kernel.cuh"

#ifdef SCP_GENERIC_WIDE
SCP_DEFINE_KERNEL_ENTRY(scp_entry_generic_shared_wide, true, 0, 0, 0)
SCP_DEFINE_KERNEL_ENTRY(scp_entry_generic_global_wide, false, 0, 0, 0)
#else
SCP_DEFINE_KERNEL_ENTRY(scp_entry_generic_shared, true, 0, 0, 0)
SCP_DEFINE_KERNEL_ENTRY(scp_entry_generic_global, false, 0, 0, 0)
#endif
