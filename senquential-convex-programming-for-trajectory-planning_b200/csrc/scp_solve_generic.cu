// scp_solve_generic.cu — K4 with run-time dimensions: working set shared-resident, or with its tail in the global
// workspace (long horizons).
#include "scp_solve_kernel.cuh"

SCP_DEFINE_KERNEL_ENTRY(scp_entry_generic_shared, true, 0, 0, 0)
SCP_DEFINE_KERNEL_ENTRY(scp_entry_generic_global, false, 0, 0, 0)
