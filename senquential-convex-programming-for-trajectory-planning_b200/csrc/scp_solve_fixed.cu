// scp_solve_fixed.cu — K4 instantiated with literal dimensions and CTA width for the shape BASELINE.json's headline
// configuration names (8 vehicles, Hp = 10, no obstacles).  Compiled once per CTA width (-DSCP_FIXED_NT=256 / 128).
#include "scp_solve_kernel.cuh"

#ifndef SCP_FIXED_NT
#error "compile with -DSCP_FIXED_NT=256 or 128"
#endif
#if SCP_FIXED_NT == 256
SCP_DEFINE_KERNEL_ENTRY(scp_entry_v8h10_t256, true, 8, 10, 256)
#elif SCP_FIXED_NT == 128
SCP_DEFINE_KERNEL_ENTRY(scp_entry_v8h10_t128, true, 8, 10, 128)
#else
#error "unsupported SCP_FIXED_NT"
#endif
