// scp_solve_fixed.cu — K4 instantiated with literal dimensions and CTA width for the shapes BASELINE.json names:
// 8 vehicles, no obstacles, Hp = 10 (configs[1]; -DSCP_FIXED_HP=10 -DSCP_FIXED_NT=256 / 128) and Hp = 20 (configs[2];
// -DSCP_FIXED_HP=20 -DSCP_FIXED_NT=256).  One compilation per (Hp, CTA width).
#if defined(SCP_FIXED_NT) && SCP_FIXED_NT > 256
#define SCP_MAX_THREADS SCP_FIXED_NT      /* sizes this unit's per-warp reduction scratch (the host plan is told: ScpKernelEntry::max_threads) */
#endif
#include "scp_solve_kernel.cuh"

#if !defined(SCP_FIXED_NT) || !defined(SCP_FIXED_HP)
#error "compile with -DSCP_FIXED_HP=10|20 -DSCP_FIXED_NT=256|128"
#endif
#if SCP_FIXED_HP == 10 && SCP_FIXED_NT == 256
SCP_DEFINE_KERNEL_ENTRY(scp_entry_v8h10_t256, true, 8, 10, 256)
#elif SCP_FIXED_HP == 10 && SCP_FIXED_NT == 128
SCP_DEFINE_KERNEL_ENTRY(scp_entry_v8h10_t128, true, 8, 10, 128)
#elif SCP_FIXED_HP == 20 && SCP_FIXED_NT == 256
SCP_DEFINE_KERNEL_ENTRY(scp_entry_v8h20_t256, true, 8, 20, 256)
#elif SCP_FIXED_HP == 20 && SCP_FIXED_NT == 512
SCP_DEFINE_KERNEL_ENTRY(scp_entry_v8h20_t512, true, 8, 20, 512)
#else
#error "unsupported (SCP_FIXED_HP, SCP_FIXED_NT)"
#endif
