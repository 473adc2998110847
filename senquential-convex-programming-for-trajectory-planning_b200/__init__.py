"""scp-b200: the SCP-QP hot path of Zhang-Xiaoxue/Senquential-Convex-Programming-for-Trajectory-Planning,
rebuilt for NVIDIA B200 (sm_100a).

Layers
  csrc/            hand-written CUDA (phase-structured CTA kernels) + the C ABI of include/scpb200.h
  _capi.py         ctypes binding of libscpb200.so (raw device pointers, no torch types)
  batch.py         BatchSCP: the batched controller stage (set-up -> fused SCP solve) over torch device buffers
  MPC_Iter.py      IterClass / MPCclass with the reference's names and attributes (MPC_Iter.py:13-149)
  SCP_controller.py SCPcontroller with the reference's call surface (SCP_controller.py:18-400)
  scenarios.py     host-side scenario constants and the synthetic batch generator

The directory name contains '-' and cannot be imported with an `import` statement; use
`importlib.import_module("senquential-convex-programming-for-trajectory-planning_b200")` or the top-level alias module
`scp_b200` at the repository root.
"""
__version__ = "0.1.0"
