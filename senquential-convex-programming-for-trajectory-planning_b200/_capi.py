"""ctypes binding of libscpb200.so (include/scpb200.h).  No torch types cross this boundary: callers pass raw
device pointers (torch.Tensor.data_ptr()) and a raw cudaStream_t.

There is no CPU fallback: if the shared library is missing the import of the compute layer raises, and every
compute entry point returns SCPB200_ERR_CUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCPB200_LIB") or os.path.join(HERE, "libscpb200.so")   # override: kernel-tuning builds only

LOG_W = 10
ST_QP_MAXITER, ST_QP_PIVOT, ST_SCP_MAXITER, ST_INFEASIBLE, ST_SETUP, ST_QP_DRES_FLOOR = 1, 2, 4, 8, 16, 32
ST_QP_WARM_RESTART = 64          # log rows only


class Dims(C.Structure):
    _fields_ = [("B", C.c_int32), ("nVeh", C.c_int32), ("Hp", C.c_int32), ("nObst", C.c_int32), ("nPts", C.c_int32)]


class Params(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("uLim", C.c_double), ("dsafeExtra", C.c_double), ("delta_tol", C.c_double),
        ("omega_weight", C.c_double), ("omega_ub", C.c_double), ("constraint_tol", C.c_double),
        ("max_scp_iter", C.c_int32), ("obstacle_eval_mode", C.c_int32),
        ("qp_abstol", C.c_double), ("qp_reltol", C.c_double), ("qp_feastol", C.c_double), ("qp_dual_reg", C.c_double),
        ("inf_bound", C.c_double), ("ipm_max_iter", C.c_int32), ("qp_warm_start", C.c_int32),
        ("trust_radius", C.c_double), ("noise_sigma", C.c_double), ("seed", C.c_uint64),
        ("instance0", C.c_uint32), ("noise_counter", C.c_uint32),
        ("qp_warm_relgap", C.c_double), ("qp_warm_max_iter", C.c_int32), ("qp_warm_min_iter", C.c_int32),
        ("qp_warm_carry", C.c_int32), ("qp_dres_floor_factor", C.c_int32),
        ("enable_rate_rows", C.c_int32), ("log_capacity", C.c_int32), ("duLim", C.c_double),
        ("noise_stream", C.c_uint32), ("reserved0", C.c_uint32),
    ]


class Rollout(C.Structure):
    """scpb200_rollout (include/scpb200.h)"""
    _fields_ = ([("nsteps", C.c_int32), ("mode", C.c_int32)] +
                [(n, C.c_void_p) for n in ("veh", "poly", "dsafe", "dsafe_obst", "obst", "x0", "u0", "x_meas", "u_act", "ref", "g",
                                           "cterm", "H", "qv", "gamma0", "abe", "setup_status", "u", "traj", "U", "obj",
                                           "max_violation", "scp_iters", "ipm_iters", "status")] +
                [(n, C.c_double) for n in ("uMax", "duLim", "mech_limit", "lat_acc_limit", "delay")] +
                [("nsub_delay", C.c_int32), ("nsub_plant", C.c_int32)] +
                [(n, C.c_void_p) for n in ("qp_total", "ipm_total", "status_or", "scp_iters_hist", "status_hist", "U_hist", "x_hist")])


def default_params_py() -> Params:
    """The values scpb200_default_params() writes (kept in sync by tests/test_capi.py)."""
    import math
    p = Params()
    p.dt, p.uLim, p.dsafeExtra, p.delta_tol = 0.4, math.pi / 180.0 * 3.0, 1.0, 1e-3
    p.omega_weight, p.omega_ub, p.constraint_tol = 1e5, 1e25, 2 * 2.1 * 1e-3
    p.max_scp_iter, p.obstacle_eval_mode = 20, 0
    p.qp_abstol, p.qp_reltol, p.qp_feastol, p.qp_dual_reg, p.inf_bound = 1e-7, 1e-13, 1e-9, 1e-11, 1e20
    p.ipm_max_iter, p.trust_radius, p.noise_sigma, p.seed, p.instance0, p.noise_counter = 60, 1e308, 0.0, 0, 0, 0
    p.qp_warm_start, p.qp_warm_relgap, p.qp_warm_max_iter, p.qp_warm_min_iter, p.qp_warm_carry = 1, 1.0, 30, 5, 0
    p.qp_dres_floor_factor = 100
    p.enable_rate_rows, p.log_capacity, p.duLim = 0, 0, math.pi / 180.0 * 6.0
    return p


_P = C.c_void_p

#: name -> argtypes (after the implicit int return); mirrors include/scpb200.h declaration by declaration
PROTOTYPES = {
    "scpb200_version": [],
    "scpb200_last_error": [],
    "scpb200_default_params": [C.POINTER(Params)],
    "scpb200_device_count": [],
    "scpb200_workspace_bytes": [C.POINTER(Dims), C.POINTER(C.c_size_t)],
    "scpb200_qp_workspace_bytes": [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)],
    "scpb200_scp_plan": [C.POINTER(Dims), C.POINTER(C.c_int64)],
    "scpb200_mpc_setup": [C.POINTER(Dims), C.POINTER(Params)] + [_P] * 13,
    "scpb200_assemble_dense": [C.POINTER(Dims), C.POINTER(Params)] + [_P] * 15,
    "scpb200_qcqp_evaluate": [C.POINTER(Dims), C.POINTER(Params)] + [_P] * 16,
    "scpb200_forward_u": [C.POINTER(Dims)] + [_P] * 6,
    "scpb200_ode_predict": [C.POINTER(Dims), C.POINTER(Params), _P, _P, _P, C.c_double, C.c_int32, C.c_int32, _P, _P],
    "scpb200_noise_draws": [C.POINTER(Dims), C.POINTER(Params), C.c_uint32, C.c_uint32, C.c_int32, _P, _P],
    "scpb200_plant_step": [C.POINTER(Dims), C.POINTER(Params), _P, _P, C.c_double, C.c_double, C.c_double, C.c_double,
                           C.c_int32, _P, _P, _P, _P, _P],
    "scpb200_advance_linear": [C.POINTER(Dims), _P, _P, C.c_double, C.c_double, _P, _P, _P],
    "scpb200_qp_solve_dense": [C.POINTER(Dims), C.POINTER(Params), C.c_int32, C.c_int32] + [_P] * 13,
    "scpb200_scp_solve": [C.POINTER(Dims), C.POINTER(Params)] + [_P] * 19,
    "scpb200_scp_solve_ordered": [C.POINTER(Dims), C.POINTER(Params)] + [_P] * 20,
    "scpb200_scp_solve_rate": [C.POINTER(Dims), C.POINTER(Params)] + [_P] * 21,
    "scpb200_work_order": [C.c_int32, _P, _P, _P],
    "scpb200_mpc_rollout": [C.POINTER(Dims), C.POINTER(Params), C.POINTER(Rollout), _P, _P],
}

_lib = None


class Scpb200Error(RuntimeError):
    pass


def load():
    """Load libscpb200.so (built in-tree by build.py).  Raises if it is missing: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Scpb200Error(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` "
                               "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, args in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
        lib.scpb200_last_error.restype = C.c_char_p
        lib.scpb200_default_params.restype = None
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().scpb200_last_error().decode(errors="replace")
        raise Scpb200Error(f"{what} failed with code {rc}: {msg}")
