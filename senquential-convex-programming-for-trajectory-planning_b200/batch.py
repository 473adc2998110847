"""BatchSCP — the batched controller stage (MPCclass set-up -> fused SCP solve) on one B200.

PyTorch is used for what it is good at here: owning device buffers, pinned host staging buffers and the CUDA
stream.  All arithmetic happens in libscpb200.so (hand-written sm_100a kernels) behind the C ABI of
include/scpb200.h; tensors cross that boundary as raw `data_ptr()` values.  There is no CPU fallback: constructing
a BatchSCP without a CUDA device, or without the built library, raises.

One BatchSCP instance serves B independent scenarios / noise samples of the same shape (nVeh, Hp, nObst, nPts);
multi-GPU runs create one instance per rank over a contiguous slice of the global batch (see parallel.py).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _capi
from ._capi import Dims, Params, check


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_contiguous(), "tensors crossing the C ABI must be contiguous"
    return C.c_void_p(t.data_ptr())


class BatchSCP:
    """Device-resident state of one batch.

    Inputs (device, float64):   x0[B,nVeh,6] u0[B,nVeh] veh[B,nVeh,5]=(Lf,Lr,Q,Q_final,R) poly[B,nVeh,nPts,2]
                                dsafe[B,nVeh,nVeh] (dsafe_obst[B,nVeh,nObst], obst[B,nObst,Hp,2])
    K1 outputs:                 ref g cterm [B,nVeh,Hp,2]  H[B,nVeh,Hp,Hp]  qv[B,nVeh,Hp]  gamma0[B]  abe[B,nVeh,48]
    K4 in/out:                  u[B,n] (warm start in, solution out)
    K4 outputs:                 traj[B,Hp,2,nVeh] U[B,Hp,nVeh] log[B,max_scp_iter,10] scp_iters ipm_iters status obj max_violation
    """

    def __init__(self, B: int, nVeh: int, Hp: int, nObst: int = 0, nPts: int = 2, params: Optional[Params] = None,
                 device: Optional[torch.device] = None, keep_log: bool = True):
        self.lib = _capi.load()
        if not torch.cuda.is_available() or self.lib.scpb200_device_count() < 1:
            raise _capi.Scpb200Error("BatchSCP needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.B, self.nVeh, self.Hp, self.nObst, self.nPts = int(B), int(nVeh), int(Hp), int(nObst), int(nPts)
        self.n = self.nVeh * self.Hp
        self.n1 = self.n + 1
        self.mc = self.Hp * (self.nVeh * (self.nVeh - 1) // 2 + self.nVeh * self.nObst)
        self.dims = Dims(self.B, self.nVeh, self.Hp, self.nObst, self.nPts)
        if params is None:
            params = Params()
            self.lib.scpb200_default_params(C.byref(params))
        self.params = params
        self.keep_log = keep_log
        f64 = dict(dtype=torch.float64, device=self.device)
        i32 = dict(dtype=torch.int32, device=self.device)
        B_, V, Hp_ = self.B, self.nVeh, self.Hp
        with torch.cuda.device(self.device):
            # The caller-facing inputs and results live in ONE contiguous device arena (self.io), every tensor a view of it:
            # inputs only | in/out (x0, u0, u) | results.  A caller with HOST buffers (HostIO) then moves one contiguous range
            # up and one down per MPC step instead of a copy per tensor.
            spec = [("veh", (B_, V, 5), torch.float64), ("poly", (B_, V, self.nPts, 2), torch.float64),
                    ("dsafe", (B_, V, V), torch.float64)]
            if self.nObst:
                spec += [("dsafe_obst", (B_, V, self.nObst), torch.float64), ("obst", (B_, self.nObst, Hp_, 2), torch.float64)]
            spec += [("x0", (B_, V, 6), torch.float64), ("u0", (B_, V), torch.float64), ("u", (B_, self.n), torch.float64),
                     ("U", (B_, Hp_, V), torch.float64), ("traj", (B_, Hp_, 2, V), torch.float64), ("obj", (B_,), torch.float64),
                     ("max_violation", (B_,), torch.float64), ("scp_iters", (B_,), torch.int32),
                     ("ipm_iters", (B_,), torch.int32), ("status", (B_,), torch.int32)]
            self.io_layout, off = {}, 0
            for name, shape, dt in spec:
                nbytes = int(np.prod(shape)) * (8 if dt == torch.float64 else 4)
                self.io_layout[name] = (off, nbytes, shape, dt)
                off += (nbytes + 15) & ~15
            self.io = torch.zeros(off, dtype=torch.uint8, device=self.device)
            for name, (o, nbytes, shape, dt) in self.io_layout.items():
                setattr(self, name, self.io[o:o + nbytes].view(dt).view(shape))
            if not self.nObst:
                self.dsafe_obst = self.obst = None
            #: byte ranges of the arena a host caller sends per step (inputs + in/out) and reads back (in/out + results)
            self.io_in_range = (0, self.io_layout["u"][0] + self.io_layout["u"][1])
            self.io_out_range = (self.io_layout["x0"][0], off)
            self.ref = torch.zeros(B_, V, Hp_, 2, **f64)
            self.g = torch.zeros(B_, V, Hp_, 2, **f64)
            self.cterm = torch.zeros(B_, V, Hp_, 2, **f64)
            self.H = torch.zeros(B_, V, Hp_, Hp_, **f64)
            self.qv = torch.zeros(B_, V, Hp_, **f64)
            self.gamma0 = torch.zeros(B_, **f64)
            self.abe = torch.zeros(B_, V, 48, **f64)
            self.setup_status = torch.zeros(B_, **i32)
            self.log = None
            self._size_log()
            nbytes = C.c_size_t(0)
            check(self.lib.scpb200_workspace_bytes(C.byref(self.dims), C.byref(nbytes)), "scpb200_workspace_bytes")
            self.ws = torch.zeros(max(int(nbytes.value), 256), dtype=torch.uint8, device=self.device)
            self.order = torch.zeros(B_, **i32)
        self.kernel_launches = 0
        #: pull the work queue in descending order of the previous solve's interior-point iteration counts
        #: (longest-processing-time-first).  Scheduling only: results do not depend on it.
        self.schedule_by_previous_work = True
        self._have_work = False

    # ------------------------------------------------------------------------------------------------ plumbing
    def _size_log(self):
        """The per-iteration log holds `log_capacity` rows per instance; the kernel indexes it with that stride and the
        C entry rejects max_scp_iter > log_capacity.  Raising params.max_scp_iter after construction (the facade does,
        per scenario) therefore reallocates here instead of writing past the rows."""
        if not self.keep_log:
            self.params.log_capacity = 0
            return
        need = int(self.params.max_scp_iter)
        if self.log is None or self.log.shape[1] < need:
            with torch.cuda.device(self.device):
                self.log = torch.zeros(self.B, max(need, 1), _capi.LOG_W, dtype=torch.float64, device=self.device)
        self.params.log_capacity = int(self.log.shape[1])

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, t):
        """A contiguous float64 device tensor with the same values (NumPy arrays loaded from .npz may be F-ordered)."""
        return torch.as_tensor(t, dtype=torch.float64).to(self.device).contiguous()

    def plan(self):
        out = (C.c_int64 * 6)()
        check(self.lib.scpb200_scp_plan(C.byref(self.dims), out), "scpb200_scp_plan")
        return dict(grid=out[0], threads=out[1], smem_bytes=out[2], S_in_shared=bool(out[3]), sms=out[4],
                    pair_block_scratch_slots=out[5])

    def load_inputs(self, x0=None, u0=None, veh=None, poly=None, dsafe=None, dsafe_obst=None, obst=None, u=None):
        """Copy whichever inputs are given (numpy or torch, host or device) into the device buffers."""
        for name, val in (("x0", x0), ("u0", u0), ("veh", veh), ("poly", poly), ("dsafe", dsafe),
                          ("dsafe_obst", dsafe_obst), ("obst", obst), ("u", u)):
            if val is None:
                continue
            dst = getattr(self, name)
            src = torch.as_tensor(val, dtype=torch.float64)
            dst.copy_(src.reshape(dst.shape), non_blocking=True)

    def host_io(self) -> "HostIO":
        """A pinned host mirror of the I/O arena (same layout, same tensor names as NumPy views)."""
        return HostIO(self)

    def upload(self, hio: "HostIO"):
        """Host -> device: the step's inputs (veh, poly, dsafe[, obstacles], x0, u0, warm start u) in one copy."""
        a, b = self.io_in_range
        self.io[a:b].copy_(hio.buf[a:b], non_blocking=True)

    def download(self, hio: "HostIO"):
        """Device -> host: x0, u0, u and every result of the step in one copy (asynchronous: synchronise before reading)."""
        a, b = self.io_out_range
        hio.buf[a:b].copy_(self.io[a:b], non_blocking=True)

    # ------------------------------------------------------------------------------------------------ kernels
    def setup(self):
        """K1: MPCclass.__init__ (MPC_Iter.py:57-97) + reference sampling for the whole batch."""
        with torch.cuda.device(self.device):
            check(self.lib.scpb200_mpc_setup(
                C.byref(self.dims), C.byref(self.params), _ptr(self.x0), _ptr(self.u0), _ptr(self.veh), _ptr(self.poly),
                _ptr(self.ref), _ptr(self.g), _ptr(self.cterm), _ptr(self.H), _ptr(self.qv), _ptr(self.gamma0),
                _ptr(self.abe), _ptr(self.setup_status), self._stream()), "scpb200_mpc_setup")
        self.kernel_launches += 1

    def solve(self):
        """K4: SCPcontroller.SCP_controller (SCP_controller.py:40-197) for the whole batch; u is warm start and result."""
        self._size_log()
        with torch.cuda.device(self.device):
            order = None
            if self.schedule_by_previous_work and self._have_work and self.B > 1:
                check(self.lib.scpb200_work_order(self.B, _ptr(self.ipm_iters), _ptr(self.order), self._stream()),
                      "scpb200_work_order")
                self.kernel_launches += 1
                order = self.order
            # u0 (the command being actuated, the set-up's input) anchors the steering-rate rows when they are enabled
            check(self.lib.scpb200_scp_solve_rate(
                C.byref(self.dims), C.byref(self.params), _ptr(self.g), _ptr(self.cterm), _ptr(self.H), _ptr(self.qv),
                _ptr(self.gamma0), _ptr(self.dsafe), _ptr(self.dsafe_obst), _ptr(self.obst), _ptr(self.u),
                _ptr(self.traj), _ptr(self.U), _ptr(self.log), _ptr(self.scp_iters), _ptr(self.ipm_iters),
                _ptr(self.status), _ptr(self.obj), _ptr(self.max_violation), _ptr(order),
                _ptr(self.u0) if self.params.enable_rate_rows else None, _ptr(self.ws), self._stream()),
                "scpb200_scp_solve_rate")
        self.kernel_launches += 2          # k_queue_init + k_scp_solve
        self._have_work = True

    def controller_step(self):
        """The controller stage of one MPC step: K1 then K4 (what main.py:131-134 does per step)."""
        self.setup()
        self.solve()

    def evaluate(self, u: Optional[torch.Tensor] = None, want_ci: bool = False):
        """QCQP_evaluate (SCP_controller.py:215-265) for the batch at u (default: the current solution)."""
        u = self.u if u is None else self._dev(u)
        f64 = dict(dtype=torch.float64, device=self.device)
        obj, mv, sv = torch.zeros(self.B, **f64), torch.zeros(self.B, **f64), torch.zeros(self.B, **f64)
        feas = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        ci = torch.zeros(self.B, self.nVeh, self.nVeh, self.Hp, **f64) if want_ci else None
        cio = torch.zeros(self.B, self.nVeh, self.nObst, self.Hp, **f64) if (want_ci and self.nObst) else None
        with torch.cuda.device(self.device):
            check(self.lib.scpb200_qcqp_evaluate(
                C.byref(self.dims), C.byref(self.params), _ptr(self.g), _ptr(self.cterm), _ptr(self.H), _ptr(self.qv),
                _ptr(self.gamma0), _ptr(u), _ptr(self.dsafe), _ptr(self.dsafe_obst), _ptr(self.obst), _ptr(obj), _ptr(mv),
                _ptr(sv), _ptr(feas), _ptr(ci), _ptr(cio), self._stream()), "scpb200_qcqp_evaluate")
        self.kernel_launches += 1
        return dict(obj=obj, max_violation=mv, sum_violations=sv, feasible=feas, ci=ci, ci_obst=cio)

    def assemble_dense(self, ubar: Optional[torch.Tensor] = None):
        """K2: the dense QP (P,q,Aineq,bineq,lb,ub) of SCP_controller.py:93-128 about ubar (default: current u)."""
        ubar = self.u if ubar is None else self._dev(ubar)
        f64 = dict(dtype=torch.float64, device=self.device)
        B, n1, mc = self.B, self.n1, self.mc
        out = dict(P=torch.empty(B, n1, n1, **f64), q=torch.empty(B, n1, **f64), A=torch.empty(B, mc, n1, **f64),
                   b=torch.empty(B, mc, **f64), lb=torch.empty(B, n1, **f64), ub=torch.empty(B, n1, **f64))
        self.assemble_dense_into(ubar, out)
        return out

    def assemble_dense_into(self, ubar, out):
        with torch.cuda.device(self.device):
            check(self.lib.scpb200_assemble_dense(
                C.byref(self.dims), C.byref(self.params), _ptr(self.g), _ptr(self.cterm), _ptr(self.H), _ptr(self.qv),
                _ptr(ubar), _ptr(self.dsafe), _ptr(self.dsafe_obst), _ptr(self.obst), _ptr(out["P"]), _ptr(out["q"]),
                _ptr(out["A"]), _ptr(out["b"]), _ptr(out["lb"]), _ptr(out["ub"]), self._stream()), "scpb200_assemble_dense")
        self.kernel_launches += 1

    def forward_u(self, u: Optional[torch.Tensor] = None):
        u = self.u if u is None else self._dev(u)
        f64 = dict(dtype=torch.float64, device=self.device)
        traj, U = torch.empty(self.B, self.Hp, 2, self.nVeh, **f64), torch.empty(self.B, self.Hp, self.nVeh, **f64)
        with torch.cuda.device(self.device):
            check(self.lib.scpb200_forward_u(C.byref(self.dims), _ptr(self.g), _ptr(self.cterm), _ptr(u), _ptr(traj), _ptr(U),
                                             self._stream()), "scpb200_forward_u")
        self.kernel_launches += 1
        return traj, U

    def advance_linear(self, uMax: float, duLim: float):
        """Close the loop on the controller's linear model: x0 <- Ad x0 + Bd U[0] + Ed, u0 <- U[0] (clamped)."""
        with torch.cuda.device(self.device):
            check(self.lib.scpb200_advance_linear(C.byref(self.dims), _ptr(self.abe), _ptr(self.U), C.c_double(uMax),
                                                  C.c_double(duLim), _ptr(self.x0), _ptr(self.u0), self._stream()),
                  "scpb200_advance_linear")
        self.kernel_launches += 1

    def plant_step(self, x_meas: torch.Tensor, u_act: torch.Tensor, mech_limit: float, lat_acc_limit: float, duLim: float,
                   T: Optional[float] = None, nsub: int = 64, want_clamped: bool = False):
        """The caller's half of an MPC step (main.py:104-109, 164-191) on device: clamp U (K4 output) to the dynamic
        steering limit and the rate limit, integrate the plant over one sample time with the command being actuated.
        x_meas[B,nVeh,6] and u_act[B,nVeh] are float64 device tensors updated IN PLACE; returns (uMax, U_clamped|None)."""
        assert x_meas.is_cuda and u_act.is_cuda and x_meas.dtype == torch.float64 and u_act.dtype == torch.float64
        assert x_meas.is_contiguous() and u_act.is_contiguous()
        umax = torch.empty(self.B, self.nVeh, dtype=torch.float64, device=self.device)
        uc = torch.empty_like(self.U) if want_clamped else None
        with torch.cuda.device(self.device):
            check(self.lib.scpb200_plant_step(
                C.byref(self.dims), C.byref(self.params), _ptr(self.veh), _ptr(self.U), C.c_double(mech_limit),
                C.c_double(lat_acc_limit), C.c_double(duLim), C.c_double(self.params.dt if T is None else T),
                C.c_int32(nsub), _ptr(x_meas), _ptr(u_act), _ptr(umax), _ptr(uc), self._stream()), "scpb200_plant_step")
        self.kernel_launches += 1
        return umax, uc

    def mpc_step(self, x_meas: torch.Tensor, u_act: torch.Tensor, mech_limit: float, lat_acc_limit: float, duLim: float,
                 delay: float, nsub_delay: int = 16, nsub_plant: int = 64):
        """One whole MPC step of main.py:98-191 with every stage on the device and no host round trip:
        IterClass delay compensation (MPC_Iter.py:25-33) -> K1 -> K4 -> clamp + plant.  `delay` = delay_x + dt + delay_u.
        x_meas / u_act are updated in place; the warm start is the previous solution left in self.u (SCP_controller.py:42-43)."""
        pred = self.ode_predict(x_meas, u_act, delay, steps=2, nsub=nsub_delay * 9)
        self.x0.copy_(pred[:, :, 1, :])
        self.u0.copy_(u_act)
        self.setup()
        self.solve()
        return self.plant_step(x_meas, u_act, mech_limit, lat_acc_limit, duLim, nsub=nsub_plant)

    def rollout(self, nsteps: int, uMax: float, duLim: float, x_meas: Optional[torch.Tensor] = None,
                u_act: Optional[torch.Tensor] = None, mech_limit: float = 0.0, lat_acc_limit: float = 0.0, delay: float = 0.0,
                nsub_delay: int = 144, nsub_plant: int = 64, history: bool = False):
        """`nsteps` closed-loop MPC steps of every instance in ONE launch (scpb200_mpc_rollout).

        x_meas is None: mode 0, the loop is closed on the controller's linear model exactly as
        setup() -> solve() -> advance_linear(uMax, duLim) per step would (self.x0 / self.u0 advance in place).
        x_meas / u_act given (device tensors, advanced in place): mode 1, every step is mpc_step(): delay compensation,
        set-up, SCP solve, clamp + plant integration.  noise_counter of step s is params.noise_counter + s.
        Returns dict(qp_total, ipm_total, status_or[, scp_iters_hist, status_hist, U_hist, x_hist]) of device tensors."""
        mode = 0 if x_meas is None else 1
        if mode == 1:
            assert u_act is not None and x_meas.is_cuda and u_act.is_cuda and x_meas.is_contiguous() and u_act.is_contiguous()
            assert x_meas.dtype == torch.float64 and u_act.dtype == torch.float64
        i32 = dict(dtype=torch.int32, device=self.device)
        f64 = dict(dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            out = dict(qp_total=torch.zeros(self.B, **i32), ipm_total=torch.zeros(self.B, **i32), status_or=torch.zeros(self.B, **i32))
            if history:
                out.update(scp_iters_hist=torch.zeros(self.B, nsteps, **i32), status_hist=torch.zeros(self.B, nsteps, **i32),
                           U_hist=torch.zeros(self.B, nsteps, self.Hp, self.nVeh, **f64),
                           x_hist=torch.zeros(self.B, nsteps + 1, self.nVeh, 6, **f64))
            r = _capi.Rollout()
            r.nsteps, r.mode = int(nsteps), mode
            for name in ("veh", "poly", "dsafe", "dsafe_obst", "obst", "x0", "u0", "ref", "g", "cterm", "H", "qv", "gamma0", "abe",
                         "setup_status", "u", "traj", "U", "obj", "max_violation", "scp_iters", "ipm_iters", "status"):
                t = getattr(self, name)
                setattr(r, name, None if t is None else t.data_ptr())
            r.x_meas = None if x_meas is None else x_meas.data_ptr()
            r.u_act = None if u_act is None else u_act.data_ptr()
            r.uMax, r.duLim, r.mech_limit, r.lat_acc_limit, r.delay = float(uMax), float(duLim), float(mech_limit), float(lat_acc_limit), float(delay)
            r.nsub_delay, r.nsub_plant = int(nsub_delay), int(nsub_plant)
            for name in ("qp_total", "ipm_total", "status_or", "scp_iters_hist", "status_hist", "U_hist", "x_hist"):
                setattr(r, name, out[name].data_ptr() if name in out else None)
            check(self.lib.scpb200_mpc_rollout(C.byref(self.dims), C.byref(self.params), C.byref(r), _ptr(self.ws), self._stream()),
                  "scpb200_mpc_rollout")
        self.kernel_launches += 2          # k_queue_init + k_scp_solve
        self._have_work = False
        return out

    def ode_predict(self, x: torch.Tensor, u_ref: torch.Tensor, T: float, steps: int = 10, nsub: int = 16):
        """Delay-compensation prediction (MPC_Iter.py:25-33) for the batch; returns [B,nVeh,steps,6]."""
        x, u_ref = self._dev(x), self._dev(u_ref)
        out = torch.empty(self.B, self.nVeh, steps, 6, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.scpb200_ode_predict(C.byref(self.dims), C.byref(self.params), _ptr(x), _ptr(u_ref), _ptr(self.veh),
                                               C.c_double(T), C.c_int32(steps), C.c_int32(nsub), _ptr(out), self._stream()),
                  "scpb200_ode_predict")
        self.kernel_launches += 1
        return out


class HostIO:
    """Pinned host staging buffer with the layout of BatchSCP.io; `h.x0`, `h.U`, ... are NumPy views of it.

    The reference's caller keeps its state in host arrays (main.py:112-117); this is that state for B scenarios, laid
    out so that BatchSCP.upload / download move it with one host<->device copy each."""

    def __init__(self, bs: "BatchSCP"):
        self.buf = torch.zeros(bs.io.numel(), dtype=torch.uint8).pin_memory()
        self.nbytes_in = bs.io_in_range[1] - bs.io_in_range[0]
        self.nbytes_out = bs.io_out_range[1] - bs.io_out_range[0]
        raw = self.buf.numpy()
        for name, (o, nbytes, shape, dt) in bs.io_layout.items():
            setattr(self, name, raw[o:o + nbytes].view(np.float64 if dt == torch.float64 else np.int32).reshape(shape))


def qp_solve_dense(P, q, A, b, lb, ub, params: Optional[Params] = None):
    """The CVXOPT/Gurobi-replacement entry on dense device tensors: P[B,n1,n1] q[B,n1] A[B,mc,n1] b[B,mc] lb,ub[B,n1].

    Returns dict(x[B,n1], fval[B], iters[B], status[B], zA[B,mc]) of device tensors."""
    lib = _capi.load()
    if not torch.cuda.is_available():
        raise _capi.Scpb200Error("qp_solve_dense needs a CUDA device; there is no CPU fallback")
    dev = P.device
    B, n1, mc = P.shape[0], P.shape[1], A.shape[1]
    if params is None:
        params = Params()
        lib.scpb200_default_params(C.byref(params))
    f64 = dict(dtype=torch.float64, device=dev)
    args = [t.to(torch.float64).contiguous() for t in (P, q, A, b, lb, ub)]
    x, fval = torch.empty(B, n1, **f64), torch.empty(B, **f64)
    iters, status = torch.zeros(B, dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)
    zA = torch.empty(B, mc, **f64)
    with torch.cuda.device(dev):
        nbytes = C.c_size_t(0)
        check(lib.scpb200_qp_workspace_bytes(C.c_int32(n1), C.c_int32(mc), C.byref(nbytes)), "scpb200_qp_workspace_bytes")
        ws = torch.zeros(max(int(nbytes.value), 256), dtype=torch.uint8, device=dev)
        d = Dims(B, 1, 1, 0, 2)
        check(lib.scpb200_qp_solve_dense(C.byref(d), C.byref(params), C.c_int32(n1), C.c_int32(mc), *[_ptr(t) for t in args],
                                         _ptr(x), _ptr(fval), _ptr(iters), _ptr(status), _ptr(zA), _ptr(ws),
                                         C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "scpb200_qp_solve_dense")
        torch.cuda.current_stream(dev).synchronize()      # ws must outlive the kernel
    return dict(x=x, fval=fval, iters=iters, status=status, zA=zA)
