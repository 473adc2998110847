"""IterClass / MPCclass with the reference's names, constructor signatures and attributes (MPC_Iter.py:13-149),
computed on the B200 through libscpb200.so.

    from scp_b200.MPC_Iter import MPCclass, IterClass        # instead of `from MPC_Iter import ...` (main.py:15)

This is the batch-of-one view of the path: every number below comes out of the same CUDA kernels the batched
controller uses (scpb200_ode_predict, scpb200_mpc_setup); NumPy is only the container the caller expects back.
Without a CUDA device the constructors raise (no CPU fallback).
"""
from __future__ import annotations

import numpy as np

from . import _capi

_ENGINES = {}


def scenario_arrays(scenario):
    """The per-vehicle constants of a reference `Scenario` object as C-ABI arrays (veh[1,nVeh,5], poly[1,nVeh,nPts,2])."""
    nVeh = scenario.nVeh
    veh = np.stack([np.asarray(scenario.Lf, float), np.asarray(scenario.Lr, float), np.asarray(scenario.Q, float),
                    np.asarray(scenario.Q_final, float), np.asarray(scenario.R, float)], axis=1).reshape(1, nVeh, 5)
    polys = [np.asarray(r, float) for r in scenario.referenceTrajectories]
    npts = {p.shape[0] for p in polys}
    if len(npts) != 1:
        raise NotImplementedError("reference polylines with different numbers of points per vehicle")
    poly = np.stack(polys)[None]
    return veh, poly


def engine_for(scenario):
    """One BatchSCP(B=1) per problem shape; parameters are refreshed from the scenario on every use."""
    import ctypes as C
    from .batch import BatchSCP
    nObst = int(getattr(scenario, "nObst", 0) or 0)
    veh, poly = scenario_arrays(scenario)
    key = (scenario.nVeh, scenario.Hp, nObst, poly.shape[2])
    eng = _ENGINES.get(key)
    if eng is None:
        eng = BatchSCP(1, scenario.nVeh, scenario.Hp, nObst=nObst, nPts=poly.shape[2])
        _ENGINES[key] = eng
    p = eng.params
    eng.lib.scpb200_default_params(C.byref(p))
    p.dt = float(scenario.dt)
    p.dsafeExtra = float(scenario.dsafeExtra)
    # SCP_controller.py:34 reads scenario.uLim, which the reference never defines: fall back to the only
    # steering bound the scenario has
    p.uLim = float(getattr(scenario, "uLim", scenario.mechanicalSteeringLimit))
    if getattr(scenario.model, "is_noise", False):
        p.noise_sigma = 0.000003                         # Model.py:84-86
        p.seed = int(getattr(scenario, "noise_seed", 0))
    for name in ("trust_radius", "qp_abstol", "qp_reltol", "qp_feastol", "max_scp_iter"):
        if hasattr(scenario, name):
            setattr(p, name, getattr(scenario, name))
    eng.load_inputs(veh=veh, poly=poly, dsafe=np.asarray(scenario.dsafeVehicles, float)[None])
    if nObst:
        eng.load_inputs(dsafe_obst=np.asarray(scenario.dsafeObstacles, float)[None])
    return eng


class IterClass:
    """MPC_Iter.py:13-54: delay-compensated initial state, sampled reference points, obstacle extrapolation."""

    def __init__(self, scenario, x_measured, u_path, obstacleState, uMax):
        import torch
        nVeh, nx, nu, Hp = scenario.nVeh, scenario.model.nx, scenario.model.nu, scenario.Hp
        steps = 10
        T = scenario.delay_x + scenario.dt + scenario.delay_u
        assert (u_path.shape[1] * scenario.tick_length - T < 1e-10)                       # MPC_Iter.py:23
        eng = engine_for(scenario)
        eng.params.noise_counter = int(getattr(scenario, "_iter_counter", 0))
        scenario._iter_counter = eng.params.noise_counter + 1
        self._noise_counter = eng.params.noise_counter
        xm = torch.as_tensor(np.ascontiguousarray(x_measured, dtype=float).reshape(1, nVeh, nx), device=eng.device)
        ur = torch.as_tensor(np.ascontiguousarray(u_path[:, -1], dtype=float).reshape(1, nVeh), device=eng.device)
        Y = eng.ode_predict(xm, ur, float(T), steps=steps, nsub=16)                        # [1,nVeh,steps,6]
        eng.x0.copy_(Y[:, :, -1, :])
        eng.u0.copy_(ur)
        if scenario.nObst:
            # MPC_Iter.py:45-51 (x = x0 + v t along the heading), stored step-major for the kernels
            k = np.arange(Hp)
            step = ((k[None, :] + 1) * scenario.dt + T) * np.asarray(scenario.obstacles)[:, 3].reshape(-1, 1)
            head = np.asarray(scenario.obstacles)[:, 2].reshape(-1, 1)
            ob = np.stack([step * np.cos(head) + np.asarray(obstacleState)[:, 0].reshape(-1, 1),
                           step * np.sin(head) + np.asarray(obstacleState)[:, 1].reshape(-1, 1)], axis=-1)   # [nObst,Hp,2]
            eng.load_inputs(obst=ob[None])
            self.obstacleFutureTrajectories = np.transpose(ob, (0, 2, 1)).copy()           # [nObst,2,Hp]
        eng.setup()
        Yh = Y[0].cpu().numpy()
        self.MPC_delay_compensation_trajectory = np.transpose(Yh, (1, 2, 0)).copy()        # [steps,nx,nVeh]
        self.x0 = Yh[:, -1, :].copy()
        self.u0 = np.asarray(u_path[:, -1], dtype=float).reshape(nVeh, nu).copy()
        self.ReferenceTrajectoryPoints = np.transpose(eng.ref[0].cpu().numpy(), (1, 2, 0)).copy()   # [Hp,2,nVeh]
        self.uMax = uMax
        self.reset = 0
        self._engine = eng
        self._setup_id = _snapshot_id(eng)


def _snapshot_id(eng):
    eng._setup_serial = getattr(eng, "_setup_serial", 0) + 1
    return eng._setup_serial


class MPCclass:
    """MPC_Iter.py:57-149: discretisation, condensed prediction matrices and cost matrices for all vehicles."""

    def __init__(self, scenario, Iter):
        nx, nu, ny = scenario.model.nx, scenario.model.nu, scenario.model.ny
        nVeh, Hp = scenario.nVeh, scenario.Hp
        eng = getattr(Iter, "_engine", None)
        if eng is None or getattr(Iter, "_setup_id", -1) != getattr(eng, "_setup_serial", -2):
            # an Iter that did not come from this module (or a stale one): run K1 on its x0 / u0
            eng = engine_for(scenario)
            eng.load_inputs(x0=np.asarray(Iter.x0, float)[None], u0=np.asarray(Iter.u0, float).reshape(1, nVeh))
            if scenario.nObst:
                eng.load_inputs(obst=np.transpose(np.asarray(Iter.obstacleFutureTrajectories, float), (0, 2, 1))[None])
            eng.setup()
            Iter._engine, Iter._setup_id = eng, _snapshot_id(eng)
        self._engine = eng
        g = eng.g[0].cpu().numpy()            # [nVeh,Hp,2]
        cterm = eng.cterm[0].cpu().numpy()
        abe = eng.abe[0].cpu().numpy()
        self._g, self._cterm, self._abe = g, cterm, abe
        self.Phi_0 = np.transpose(eng.H[0].cpu().numpy(), (1, 2, 0)).copy()                # [Hp,Hp,nVeh]
        self.Psi_0 = np.transpose(eng.qv[0].cpu().numpy(), (1, 0))[:, None, :].copy()      # [Hp,1,nVeh]
        gam = _per_vehicle_gamma(eng)
        self.gamma_0 = gam.reshape(1, nVeh)
        self.const_term = np.transpose(cterm.reshape(nVeh, 2 * Hp), (1, 0))[:, None, :].copy()   # [2Hp,1,nVeh]
        self.Reference = np.transpose(eng.ref[0].cpu().numpy().reshape(nVeh, 2 * Hp), (1, 0)).copy()
        # Toeplitz fill of Mathcal_B from its generator (MPC_Iter.py:146-147)
        MB = np.zeros([ny * Hp, nu * Hp, nVeh])
        for i in range(Hp):
            for j in range(i + 1):
                MB[ny * i:ny * (i + 1), j, :] = g[:, i - j, :].T
        self.Mathcal_B = MB
        self.A = np.repeat(np.transpose(abe[:, :36].reshape(nVeh, nx, nx), (1, 2, 0))[:, :, None, :], Hp, axis=2)
        self.B = np.repeat(np.transpose(abe[:, 36:42], (1, 0))[:, None, None, :], Hp, axis=2)
        self.E = np.repeat(np.transpose(abe[:, 42:48], (1, 0))[:, None, :], Hp, axis=1)

    # Mathcal_A / Mathcal_C are not consumed by the SCP path (only const_term = Mathcal_A x0 + Mathcal_C is,
    # MPC_Iter.py:90, and that comes from the kernel); they are rebuilt on demand for API compatibility.
    @property
    def Mathcal_A(self):
        nVeh, Hp = self._g.shape[0], self._g.shape[1]
        out = np.zeros([2 * Hp, 6, nVeh])
        for v in range(nVeh):
            Ad = self._abe[v, :36].reshape(6, 6)
            CA = np.eye(2, 6)
            for i in range(Hp):
                CA = CA @ Ad
                out[2 * i:2 * i + 2, :, v] = CA
        return out

    @property
    def Mathcal_C(self):
        nVeh, Hp = self._g.shape[0], self._g.shape[1]
        out = np.zeros([2 * Hp, 1, nVeh])
        for v in range(nVeh):
            Ad, Ed = self._abe[v, :36].reshape(6, 6), self._abe[v, 42:48]
            CA, S = np.eye(2, 6), np.zeros((2, 6))
            for i in range(Hp):
                S = S + CA
                out[2 * i:2 * i + 2, 0, v] = S @ Ed
                CA = CA @ Ad
        return out


def _per_vehicle_gamma(eng):
    """gamma_0[v] = Err_v' Q Err_v (MPC_Iter.py:126); the kernel returns their sum, the per-vehicle split is
    recovered from ref / cterm for the attribute only."""
    ref, c, veh = eng.ref[0].cpu().numpy(), eng.cterm[0].cpu().numpy(), eng.veh[0].cpu().numpy()
    Hp = ref.shape[1]
    w = np.repeat(veh[:, 2:3], Hp, axis=1)
    w[:, -1] = veh[:, 3]
    return (w * ((ref - c) ** 2).sum(-1)).sum(-1)
