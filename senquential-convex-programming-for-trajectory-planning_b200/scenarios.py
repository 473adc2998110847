"""Host-side scenario constants and the synthetic batch generator.

The reference builds its scenarios once, on the host, in Scenarios.py:40-252; that code is set-up, not hot path, and
only its OUTPUTS feed the path.  This module reproduces those outputs for the circle / intersection family
(Scenarios.py:109-125) for a whole batch of perturbed instances, as plain NumPy arrays in the C-ABI layouts.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# Model.py:20-28 (DefaultVehicle), Scenarios.py:45-58
V_SPEED, V_LENGTH, V_WIDTH, V_LF, V_LR, V_Q, V_QF, V_R = 4.0, 0.98, 0.88, 0.34, 0.34, 1.0, 20.0, 4000.0
DT, TICK, DELAY_X, DELAY_U = 0.4, 0.01, 0.0, 0.03
MECH_LIMIT = math.pi / 180 * 3
DU_LIM = 2 * MECH_LIMIT
DSAFE_EXTRA = 1.0
LAT_ACC_LIMIT = 9.81 / 2

#: circle radius per horizon: the stock radius 30 is infeasible at u=0 for Hp >= 20
RADIUS_FOR_HP = {10: 30.0, 20: 45.0, 50: 90.0}


def safety_distance(v1, v2, dt=DT, L1=V_LENGTH, W1=V_WIDTH, L2=V_LENGTH, W2=V_WIDTH):
    """Scenarios.py:236-243: dsafe = sqrt((sum v * dt / 2)^2 + (r1 + r2)^2)."""
    chord = (v1 + v2) * dt
    R = np.sqrt((L1 / 2) ** 2 + (W1 / 2) ** 2) + np.sqrt((L2 / 2) ** 2 + (W2 / 2) ** 2)
    return np.sqrt((chord / 2) ** 2 + R ** 2)


@dataclass
class CircleBatch:
    """A batch of perturbed circle scenarios positioned at MPC step `step[b]` of the open-loop (u = 0) run."""
    B: int
    nVeh: int
    Hp: int
    x0: np.ndarray          # [B,nVeh,6]
    u0: np.ndarray          # [B,nVeh]
    veh: np.ndarray         # [B,nVeh,5] = (Lf, Lr, Q, Q_final, R)
    poly: np.ndarray        # [B,nVeh,2,2]
    dsafe: np.ndarray       # [B,nVeh,nVeh]
    step: np.ndarray        # [B]
    radius: float
    meta: dict = field(default_factory=dict)


def circle_batch(B: int, nVeh: int = 8, Hp: int = 10, instance0: int = 0, seed: int = 20261018, radius: float | None = None,
                 step_lo: int = 6, step_hi: int = 14, perturb: bool = True, dt: float = DT) -> CircleBatch:
    """Instances instance0 .. instance0+B-1 of the synthetic family of SURVEY 8d.

    Vehicle i starts on a circle of radius R_c at angle 2 pi (i+1)/nVeh + eps_theta heading for the antipode at speed
    4 (1 + eps_v); eps_theta ~ U(-0.05, 0.05) rad, eps_v ~ U(-0.05, 0.05), step ~ U{step_lo..step_hi}, all from
    numpy.random.default_rng(seed + instance_id) so that an instance does not depend on how the batch is sharded.
    """
    R_c = float(radius if radius is not None else RADIUS_FOR_HP.get(Hp, 30.0 * max(1.0, Hp / 10.0) * 0.9))
    x0 = np.zeros((B, nVeh, 6))
    u0 = np.zeros((B, nVeh))
    veh = np.tile(np.array([V_LF, V_LR, V_Q, V_QF, V_R]), (B, nVeh, 1))
    poly = np.zeros((B, nVeh, 2, 2))
    dsafe = np.zeros((B, nVeh, nVeh))
    steps = np.zeros(B, dtype=np.int64)
    base = 2 * math.pi * (np.arange(nVeh) + 1) / nVeh
    for b in range(B):
        rng = np.random.default_rng(seed + instance0 + b)
        if perturb:
            eth = rng.uniform(-0.05, 0.05, nVeh)
            ev = rng.uniform(-0.05, 0.05, nVeh)
            s = int(rng.integers(step_lo, step_hi + 1))
        else:
            eth, ev, s = np.zeros(nVeh), np.zeros(nVeh), step_lo
        th = base + eth
        v = V_SPEED * (1 + ev)
        c, sn = np.cos(th), np.sin(th)
        travelled = v * (s * dt)
        x0[b, :, 0] = -c * R_c + c * travelled
        x0[b, :, 1] = -sn * R_c + sn * travelled
        x0[b, :, 2] = th
        x0[b, :, 3] = v
        poly[b, :, 0, 0], poly[b, :, 0, 1] = -c * R_c, -sn * R_c
        poly[b, :, 1, 0], poly[b, :, 1, 1] = c * R_c, sn * R_c
        dsafe[b] = safety_distance(v[:, None], v[None, :], dt)
        steps[b] = s
    return CircleBatch(B=B, nVeh=nVeh, Hp=Hp, x0=x0, u0=u0, veh=veh, poly=poly, dsafe=dsafe, step=steps, radius=R_c,
                       meta=dict(seed=seed, instance0=instance0, step_lo=step_lo, step_hi=step_hi, perturb=perturb))


def reference_circle(nVeh: int = 8, Hp: int = 10, radius: float | None = None):
    """The unperturbed reference scenario (Scenarios.py:109-125 + complete_scenario) at simulation start."""
    return circle_batch(1, nVeh=nVeh, Hp=Hp, radius=radius, step_lo=0, step_hi=0, perturb=False)
