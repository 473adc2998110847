"""Batched closed-loop rollouts recorded in the reference's result format.

`BatchRollout` is `Simulation.runsimulation('SCP')` (main.py:83-231) for B scenarios at once with every stage on the
device: delay compensation (MPC_Iter.py:25-33) -> K1 -> K4 -> clamp (main.py:164-174) -> plant (main.py:176-191).
`result_for_plot(b)` returns, for instance b, exactly the dictionary main.py:213-224 dumps to
`Data/<scenario>_num_<nVeh>_control_SCP.json` — same keys, same array shapes, nested lists — so `draw_video.py:44-56`
(which reshapes every entry with order='F') and the other plotting scripts read a batched run like one of their own.

What is and is not reproduced of the tick-resolution arrays:
  * `controlPathFullRes[v, :]` — the actuator's command per tick: step i's first clamped command occupies ticks
    [(i+1)*ticks_per_sim + 1 + ticks_delay_u, (i+2)*ticks_per_sim + 1 + ticks_delay_u) (main.py:180-182), the initial
    command everything before; identical construction.
  * `vehiclePathFullRes[:, v, tick]` — main.py:184-191 restarts its integrator from the step's first tick for each of the
    41 tick times with the command the actuator holds at that tick.  Here: one RK4 integration per command over the step,
    sampled at the ticks (the first `ticks_delay_u - 1` ticks of a step still see the command of two steps ago, the rest
    the previous step's, as in the reference).  Only delay_x = 0 (every shipped scenario) is supported.
  * `controllerRuntime`, `stepTime` — CUDA-event times of the whole batch's controller stage / MPC step, in seconds.
"""
from __future__ import annotations

import json
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .batch import BatchSCP

#: keys of the reference's result file in its own order (main.py:213-224)
RESULT_KEYS = ("vehiclePathFullRes", "obstaclePathFullRes", "controlPathFullRes", "controlPredictions",
               "trajectoryPredictions", "initial_pos", "ReferenceTrajectory", "MPC_delay_compensation_trajectory",
               "evaluations_obj_value", "controllerRuntime", "stepTime")


class BatchRollout:
    def __init__(self, bs: BatchSCP, x_init, u_init, *, mech_limit: float, lat_acc_limit: float, duLim: float,
                 delay_u: float, tick_length: float, Nsim: int, delay_x: float = 0.0,
                 record: Optional[Sequence[int]] = None, obstacles=None):
        """x_init[B,nVeh,6], u_init[B,nVeh]: Scenario.x0 / u0.  `record`: instances whose full history is kept on the
        host (default: all).  `obstacles[nObst,6]` (x, y, heading, speed, length, width; Scenarios.py:105-106) only
        fills `obstaclePathFullRes`."""
        if delay_x != 0.0:
            raise ValueError("only delay_x = 0 (every shipped scenario) is supported")
        self.bs = bs
        self.B, self.nVeh, self.Hp = bs.B, bs.nVeh, bs.Hp
        self.dt = float(bs.params.dt)
        self.mech, self.lat, self.duLim = float(mech_limit), float(lat_acc_limit), float(duLim)
        self.delay = delay_x + self.dt + delay_u
        self.tick = float(tick_length)
        self.tps = int(round(self.dt / self.tick))
        self.tdu = int(round(delay_u / self.tick))
        self.Nsim = int(Nsim)
        self.ticks_total = self.Nsim * self.tps
        self.record = list(range(self.B)) if record is None else [int(b) for b in record]
        self.obstacles = None if obstacles is None else np.asarray(obstacles, float).reshape(-1, 6)
        dev = bs.device
        self.x = torch.as_tensor(np.asarray(x_init, float)).to(dev).contiguous()
        self.u_act = torch.as_tensor(np.asarray(u_init, float)).to(dev).contiguous()
        self._u_init = self.u_act.clone()
        R, V, Hp, N = len(self.record), self.nVeh, self.Hp, self.Nsim
        self.vehiclePathFullRes = np.full((R, 6, V, self.ticks_total + 1), np.nan)
        self.first_commands = np.zeros((R, N, V))
        self.controlPredictions = np.zeros((R, Hp, V, N))
        self.trajectoryPredictions = np.zeros((R, Hp, 2, V, N))
        self.initial_pos = np.zeros((R, 2, V, N))
        self.ReferenceTrajectory = np.zeros((R, Hp, 2, V, N))
        self.delay_traj = np.zeros((R, 10, 6, V, N))
        self.evaluations_obj_value = np.zeros((R, N))
        self.controllerRuntime = np.zeros((N, 1))
        self.stepTime = np.zeros((N, 1))
        self.scp_iters = np.zeros((R, N), dtype=np.int32)
        self.status = np.zeros((R, N), dtype=np.int32)
        self.steps_done = 0
        self.vehiclePathFullRes[:, :, :, 0] = self._rec(self.x).transpose(0, 2, 1)

    def _rec(self, t: torch.Tensor) -> np.ndarray:
        return t[self.record].detach().cpu().numpy()

    def prediction_objective(self) -> torch.Tensor:
        """`predictionObjectiveValue` of evaluateInOriginalProblem (SCP_controller.py:349-362) for the batch, from the
        clamped control prediction and the trajectory prediction of the step just solved (device, [B])."""
        bs = self.bs
        Q, Qf, Rw = bs.veh[:, :, 2], bs.veh[:, :, 3], bs.veh[:, :, 4]                        # [B,nVeh]
        err = (bs.ref.permute(0, 2, 3, 1) - bs.traj) ** 2                                    # [B,Hp,2,nVeh]
        ex = (err[:, :-1].sum(dim=(1, 2)) * Q + err[:, -1].sum(dim=1) * Qf).sum(dim=1)
        eu = ((self._U_clamped ** 2).sum(dim=1) * Rw).sum(dim=1)
        return ex + eu

    def step(self) -> None:
        """One MPC step for the whole batch (main.py:98-210)."""
        i, bs, tps = self.steps_done, self.bs, self.tps
        if i >= self.Nsim:
            raise RuntimeError("rollout already has Nsim steps")
        bs.params.noise_counter = i
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        pred = bs.ode_predict(self.x, self.u_act, self.delay, steps=10, nsub=16)           # [B,nVeh,10,6]
        bs.x0.copy_(pred[:, :, -1, :])
        bs.u0.copy_(self.u_act)
        ev[1].record()
        bs.controller_step()
        ev[2].record()
        # tick-resolution plant path of this step, sampled before the state is advanced
        u_prev = self.u_act.clone()
        # (each ode_predict caller draws from its own Philox stream: 0 delay compensation, 1 / 2 the tick paths)
        bs.params.noise_stream = 1
        ticks = bs.ode_predict(self.x, u_prev, self.dt, steps=tps + 1, nsub=4)             # [B,nVeh,tps+1,6]
        if self.tdu > 1:
            older = self._u_older if i > 0 else self._u_init
            bs.params.noise_stream = 2
            early = bs.ode_predict(self.x, older, (self.tdu - 1) * self.tick, steps=self.tdu, nsub=4)
            ticks[:, :, 1:self.tdu, :] = early[:, :, 1:, :]
        bs.params.noise_stream = 0
        self._u_older = u_prev
        _, self._U_clamped = bs.plant_step(self.x, self.u_act, self.mech, self.lat, self.duLim, want_clamped=True)
        ticks[:, :, tps, :] = self.x                                                        # the state the next step measures
        obj = self.prediction_objective()
        ev[3].record()
        torch.cuda.synchronize(bs.device)
        self.controllerRuntime[i, 0] = ev[1].elapsed_time(ev[2]) * 1e-3
        self.stepTime[i, 0] = ev[0].elapsed_time(ev[3]) * 1e-3
        sl = slice(i * tps + 1, (i + 1) * tps + 1)
        self.vehiclePathFullRes[:, :, :, sl] = self._rec(ticks)[:, :, 1:, :].transpose(0, 3, 1, 2)
        self.first_commands[:, i, :] = self._rec(self._U_clamped)[:, 0, :]
        self.controlPredictions[:, :, :, i] = self._rec(self._U_clamped)
        self.trajectoryPredictions[:, :, :, :, i] = self._rec(bs.traj)
        self.initial_pos[:, :, :, i] = self._rec(bs.x0)[:, :, :2].transpose(0, 2, 1)
        self.ReferenceTrajectory[:, :, :, :, i] = self._rec(bs.ref).transpose(0, 2, 3, 1)
        self.delay_traj[:, :, :, :, i] = self._rec(pred).transpose(0, 2, 3, 1)
        self.evaluations_obj_value[:, i] = self._rec(obj)
        self.scp_iters[:, i] = self._rec(bs.scp_iters)
        self.status[:, i] = self._rec(bs.status)
        self.steps_done += 1

    def run(self, steps: Optional[int] = None) -> "BatchRollout":
        for _ in range(self.Nsim - self.steps_done if steps is None else steps):
            self.step()
        return self

    # ------------------------------------------------------------------------------------------------ result format
    def control_path(self, r: int) -> np.ndarray:
        """controlPathFullRes[nVeh, ticks_total+1] of recorded instance r (main.py:80, 180-182)."""
        T1 = self.ticks_total + 1
        out = np.full((self.nVeh, T1), np.nan)
        out[:, 0:min(T1, self.tdu + self.tps + 1)] = self._u_init[self.record[r]].cpu().numpy()[:, None]
        for i in range(self.steps_done):
            idx = np.arange((i + 1) * self.tps + 1 + self.tdu, (i + 2) * self.tps + 1 + self.tdu)
            idx[idx >= T1 - 1] = T1 - 1
            out[:, idx] = self.first_commands[r, i][:, None]
        return out

    def obstacle_path(self) -> np.ndarray:
        """obstaclePathFullRes[nObst, 2, ticks_total+1] (main.py:66-74)."""
        T1 = self.ticks_total + 1
        if self.obstacles is None or len(self.obstacles) == 0:
            return np.zeros((0, 2, T1))
        t = np.arange(T1) * self.tick
        ob = self.obstacles
        return np.stack([t[None] * ob[:, 3:4] * np.cos(ob[:, 2:3]) + ob[:, 0:1],
                         t[None] * ob[:, 3:4] * np.sin(ob[:, 2:3]) + ob[:, 1:2]], axis=1)

    def result_arrays(self, r: int) -> Dict[str, np.ndarray]:
        """The arrays of main.py:213-224 for recorded instance r, in the reference's shapes."""
        return {
            "vehiclePathFullRes": self.vehiclePathFullRes[r],                               # (nx, nVeh, ticks_total+1)
            "obstaclePathFullRes": self.obstacle_path(),                                    # (nObst, 2, ticks_total+1)
            "controlPathFullRes": self.control_path(r),                                     # (nVeh, ticks_total+1)
            "controlPredictions": self.controlPredictions[r],                               # (Hp, nVeh, Nsim)
            "trajectoryPredictions": self.trajectoryPredictions[r],                         # (Hp, ny, nVeh, Nsim)
            "initial_pos": self.initial_pos[r],                                             # (2, nVeh, Nsim)
            "ReferenceTrajectory": self.ReferenceTrajectory[r],                             # (Hp, 2, nVeh, Nsim)
            "MPC_delay_compensation_trajectory": self.delay_traj[r],                        # (10, nx, nVeh, Nsim)
            "evaluations_obj_value": self.evaluations_obj_value[r],                         # Nsim
            "controllerRuntime": self.controllerRuntime,                                    # (Nsim, 1)
            "stepTime": self.stepTime,                                                      # (Nsim, 1)
        }

    def result_for_plot(self, r: int) -> dict:
        """`result_for_plot1` of main.py:213-224: nested lists under the reference's keys."""
        a = self.result_arrays(r)
        return {k: a[k].tolist() for k in RESULT_KEYS}

    def dump_json(self, r: int, path: str) -> None:
        """What main.py:227-231 writes for one scenario."""
        with open(path, "w") as f:
            json.dump(self.result_for_plot(r), f)


def load_result(path: str, *, nx: int, nVeh: int, nObst: int, Hp: int, Nsim: int, ticks_total: int) -> Dict[str, np.ndarray]:
    """Read a result file the way draw_video.py:42-56 does (reshape with order='F' to the documented shapes)."""
    with open(path) as f:
        res = json.load(f)
    shapes = {"vehiclePathFullRes": (nx, nVeh, ticks_total + 1), "obstaclePathFullRes": (nObst, 2, ticks_total + 1),
              "controlPathFullRes": (nVeh, ticks_total + 1), "controlPredictions": (Hp, nVeh, Nsim),
              "trajectoryPredictions": (Hp, 2, nVeh, Nsim), "initial_pos": (1, 2, nVeh, Nsim),
              "MPC_delay_compensation_trajectory": (10, nx, nVeh, Nsim), "evaluations_obj_value": (Nsim, 1),
              "controllerRuntime": (Nsim, 1), "stepTime": (Nsim, 1), "ReferenceTrajectory": (Hp, 2, nVeh, Nsim)}
    return {k: np.reshape(np.asarray(res[k], dtype=float), shp, order="F") for k, shp in shapes.items()}
