"""SCPcontroller with the reference's call surface (SCP_controller.py:18-400), computed on the B200.

    from scp_b200.SCP_controller import SCPcontroller        # instead of `from SCP_controller import ...` (main.py:13)

Used exactly as main.py:130-134 does:

    controller = SCPcontroller(scenario, Iter, prevOutput)            # prevOutput: [] at step 0, else the last controllerOutput
    U, trajectoryPrediction, controllerOutput = controller.SCP_controller(Iter)

The whole SCP loop (linearise -> QP -> evaluate -> merit/stop test) runs inside one kernel launch
(scpb200_scp_solve); this class only moves the batch-of-one in and out.  Differences from the reference that a
caller can observe are listed in INTEGRATION.md (scalar-only optimisation log; scenario.uLim defaulting to
mechanicalSteeringLimit; the nVeh == 1 retry implemented as the reference evidently intended it).
"""
from __future__ import annotations

import time

import numpy as np

from . import _capi
from .MPC_Iter import MPCclass


class SCPcontroller:
    def __init__(self, scenario, Iter, prevOutput):
        self.scenario = scenario
        self.Iter = Iter
        self.prevOutput = prevOutput
        self.nu = scenario.model.nu
        self.ny = scenario.model.ny
        self.Hp = scenario.Hp
        self.nVeh = scenario.nVeh
        self.nObst = scenario.nObst
        self.dsafeExtra = scenario.dsafeExtra
        self.scenario_uLim = getattr(scenario, "uLim", scenario.mechanicalSteeringLimit)     # SCP_controller.py:34, SURVEY F1
        self.mpc = MPCclass(scenario, Iter)            # K1 on device (or the set-up IterClass already ran)
        self._eng = self.mpc._engine
        self._setup_id = Iter._setup_id                # the engine state this controller's problem data belong to
        self.qcqp = self.QCQP_formulate(scenario)
        self.u = np.zeros([self.nVeh * self.Hp, 1])

    def _engine(self):
        """The engine (one set of device buffers per problem shape) is shared between every Iter / controller of that
        shape.  If another IterClass / MPCclass has used it since this controller was built, its buffers hold the other
        problem's data: reload this controller's problem (scenario constants, x0, u0, obstacles) and re-run K1.  The
        reference's objects are self-contained; this keeps that property."""
        eng = self._eng
        if getattr(eng, "_setup_serial", None) != self._setup_id:
            from .MPC_Iter import engine_for, _snapshot_id
            eng = engine_for(self.scenario)
            nVeh = self.nVeh
            eng.load_inputs(x0=np.asarray(self.Iter.x0, float)[None], u0=np.asarray(self.Iter.u0, float).reshape(1, nVeh))
            if self.nObst:
                eng.load_inputs(obst=np.transpose(np.asarray(self.Iter.obstacleFutureTrajectories, float), (0, 2, 1))[None])
            eng.params.noise_counter = int(getattr(self.Iter, "_noise_counter", 0))
            eng.setup()
            self._eng = eng
            self._setup_id = self.Iter._setup_id = _snapshot_id(eng)
            self.Iter._engine = eng
        return eng

    # ---------------------------------------------------------------------------------------------- the controller
    def SCP_controller(self, Iter):
        if self.prevOutput and ("u" in self.prevOutput):
            self.u = np.asarray(self.prevOutput["u"], dtype=float).reshape([self.Hp * self.nVeh, 1], order="F")
        controllerOutput = {"resultInvalid": False}
        optimizerTimer = time.time()
        self.u, feasible, _, controllerOutput["optimization_log"] = self.SCP_optimizer(self.u)
        if self.nVeh == 1 and not feasible:
            # SCP_controller.py:51-66: retry from full-left, then full-right steering (the reference's version of this
            # branch cannot run: it unpacks three values from a four-tuple and builds an n x n start vector)
            for sign in (+1.0, -1.0):
                u_try, feas_try, _, _ = self.SCP_optimizer(sign * np.ones_like(self.u) * self.scenario_uLim)
                if feas_try:
                    self.u = u_try
                    break
            else:
                print("INFEASIBLE PROBLEM")
                controllerOutput["resultInvalid"] = True
        controllerOutput["u"] = self.u
        trajectoryPrediction, U = self.forward_U(self.u)
        U = np.squeeze(U[:, 0, :])
        controllerOutput["optimizerTime"] = time.time() - optimizerTimer
        return U, trajectoryPrediction, controllerOutput

    def SCP_optimizer(self, u_approx):
        """SCP_controller.py:74-197 in one kernel launch.  Returns (u[n,1], feasible, objValue, optimization_log) with
        the scalar fields of the reference's log (dense P / Aineq per iteration are available on demand through
        BatchSCP.assemble_dense)."""
        import torch
        eng = self._engine()
        if abs(u_approx[0, 0]) < np.spacing(1):
            u_approx[0] = np.spacing(1)                 # :75-76 (mutates the caller's array, as the reference does)
        eng.load_inputs(u=np.ascontiguousarray(u_approx, dtype=float).reshape(1, -1))
        eng.solve()
        torch.cuda.current_stream(eng.device).synchronize()
        its = int(eng.scp_iters[0])
        log = eng.log[0, :its].cpu().numpy()
        u = eng.u[0].cpu().numpy().reshape(-1, 1)
        self._last_traj = eng.traj[0].cpu().numpy()
        self._last_status = int(eng.status[0])
        optimization_log = {
            "slack": [row[0] for row in log], "SCP_ObjVal": [row[1] for row in log], "QCQP_ObjVal": [row[2] for row in log],
            "delta_hat": [row[3] for row in log], "delta": [row[4] for row in log], "feasible": [bool(row[5]) for row in log],
            "max_violation": [row[6] for row in log], "sum_violations": [row[7] for row in log],
            "ipm_iterations": [int(row[8]) for row in log], "qp_status": [int(row[9]) for row in log],
            "u": [u], "status": self._last_status,
        }
        print("iterations: ", its)
        feasible = bool(log[-1, 5]) if its else False
        objValue = float(log[-1, 2]) if its else float("nan")
        return u, feasible, objValue, optimization_log

    def forward_U(self, u):
        """SCP_controller.py:199-213: Traj[Hp,ny,nVeh], U[Hp,nu,nVeh]."""
        import torch
        eng = self._engine()
        ut = torch.as_tensor(np.ascontiguousarray(u, dtype=float).reshape(1, -1), device=eng.device)
        traj, U = eng.forward_u(ut)
        return traj[0].cpu().numpy(), U[0].cpu().numpy()[:, None, :]

    def QCQP_evaluate(self, U):
        """SCP_controller.py:215-265.  Items 3-4 of the reference's tuple (a penalty score and its gradient that no
        caller reads) are returned as None."""
        import torch
        eng = self._engine()
        ut = torch.as_tensor(np.ascontiguousarray(U, dtype=float).reshape(1, -1), device=eng.device)
        ev = eng.evaluate(ut, want_ci=True)
        ci = ev["ci"][0].cpu().numpy()
        cio = ev["ci_obst"][0].cpu().numpy() if ev["ci_obst"] is not None else np.full([self.nVeh, 0, self.Hp], -np.inf)
        return (bool(ev["feasible"][0]), np.array([[float(ev["obj"][0])]]), None, None, float(ev["max_violation"][0]),
                float(ev["sum_violations"][0]), ci, cio)

    def QCQP_formulate(self, scenario):
        """The reference materialises dense Phi/Psi/gamma per (pair, step) here (SCP_controller.py:278-341, 28.7 MB at
        the default size).  The kernels work on the rank structure instead, so only the cost part is exposed."""
        n = self.nVeh * self.Hp
        Phi0 = np.zeros([n, n])
        Psi0 = np.zeros([n, 1])
        for v in range(self.nVeh):
            sl = slice(self.Hp * v, self.Hp * (v + 1))
            Phi0[sl, sl] = self.mpc.Phi_0[:, :, v]
            Psi0[sl, 0] = self.mpc.Psi_0[:, 0, v]
        return {"Phi0": Phi0, "Psi0": Psi0, "gamma0": self.mpc.gamma_0.sum(axis=1)}

    def evaluateInOriginalProblem(self, controlPrediction, trajectoryPrediction, options):
        """SCP_controller.py:343-400 (reporting on the caller's side of the path; the QCQP part goes to the device)."""
        ev = {}
        sq = (self.Iter.ReferenceTrajectoryPoints - trajectoryPrediction) ** 2
        ev["predictionObjectiveValueX"] = sum(self.scenario.Q[v] * sq[0:-1, :, v].sum() + self.scenario.Q_final[v] * sq[-1, :, v].sum()
                                              for v in range(self.nVeh))
        u = controlPrediction[0:self.Hp, :]
        ev["predictionObjectiveValueU"] = sum(self.scenario.R[v] * (u[:, v] ** 2).sum() for v in range(self.nVeh))
        ev["predictionObjectiveValue"] = ev["predictionObjectiveValueX"] + ev["predictionObjectiveValueU"]
        uf = u.reshape(u.shape[0] * u.shape[1], 1, order="F")
        r = self.QCQP_evaluate(uf)
        ev["predictionFeasibleQCQP"], ev["constraintValuesVehicleQCQP"], ev["constraintValuesObstacleQCQP"] = r[0], r[6], r[7]
        tol = 2 * 2.1 * 1e-3                                                   # Config.py:18
        d = trajectoryPrediction[:, :, :, None] - trajectoryPrediction[:, :, None, :]          # [Hp,2,v,v2]
        ci = np.asarray(self.scenario.dsafeVehicles)[None] ** 2 - (d ** 2).sum(axis=1)           # [Hp,v,v2]
        ci = np.transpose(ci, (1, 2, 0))
        iu = np.triu_indices(self.nVeh, 1)
        cv = np.zeros([self.nVeh, self.nVeh, self.Hp])
        cv[iu] = ci[iu]
        cv[(iu[1], iu[0])] = ci[iu]
        ev["constraintValuesVehicle_trajPred"] = cv
        feas = not (ci[iu] > tol).any()
        if self.nObst:
            ob = self.Iter.obstacleFutureTrajectories                          # [nObst,2,Hp]
            do = trajectoryPrediction[:, :, :, None] - np.transpose(ob, (2, 1, 0))[:, :, None, :]   # [Hp,2,v,o]
            co = np.transpose(np.asarray(self.scenario.dsafeObstacles)[None] ** 2 - (do ** 2).sum(axis=1), (1, 2, 0))
            ev["constraintValuesObstacle_trajPred"] = co
            feas = feas and not (co > tol).any()
        ev["predictionFeasible_trajPred"] = feas
        if ev["predictionFeasibleQCQP"] != feas:                               # the reference's flag test is dead code (:391)
            print("feasibility criteria disagree\n")
        ev["predictionFeasible"] = feas
        ev["constraintValuesVehicle"] = cv
        if self.nObst:
            ev["constraintValuesObstacle"] = ev["constraintValuesObstacle_trajPred"]
        return ev
