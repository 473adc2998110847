"""Generate tests/golden/*.npz by running the reference's OWN Python code (read-only /root/reference).

TEST INFRASTRUCTURE ONLY.  Runs in the build container (the GPU box has no /root/reference); the fixtures
it writes are committed together with this script.

Mechanism: empty stub modules for the reference's unused imports
(oracle/stubs/{autograd,ode,qpsolvers,gurobipy,matplotlib}) and a capture stub named `cvxpy`
(oracle/stubs/cvxpy.py) go ahead of /root/reference on sys.path; `scenario.uLim` (undefined in the
reference, SCP_controller.py:34) is set to mechanicalSteeringLimit; `main.Simulation.runsimulation('SCP')`
then runs unmodified.  The only non-reference arithmetic is the QP solve itself (a8), which goes to the
oracle's coneqp restatement (run in __float128 down to 1e-13, i.e. the converged minimiser of SURVEY F10)
because neither Gurobi nor CVXOPT is installable here.

Usage:  python oracle/make_golden.py [--only hp10|hp20|hp50]
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys
import tempfile
import time
from math import cos, pi, sin

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, HERE)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "stubs"))

import oracle  # noqa: E402
import cvxpy as cp_stub  # noqa: E402  (the capture stub)

QP_ITERS = []
QP_MULT = []           # per QP, in solve order: the multipliers (zA, zub, zlb) of the extended-precision solution


#: the QP "truth": the interior-point iteration run in __float128 down to 1e-13 (see scp_oracle.c, a8)
QUAD_TOL = dict(abstol=1e-13, reltol=1e-13, feastol=1e-13, maxiters=100)


def _solver(P, q, A, b, lb, ub):
    r = oracle.qp_boxed(P, q, A, b, lb, ub, opts=QUAD_TOL, inf_bound=1e20, quad=True)
    assert r["status"] == 0, {k: v for k, v in r.items() if np.ndim(v) == 0}
    QP_ITERS.append(r["iterations"])
    QP_MULT.append((r["zA"].copy(), r["zub"].copy(), r["zlb"].copy()))
    return r["x"]


cp_stub.SOLVER = _solver

import main as ref_main  # noqa: E402
import MPC_Iter as ref_mpc_iter  # noqa: E402
from Model import DefaultVehicle  # noqa: E402
from Scenarios import Scenario  # noqa: E402

ITER_INPUTS = []
_OrigIter = ref_mpc_iter.IterClass


class RecordingIter(_OrigIter):
    def __init__(self, scenario, x_measured, u_path, obstacleState, uMax):
        ITER_INPUTS.append(dict(x_measured=np.array(x_measured, dtype=float), u_path=np.array(u_path, dtype=float),
                                uMax=np.array(uMax, dtype=float)))
        super().__init__(scenario, x_measured, u_path, obstacleState, uMax)


ref_main.IterClass = RecordingIter


def circle_scenario(nveh, radius, Hp):
    """Scenarios.py:109-125 with the radius as a parameter (it is a local constant there; F11)."""
    sc = Scenario(False)
    angles = [2 * pi / nveh * (i + 1) for i in range(nveh)]
    if radius == 30:
        sc.get_circle_scenario(angles)
    else:
        for angle in angles:
            s, c = sin(angle), cos(angle)
            veh = DefaultVehicle()
            veh.x_start, veh.y_start, veh.heading = -c * radius, -s * radius, angle
            veh.referenceTrajectory = np.array([[-c * radius, -s * radius], [c * radius, s * radius]])
            sc.addVehicle(veh)
    sc.Hp = sc.Hu = Hp
    sc.uLim = sc.mechanicalSteeringLimit           # F1
    return sc


def frog_scenario(Hp):
    """Scenarios.py:127-146: one vehicle crossing two lanes of moving obstacles (nVeh = 1, nObst = 22)."""
    sc = Scenario(False)
    sc.get_frog_scenario()
    sc.Hp = sc.Hu = Hp
    sc.uLim = sc.mechanicalSteeringLimit           # F1
    return sc


def run(nveh, radius, Hp, nsim, frog=False):
    ITER_INPUTS.clear()
    QP_ITERS.clear()
    QP_MULT.clear()
    cp_stub.CAPTURE.clear()
    sc = frog_scenario(Hp) if frog else circle_scenario(nveh, radius, Hp)
    sim = ref_main.Simulation(sc, doOnlinePlot=False, isNoise=False)
    if nsim is not None:
        sim.scenario.Nsim = nsim
    ref_main.scenario_choice = "Frog" if frog else "Circle"
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "Data"))
    os.chdir(tmp)
    t0 = time.time()
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            sim.runsimulation("SCP")
    finally:
        os.chdir(cwd)
    print(f"  reference run nVeh={nveh} Hp={Hp} r={radius}: {sim.scenario.Nsim} steps, "
          f"{len(cp_stub.CAPTURE)} QPs, {time.time() - t0:.1f}s")
    return sim


def scenario_constants(sc):
    nVeh = sc.nVeh
    return dict(
        nVeh=nVeh, Hp=sc.Hp, dt=sc.dt, uLim=sc.uLim, dsafeExtra=sc.dsafeExtra, duLim=sc.duLim,
        mechanicalSteeringLimit=sc.mechanicalSteeringLimit, lateralAccelerationLimit=sc.lateralAccelerationLimit,
        tick_length=sc.tick_length, delay_x=sc.delay_x, delay_u=sc.delay_u, ticks_per_sim=sc.ticks_per_sim,
        ticks_delay_u=sc.ticks_delay_u, ticks_delay_x=sc.ticks_delay_x, ticks_total=sc.ticks_total,
        dsafeVehicles=np.array(sc.dsafeVehicles), Lf=np.array(sc.Lf, float), Lr=np.array(sc.Lr, float),
        Q=np.array(sc.Q, float), Q_final=np.array(sc.Q_final, float), R=np.array(sc.R, float),
        poly=np.array([np.asarray(r, float) for r in sc.referenceTrajectories]),
        x_init=np.array([np.asarray(x0).ravel() for x0 in sc.x0]), u_init=np.array(sc.u0, float),
        nObst=sc.nObst, dsafeObstacles=np.array(sc.dsafeObstacles, float),
    )


def step_record(sim, i, dense_iters):
    """All a1-a12 intermediates of MPC step i, from the reference's own objects."""
    sc = sim.scenario
    It = sim.iterationStructs[i]
    mpc = ref_mpc_iter.MPCclass(sc, It)
    out = sim.controllerOutputs[i]
    log = out["optimization_log"]
    nit = len(log["x"])
    rec = dict(
        x_measured=ITER_INPUTS[i]["x_measured"], u_path=ITER_INPUTS[i]["u_path"], uMax=ITER_INPUTS[i]["uMax"],
        x0=It.x0, u0=It.u0, RefPts=It.ReferenceTrajectoryPoints, delay_traj=It.MPC_delay_compensation_trajectory,
        A=mpc.A[:, :, 0, :], B=mpc.B[:, 0, 0, :], E=mpc.E[:, 0, :],
        Mathcal_A=mpc.Mathcal_A, Mathcal_B=mpc.Mathcal_B, Mathcal_C=mpc.Mathcal_C[:, 0, :],
        const_term=mpc.const_term[:, 0, :], Phi_0=mpc.Phi_0, Psi_0=mpc.Psi_0[:, 0, :], gamma_0=mpc.gamma_0[0],
        u_warm=(sim.controllerOutputs[i - 1]["u"].ravel() if i > 0 else np.zeros(sc.nVeh * sc.Hp)),
        u_final=out["u"].ravel(), U=sim.controlPredictions_raw[i], Traj=sim.trajectoryPredictions[:, :, :, i],
        scp_iters=nit,
        prev_u=np.array([np.ravel(v) for v in log["prev_u"]]), x=np.array([np.ravel(v) for v in log["x"]]),
        slack=np.array([float(np.ravel(v)[0]) for v in log["slack"]]),
        SCP_ObjVal=np.array([float(np.ravel(v)[0]) for v in log["SCP_ObjVal"]]),
        QCQP_ObjVal=np.array([float(np.ravel(v)[0]) for v in log["QCQP_ObjVal"]]),
        delta_hat=np.array([float(np.ravel(v)[0]) for v in log["delta_hat"]]),
        delta=np.array([float(np.ravel(v)[0]) for v in log["delta"]]),
        feasible=np.array([bool(v) for v in log["feasible"]]),
    )
    # multipliers of every QP of this step (solve order = step order): the inputs of the independent KKT certificate in
    # tests/test_qp_certificates.py (stationarity / primal / dual sign / complementarity in NumPy longdouble)
    off = sum(len(o["optimization_log"]["x"]) for o in sim.controllerOutputs[:i])
    assert len(QP_MULT) >= off + nit
    rec["zA"] = np.array([QP_MULT[off + it][0] for it in range(nit)])
    rec["zub"] = np.array([QP_MULT[off + it][1] for it in range(nit)])
    rec["zlb"] = np.array([QP_MULT[off + it][2] for it in range(nit)])
    # the warm start the reference actually linearised about in iteration 0 (after the eps tweak of :75-76)
    for it in dense_iters:
        if it < 0:
            it += nit
        if 0 <= it < nit:
            rec[f"P_{it}"] = log["P"][it]
            rec[f"q_{it}"] = np.ravel(log["q"][it])
            rec[f"Aineq_{it}"] = log["Aineq"][it]
            rec[f"bineq_{it}"] = np.ravel(log["bineq"][it])
            rec[f"lb_{it}"] = np.ravel(log["lb"][it])
            rec[f"ub_{it}"] = np.ravel(log["ub"][it])
    if sc.nObst:
        # obstacle rows (SCP_controller.py:106-114): predicted obstacle positions in the kernels' [nObst, Hp, 2] layout
        rec["obst"] = np.transpose(np.asarray(It.obstacleFutureTrajectories, float), (0, 2, 1)).copy()
    # a9 on the final u, straight from the reference
    ctl = ref_main.SCPcontroller(sc, It, [])
    with contextlib.redirect_stdout(io.StringIO()):
        feas, obj, _, _, mv, sv, civ, _ = ctl.QCQP_evaluate(out["u"].reshape(-1, 1))
    rec.update(eval_feasible=bool(feas), eval_obj=float(np.ravel(obj)[0]), eval_max_violation=float(mv),
               eval_sum_violations=float(sv), eval_ci=civ)
    if sc.nObst:
        rec["eval_cio"] = np.asarray(ctl.QCQP_evaluate(out["u"].reshape(-1, 1))[7], float)
    return rec


def patch_raw_controls(sim_cls):
    """main.py:164-174 clamps U in place before storing it; keep the raw controller output as well."""
    pass


def collect(nveh, radius, Hp, nsim, steps, dense_iters, tag, full_run=False, frog=False):
    sim = run(nveh, radius, Hp, nsim, frog=frog)
    sc = sim.scenario
    # raw (pre-clamp) U per step = forward_U of the stored 'u' (SCP_controller.py:69-70)
    sim.controlPredictions_raw = [o["u"].reshape(sc.nVeh, sc.Hp).T.copy() for o in sim.controllerOutputs]
    const = scenario_constants(sc)
    for i in steps:
        rec = step_record(sim, i, dense_iters)
        rec.update({f"sc_{k}": v for k, v in const.items()})
        path = os.path.join(OUT, f"{tag}_step{i}.npz")
        np.savez_compressed(path, **rec)
        print(f"    wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB, "
              f"{rec['scp_iters']} SCP iterations)")
    if full_run:
        nsteps = sc.Nsim
        n = sc.nVeh * sc.Hp
        its = np.array([len(o["optimization_log"]["x"]) for o in sim.controllerOutputs])
        pos = sim.vehiclePathFullRes[0:2]
        mind = np.inf
        for a in range(sc.nVeh):
            for b in range(a + 1, sc.nVeh):
                d = np.sqrt(((pos[:, a, :] - pos[:, b, :]) ** 2).sum(0))
                mind = min(mind, np.nanmin(d))
        feas_all = []
        for o in sim.controllerOutputs:
            feas_all.append(bool(o["optimization_log"]["feasible"][-1]))
        rec = dict(
            x_measured=np.array([d["x_measured"] for d in ITER_INPUTS]),
            u_path=np.array([d["u_path"] for d in ITER_INPUTS]),
            x0=np.array([It.x0 for It in sim.iterationStructs]),
            u0=np.array([It.u0[:, 0] for It in sim.iterationStructs]),
            RefPts=np.array([It.ReferenceTrajectoryPoints for It in sim.iterationStructs]),
            u_final=np.array([o["u"].ravel() for o in sim.controllerOutputs]).reshape(nsteps, n),
            U_clamped=np.moveaxis(sim.controlPredictions, 2, 0), Traj=np.moveaxis(sim.trajectoryPredictions, 3, 0),
            scp_iters=its, qp_total=int(its.sum()), feasible_last=np.array(feas_all),
            QCQP_ObjVal_last=np.array([float(np.ravel(o["optimization_log"]["QCQP_ObjVal"][-1])[0])
                                       for o in sim.controllerOutputs]),
            evaluations_obj_value=np.array([float(np.ravel(e["predictionObjectiveValue"])[0])
                                            for e in sim.evaluations]),
            vehiclePath_every10=sim.vehiclePathFullRes[:, :, ::10], controlPath_every10=sim.controlPathFullRes[:, ::10],
            vehiclePath_step_ends=sim.vehiclePathFullRes[:, :, ::sc.ticks_per_sim],
            min_distance=float(mind), ipm_iters=np.array(QP_ITERS),
        )
        rec.update({f"sc_{k}": v for k, v in const.items()})
        path = os.path.join(OUT, f"{tag}_run.npz")
        np.savez_compressed(path, **rec)
        print(f"    wrote {os.path.relpath(path, ROOT)} ({os.path.getsize(path) / 1024:.0f} KiB): "
              f"{int(its.sum())} QPs, {sum(feas_all)}/{nsteps} feasible, min distance {mind:.4f} m")
    return sim


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--out", default=None, help="write the fixtures here instead of tests/golden (regeneration checks)")
    args = ap.parse_args()
    global OUT
    if args.out:
        OUT = os.path.abspath(args.out)
    os.makedirs(OUT, exist_ok=True)
    if args.only in (None, "hp10"):
        collect(8, 30, 10, None, [0, 6, 10, 29], [0, -1], "circle8_hp10", full_run=True)
    if args.only in (None, "hp20"):
        collect(8, 45, 20, 8, [5, 7], [0], "circle8_hp20")
    if args.only in (None, "hp50"):
        collect(8, 90, 50, 4, [3], [0], "circle8_hp50")
    if args.only in (None, "frog"):
        collect(1, 0, 10, 12, [5, 9], [0, -1], "frog1_hp10", frog=True)
    if args.only in (None, "small"):
        collect(3, 30, 10, 12, [8], [0, -1], "circle3_hp10")


if __name__ == "__main__":
    main()
