"""Host-side pieces of the result / wire format (main.py:213-231, draw_video.py:42-56).  CPU only:
the recorder's device-side loop is covered by tests/test_gpu_parity.py::test_batch_rollout_in_the_reference_result_format."""
import importlib
import json
import os
import re

import numpy as np
import pytest

PKG = "senquential-convex-programming-for-trajectory-planning_b200"
results = importlib.import_module(PKG + ".results")
REF_MAIN = "/root/reference/main.py"


def _bare_rollout(nVeh=3, Nsim=5, tps=40, tdu=3, steps_done=5, seed=0):
    """A BatchRollout with only the host-side fields its formatting methods read (no device)."""
    import torch
    rng = np.random.default_rng(seed)
    ro = object.__new__(results.BatchRollout)
    ro.nVeh, ro.Nsim, ro.tps, ro.tdu, ro.tick = nVeh, Nsim, tps, tdu, 0.01
    ro.ticks_total, ro.steps_done, ro.record = Nsim * tps, steps_done, [0]
    ro._u_init = torch.as_tensor(rng.normal(size=(1, nVeh)))
    ro.first_commands = rng.normal(size=(1, Nsim, nVeh))
    ro.obstacles = None
    return ro


def test_control_path_follows_the_reference_bookkeeping():
    """controlPathFullRes: main.py:80 (initial command up to tick ticks_delay_u + ticks_per_sim) and main.py:177-182
    (step i's first command shifted ticks_per_sim + ticks_delay_u into the future, truncated at the last tick)."""
    ro = _bare_rollout()
    got = ro.control_path(0)
    nVeh, tps, tdu, T1 = ro.nVeh, ro.tps, ro.tdu, ro.ticks_total + 1
    want = np.full((nVeh, T1), np.nan)
    u0 = ro._u_init[0].numpy()
    for v in range(nVeh):
        want[v, 0:tdu + tps + 1] = u0[v]
    for i in range(ro.steps_done):
        for v in range(nVeh):
            sl = np.array(range(i * tps + 1 + tdu + tps, (i + 1) * tps + 1 + tdu + tps))
            sl[sl >= want.shape[1] - 1] = want.shape[1] - 1
            want[v, sl] = ro.first_commands[0, i, v]
    np.testing.assert_array_equal(got, want)
    assert not np.isnan(got).any()


def test_obstacle_path_is_constant_velocity():
    ro = _bare_rollout()
    assert ro.obstacle_path().shape == (0, 2, ro.ticks_total + 1)
    ro.obstacles = np.array([[7.0, -15.0, np.pi / 2, 2.0, 4.0, 2.0], [14.0, 3.0, 0.0, 1.0, 4.0, 2.0]])
    P = ro.obstacle_path()                                                       # main.py:66-74
    assert P.shape == (2, 2, ro.ticks_total + 1)
    np.testing.assert_allclose(P[0, :, 100], [7.0, -15.0 + 2.0 * 1.0], atol=1e-12)
    np.testing.assert_allclose(P[1, :, 200], [14.0 + 2.0, 3.0], atol=1e-12)


def test_result_file_round_trip_in_the_readers_convention(tmp_path):
    """Nested lists under the reference's keys; read back with np.reshape(..., order='F') as draw_video.py:44-56 does."""
    nx, nVeh, nObst, Hp, Nsim, tt = 6, 3, 0, 10, 4, 160
    rng = np.random.default_rng(1)
    A = {"vehiclePathFullRes": rng.normal(size=(nx, nVeh, tt + 1)), "obstaclePathFullRes": np.zeros((nObst, 2, tt + 1)),
         "controlPathFullRes": rng.normal(size=(nVeh, tt + 1)), "controlPredictions": rng.normal(size=(Hp, nVeh, Nsim)),
         "trajectoryPredictions": rng.normal(size=(Hp, 2, nVeh, Nsim)), "initial_pos": rng.normal(size=(2, nVeh, Nsim)),
         "ReferenceTrajectory": rng.normal(size=(Hp, 2, nVeh, Nsim)),
         "MPC_delay_compensation_trajectory": rng.normal(size=(10, nx, nVeh, Nsim)),
         "evaluations_obj_value": rng.normal(size=Nsim), "controllerRuntime": rng.random((Nsim, 1)), "stepTime": rng.random((Nsim, 1))}
    path = tmp_path / "Circle_num_3_control_SCP.json"
    with open(path, "w") as f:
        json.dump({k: A[k].tolist() for k in results.RESULT_KEYS}, f)
    L = results.load_result(str(path), nx=nx, nVeh=nVeh, nObst=nObst, Hp=Hp, Nsim=Nsim, ticks_total=tt)
    for k in results.RESULT_KEYS:
        np.testing.assert_array_equal(L[k].ravel(), A[k].ravel())
    assert L["initial_pos"].shape == (1, 2, nVeh, Nsim)


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="the reference tree is only present in the build container")
def test_keys_are_the_references_keys_in_its_order():
    src = open(REF_MAIN).read()
    blk = src[src.index("result_for_plot1 = {"):src.index("# result_for_plot = [")]
    assert tuple(re.findall(r"'(\w+)'\s*:", blk)) == results.RESULT_KEYS
