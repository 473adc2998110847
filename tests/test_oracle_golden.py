"""The oracle against golden vectors produced by the reference's OWN code (oracle/make_golden.py),
and the survey's anchors.  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_setup_inputs, load_golden

STEP_FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*_step*.npz")))


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max())


def test_golden_files_present():
    assert "circle8_hp10_step0.npz" in STEP_FILES and "circle8_hp10_step6.npz" in STEP_FILES
    assert os.path.exists(os.path.join(GOLDEN, "circle8_hp10_run.npz"))


def test_survey_anchors(oracle):
    """Golden anchors of SURVEY.md section 8(a): default scenario, MPC step 0, vehicle 0."""
    G = load_golden("circle8_hp10_step0.npz")
    np.testing.assert_allclose(G["x0"][0], [-19.996979772, -19.996979772, 0.7853981634, 4, 0, 0], atol=2e-9)
    out = oracle.mpc_setup(*golden_setup_inputs(G), Hp=10, dt=float(G["sc_dt"]))
    abe = out["abe"][0, 0]
    Ad, Bd, Ed = abe[:36].reshape(6, 6), abe[36:42], abe[42:48]
    np.testing.assert_allclose(Bd, [-1.2556973014, 1.2556973014, 1.7754797876, 0, 0, 0.9816843611], atol=2e-10)
    np.testing.assert_allclose(Ed, [0.8885765876, -0.8885765876, 0, 0, 0, 0], atol=2e-10)
    np.testing.assert_allclose([Ad[0, 2], Ad[0, 5], Ad[2, 5], Ad[5, 5]],
                               [-1.1313708499, -0.6410126528, 0.5774613889, 0.0183156389], atol=2e-10)
    g = out["g"][0, 0]
    np.testing.assert_allclose(g[:3, 0], [-1.2556973014, -3.8936954742, -6.5553040255], atol=2e-10)
    np.testing.assert_allclose(g[9, 0], -25.1896392169, atol=2e-10)
    H = out["H"][0, 0]
    np.testing.assert_allclose([H[0, 0], H[9, 9], H[0, 1]], [32771.3866433744, 4063.071028513, 25523.513152965],
                               rtol=1e-12)
    np.testing.assert_allclose(out["ref"][0, 0, 0], [-18.86560892205709, -18.865608922057085], rtol=1e-14)


@pytest.mark.parametrize("fname", STEP_FILES)
def test_setup_matches_reference(oracle, fname):
    """a1-a5, a14: every MPCclass / IterClass intermediate of the reference, to 1e-12 relative."""
    G = load_golden(fname)
    nVeh, Hp = int(G["sc_nVeh"]), int(G["sc_Hp"])
    out = oracle.mpc_setup(*golden_setup_inputs(G), Hp=Hp, dt=float(G["sc_dt"]))
    assert out["rc"] == 0
    for v in range(nVeh):
        abe = out["abe"][0, v]
        assert rel(abe[:36].reshape(6, 6), G["A"][:, :, v]) < 1e-12
        assert rel(abe[36:42], G["B"][:, v]) < 1e-12
        assert rel(abe[42:48], G["E"][:, v]) < 1e-12
        MB = G["Mathcal_B"][:, :, v]                       # [2Hp, Hp], Toeplitz (F7)
        g_ref = MB[:, 0].reshape(Hp, 2)
        assert rel(out["g"][0, v], g_ref) < 1e-12
        for i in range(Hp):                                # the whole Mathcal_B, not only its first column
            for j in range(i + 1):
                assert np.abs(MB[2 * i:2 * i + 2, j] - out["g"][0, v, i - j]).max() <= 1e-12 * np.abs(MB).max()
        assert np.count_nonzero(np.triu(MB.reshape(Hp, 2, Hp)[:, 0, :], 1)) == 0
        assert rel(out["cterm"][0, v].ravel(), G["const_term"][:, v]) < 1e-12
        assert rel(out["ref"][0, v], G["RefPts"][:, :, v]) < 1e-13
        assert rel(out["H"][0, v], G["Phi_0"][:, :, v]) < 1e-12
        # Psi_0 = -2 B'Q(Ref - c) cancels to rounding noise when a vehicle sits on its reference (step 0):
        # measure against the magnitude of the terms that are summed, not of the result
        pscale = 2 * G["sc_Q_final"][v] * Hp * np.abs(g_ref).max() * np.abs(G["RefPts"][:, :, v]).max()
        assert np.abs(out["qv"][0, v] - G["Psi_0"][:, v]).max() < 1e-13 * pscale
    gscale = (G["sc_Q_final"].max() * Hp * nVeh) * np.abs(G["RefPts"]).max() ** 2
    assert abs(out["gamma0"][0] - G["gamma_0"].sum()) < 1e-13 * gscale


def _obst(G):
    """Obstacle arguments of a golden record (the frog scenario, Scenarios.py:127-146); {} for nObst = 0."""
    if "sc_nObst" in G and int(G["sc_nObst"]) > 0:
        return dict(dsafe_obst=G["sc_dsafeObstacles"], obst=G["obst"])
    return {}


def _setup(oracle, G):
    Hp = int(G["sc_Hp"])
    out = oracle.mpc_setup(*golden_setup_inputs(G), Hp=Hp, dt=float(G["sc_dt"]))
    return {k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in out.items()}


@pytest.mark.parametrize("fname", STEP_FILES)
def test_dense_assembly_matches_reference(oracle, fname):
    """a6/a7: dense P,q,Aineq,bineq,lb,ub of the reference's optimisation log, same ubar."""
    G = load_golden(fname)
    S = _setup(oracle, G)
    its = sorted(int(k.split("_")[1]) for k in G if k.startswith("Aineq_"))
    assert its
    for it in its:
        ubar = G["prev_u"][it]
        P, q, A, b, lb, ub = oracle.assemble_dense(S["g"], S["cterm"], S["H"], S["qv"], ubar, G["sc_dsafeVehicles"],
                                                   float(G["sc_dsafeExtra"]), float(G["sc_uLim"]), **_obst(G))
        assert rel(P, G[f"P_{it}"]) < 1e-12
        assert rel(q, G[f"q_{it}"]) < 1e-11
        scale = np.abs(G[f"Aineq_{it}"]).max()
        assert np.abs(A - G[f"Aineq_{it}"]).max() < 1e-11 * scale
        # causal sparsity (F9): every structural zero of the reference is a zero here; the converse may
        # differ only by rounding noise (head-on pairs have dbar orthogonal to g up to 1e-14)
        assert (A[G[f"Aineq_{it}"] == 0] == 0).all()
        assert np.abs(b - G[f"bineq_{it}"]).max() < 1e-11 * np.abs(G[f"bineq_{it}"]).max()
        np.testing.assert_array_equal(lb, G[f"lb_{it}"])
        np.testing.assert_array_equal(ub, G[f"ub_{it}"])


@pytest.mark.parametrize("fname", STEP_FILES)
def test_evaluate_matches_reference(oracle, fname):
    """a9: QCQP_evaluate on the final u."""
    G = load_golden(fname)
    S = _setup(oracle, G)
    # obstacle_mode=1: the reference's nesting of the obstacle check inside the v2 loop (SCP_controller.py:249-263)
    ev = oracle.qcqp_evaluate(S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], G["u_final"], G["sc_dsafeVehicles"],
                              float(G["sc_dsafeExtra"]), obstacle_mode=1, **_obst(G))
    assert ev["feasible"] == bool(G["eval_feasible"])
    assert abs(ev["obj"] - float(G["eval_obj"])) <= 1e-9 * max(1.0, abs(float(G["eval_obj"])))
    assert abs(ev["max_violation"] - float(G["eval_max_violation"])) < 1e-9
    assert abs(ev["sum_violations"] - float(G["eval_sum_violations"])) < 1e-9
    fin = np.isfinite(G["eval_ci"])
    assert (np.isfinite(ev["ci"]) == fin).all()
    if fin.any():                                                  # a single vehicle has no pair constraints
        assert np.abs(ev["ci"][fin] - G["eval_ci"][fin]).max() < 1e-9
    # a11 forward_U
    pos = oracle.forward(S["g"], S["cterm"], G["u_final"])          # [nVeh,Hp,2]
    assert np.abs(np.transpose(pos, (1, 2, 0)) - G["Traj"]).max() < 1e-10


@pytest.mark.parametrize("fname", [f for f in STEP_FILES if "hp50" not in f])
def test_scp_iterations_teacher_forced(oracle, fname):
    """a7+a8 per SCP iteration with the reference's own linearisation point (SURVEY hard part 3): the QP
    assembled and solved (extended precision on both sides) about the same ubar gives the same x."""
    G = load_golden(fname)
    S = _setup(oracle, G)
    for it in range(int(G["scp_iters"])):
        P, q, A, b, lb, ub = oracle.assemble_dense(S["g"], S["cterm"], S["H"], S["qv"], G["prev_u"][it],
                                                   G["sc_dsafeVehicles"], float(G["sc_dsafeExtra"]), float(G["sc_uLim"]),
                                                   **_obst(G))
        r = oracle.qp_boxed(P, q, A, b, lb, ub, opts=dict(abstol=1e-13, reltol=1e-13, feastol=1e-13), quad=True)
        assert r["status"] == 0
        assert np.abs(r["x"] - G["x"][it]).max() < 1e-11
        fval = r["fval"] + S["gamma0"]
        assert abs(fval - G["SCP_ObjVal"][it]) <= 1e-9 * max(1.0, abs(G["SCP_ObjVal"][it]))


@pytest.mark.parametrize("fname", [f for f in STEP_FILES if "hp50" not in f])
def test_scp_loop_free_running(oracle, fname):
    """a10-a12: the whole SCP loop from the reference's warm start.  The symmetric conflict (step 6) is an
    unstable fixed point of the SCP map: a 1e-14 difference in iteration 0 grows ~300x per iteration until
    the symmetry breaks, so intermediate iterates are NOT comparable free-running; the iteration count, the
    converged u and the logged objective are."""
    G = load_golden(fname)
    S = _setup(oracle, G)
    r = oracle.scp_optimizer(S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], G["sc_dsafeVehicles"], G["u_warm"],
                             dsafeExtra=float(G["sc_dsafeExtra"]), uLim=float(G["sc_uLim"]),
                             opts=dict(abstol=1e-13, reltol=1e-13, feastol=1e-13), quad=True, obstacle_mode=1, **_obst(G))
    assert bool(r["log"][-1, 5]) == bool(G["feasible"][-1])
    if int(G["scp_iters"]) <= 5:                              # well away from the symmetric instability
        assert r["iters"] == int(G["scp_iters"])
        assert np.abs(r["u_hist"] - G["x"][:, :-1]).max() < 1e-8
        assert np.abs(r["u"] - G["u_final"]).max() < 1e-8
        np.testing.assert_allclose(r["log"][-1, 2], G["QCQP_ObjVal"][-1], rtol=1e-8, atol=1e-9)
    else:                                                     # symmetry breaking amplifies rounding noise
        assert abs(r["iters"] - int(G["scp_iters"])) <= 2
        if np.abs(r["u"] - G["u_final"]).max() < 1e-6:        # same branch: same optimum
            np.testing.assert_allclose(r["log"][-1, 2], G["QCQP_ObjVal"][-1], rtol=1e-7)


def test_sampler_end_quirk(oracle):
    """SampleReferTraj.py:26-28 never advances the index: past the polyline end the samples oscillate."""
    poly = np.array([[0.0, 0.0], [30.0, 0.0]])
    pts = oracle.sample_reference(6, poly, 25.0, 0.3, 1.6)
    np.testing.assert_allclose(pts[:, 0], [26.6, 28.2, 29.8, 31.4, 30.2, 31.4], atol=1e-12)
    np.testing.assert_allclose(pts[:, 1], 0.0, atol=1e-15)


def test_full_run_invariants():
    """F4: the reference's own closed loop with exact QP solutions: 124 QPs, 50/50 feasible, 3.0661 m."""
    R = load_golden("circle8_hp10_run.npz")
    assert int(R["qp_total"]) == 124
    assert R["feasible_last"].all() and len(R["feasible_last"]) == 50
    assert abs(float(R["min_distance"]) - 3.0661) < 5e-5
    its = R["scp_iters"]
    assert (its[:6] == 1).all() and its[6] == 12
