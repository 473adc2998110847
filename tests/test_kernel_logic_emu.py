"""CPU tests of the CUDA kernels' LOGIC: the device source (csrc/*.cuh) compiled with g++ by tests/emu and run
phase by phase on the host, checked against the oracle and the reference's golden vectors.  The parity tests
proper run the real kernels on a B200 (tests/test_gpu_parity.py, -m gpu)."""
import glob
import importlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_setup_inputs, load_golden
from emu import emu

capi = importlib.import_module("senquential-convex-programming-for-trajectory-planning_b200._capi")
ALL_STEP_FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*_step*.npz")))
STEP_FILES = [f for f in ALL_STEP_FILES if "frog" not in f]          # vehicle-pair rows only
FROG_FILES = [f for f in ALL_STEP_FILES if "frog" in f]              # the reference's obstacle scenario (nVeh = 1, nObst = 22)
FAST_FILES = [f for f in STEP_FILES if "hp50" not in f and "hp20" not in f]


def params_for(G, **kw):
    p = capi.default_params_py()
    p.dt, p.uLim, p.dsafeExtra = float(G["sc_dt"]), float(G["sc_uLim"]), float(G["sc_dsafeExtra"])
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@pytest.fixture(autouse=True)
def _reset_emu():
    emu.config(nt=128, reverse=False)
    yield


@pytest.mark.parametrize("fname", STEP_FILES)
@pytest.mark.parametrize("reverse", [False, True])
def test_setup_kernel_vs_oracle_and_golden(oracle, fname, reverse):
    G = load_golden(fname)
    Hp = int(G["sc_Hp"])
    emu.config(nt=64, reverse=reverse)
    x0, u0, veh, poly = golden_setup_inputs(G)
    E = emu.mpc_setup(x0, u0, veh, poly, Hp, params_for(G))
    O = oracle.mpc_setup(x0, u0, veh, poly, Hp=Hp, dt=float(G["sc_dt"]))
    assert (E["setup_status"] == 0).all()
    for k in ["ref", "g", "cterm", "H", "abe"]:
        scale = np.abs(O[k]).max()
        assert np.abs(E[k] - O[k]).max() <= 1e-12 * scale, k
    assert np.abs(E["qv"] - O["qv"]).max() <= 1e-12 * np.abs(O["g"]).max() * np.abs(O["ref"]).max() * 2 * 20 * Hp
    assert abs(E["gamma0"][0] - O["gamma0"][0]) <= 1e-12 * max(1.0, abs(O["gamma0"][0])) * Hp
    # and straight against the reference's own arrays
    nVeh = int(G["sc_nVeh"])
    for v in range(nVeh):
        assert np.abs(E["H"][0, v] - G["Phi_0"][:, :, v]).max() <= 1e-12 * np.abs(G["Phi_0"]).max()
        assert np.abs(E["cterm"][0, v].ravel() - G["const_term"][:, v]).max() <= 1e-12 * np.abs(G["const_term"]).max()


def test_setup_kernel_noise_matches_oracle(oracle):
    G = load_golden("circle8_hp10_step10.npz")
    x0, u0, veh, poly = golden_setup_inputs(G)
    B = 3
    x0, u0, veh, poly = (np.repeat(a, B, axis=0) for a in (x0, u0, veh, poly))
    p = params_for(G, noise_sigma=3e-6, seed=12345, instance0=7, noise_counter=3)
    E = emu.mpc_setup(x0, u0, veh, poly, 10, p)
    O = oracle.mpc_setup(x0, u0, veh, poly, Hp=10, dt=0.4, noise_sigma=3e-6, seed=12345, instance0=7, noise_counter=3)
    assert np.abs(E["cterm"] - O["cterm"]).max() <= 1e-12 * np.abs(O["cterm"]).max()
    assert np.abs(E["abe"] - O["abe"]).max() <= 1e-12 * np.abs(O["abe"]).max()
    assert np.abs(E["cterm"][0] - E["cterm"][1]).max() > 1e-9          # instances draw different noise
    N = oracle.mpc_setup(x0, u0, veh, poly, Hp=10, dt=0.4)
    assert 1e-9 < np.abs(E["cterm"] - N["cterm"]).max() < 1e-3          # and it is small


def _setup(oracle, G):
    out = oracle.mpc_setup(*golden_setup_inputs(G), Hp=int(G["sc_Hp"]), dt=float(G["sc_dt"]))
    return out


@pytest.mark.parametrize("fname", STEP_FILES)
@pytest.mark.parametrize("reverse", [False, True])
def test_assemble_kernel_vs_reference(oracle, fname, reverse):
    G = load_golden(fname)
    S = _setup(oracle, G)
    emu.config(nt=256, reverse=reverse)
    its = sorted(int(k.split("_")[1]) for k in G if k.startswith("Aineq_"))
    for it in its:
        P, q, A, b, lb, ub = emu.assemble_dense(S["g"], S["cterm"], S["H"], S["qv"], G["prev_u"][it][None],
                                                G["sc_dsafeVehicles"][None], params_for(G))
        assert not np.isnan(P).any() and not np.isnan(A).any() and not np.isnan(b).any()
        assert np.abs(P[0] - G[f"P_{it}"]).max() <= 1e-12 * np.abs(G[f"P_{it}"]).max()
        assert np.abs(A[0] - G[f"Aineq_{it}"]).max() <= 1e-11 * np.abs(G[f"Aineq_{it}"]).max()
        assert (A[0][G[f"Aineq_{it}"] == 0] == 0).all()
        assert np.abs(b[0] - G[f"bineq_{it}"]).max() <= 1e-11 * np.abs(G[f"bineq_{it}"]).max()
        qscale = 2 * 20 * int(G["sc_Hp"]) * np.abs(S["g"]).max() * np.abs(S["ref"]).max()   # terms summed in Psi_0
        assert np.abs(q[0] - G[f"q_{it}"]).max() <= 1e-13 * qscale
        np.testing.assert_array_equal(lb[0], G[f"lb_{it}"])
        np.testing.assert_array_equal(ub[0], G[f"ub_{it}"])


@pytest.mark.parametrize("fname", STEP_FILES)
def test_evaluate_kernel_vs_reference(oracle, fname):
    G = load_golden(fname)
    S = _setup(oracle, G)
    ev = emu.qcqp_evaluate(S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], G["u_final"][None],
                           G["sc_dsafeVehicles"][None], params_for(G))
    assert bool(ev["feasible"][0]) == bool(G["eval_feasible"])
    assert abs(ev["obj"][0] - float(G["eval_obj"])) <= 1e-9 * max(1.0, abs(float(G["eval_obj"])))
    assert abs(ev["max_violation"][0] - float(G["eval_max_violation"])) < 1e-9
    assert abs(ev["sum_violations"][0] - float(G["eval_sum_violations"])) < 1e-9
    fin = np.isfinite(G["eval_ci"])
    assert (np.isfinite(ev["ci"][0]) == fin).all()
    assert np.abs(ev["ci"][0][fin] - G["eval_ci"][fin]).max() < 1e-9


def _qp_from_golden(oracle, G, S, it):
    return oracle.assemble_dense(S["g"][0], S["cterm"][0], S["H"][0], S["qv"][0], G["prev_u"][it], G["sc_dsafeVehicles"],
                                 float(G["sc_dsafeExtra"]), float(G["sc_uLim"]))


@pytest.mark.parametrize("fname", FAST_FILES)
@pytest.mark.parametrize("reverse,force_global", [(False, False), (True, False), (False, True)])
def test_dense_qp_kernel_vs_reference_solution(oracle, fname, reverse, force_global):
    """K3 on the reference's dense QPs (every SCP iteration of the golden step): x within 1e-6 of the
    extended-precision minimiser, objective within 1e-6 relative, constraints satisfied to 1e-6."""
    G = load_golden(fname)
    S = _setup(oracle, G)
    emu.config(nt=128, reverse=reverse, force_global_S=force_global)
    its = list(range(int(G["scp_iters"])))
    if reverse or force_global:
        its = its[:2]
    qps = [_qp_from_golden(oracle, G, S, it) for it in its]
    P, q, A, b, lb, ub = (np.stack([qp[k] for qp in qps]) for k in range(6))
    r = emu.qp_solve_dense(P, q, A, b, lb, ub, params_for(G))
    for j, it in enumerate(its):
        xs = G["x"][it]
        assert (r["status"][j] & ~capi.ST_QP_DRES_FLOOR) == 0, (it, r["status"][j], r["iters"][j])
        assert np.abs(r["x"][j] - xs).max() < 1e-6
        f_ref = 0.5 * xs @ P[j] @ xs + q[j] @ xs
        assert abs(r["fval"][j] - f_ref) <= 1e-6 * max(1.0, abs(f_ref))
        viol = max((A[j] @ r["x"][j] - b[j]).max(), (lb[j] - r["x"][j]).max(), (r["x"][j][:-1] - ub[j][:-1]).max())
        assert viol <= 1e-6


@pytest.mark.parametrize("nt,reverse,force_global", [(128, False, False), (256, True, True), (64, False, True)])
def test_dense_qp_kernel_long_horizon_size(oracle, nt, reverse, force_global):
    """A synthetic dense QP with 34 tile columns (n1 = 265; the golden Hp = 50 case, 51 columns, is too slow for the
    emulator): tile tables beyond their precomputed range, normal matrix in the global workspace, three CTA widths."""
    rng = np.random.default_rng(7)
    n1, mc = 265, 30
    M = rng.standard_normal((n1, n1)) / np.sqrt(n1)
    P = M @ M.T + np.diag(rng.uniform(0.5, 2.0, n1))
    q = rng.standard_normal(n1)
    A = rng.standard_normal((mc, n1)) * (rng.uniform(size=(mc, n1)) < 0.3)
    b = rng.uniform(0.1, 1.0, mc)
    lb, ub = -np.full(n1, 0.4), np.full(n1, 0.4)
    emu.config(nt=nt, reverse=reverse, force_global_S=force_global)
    prm = capi.default_params_py()
    prm.qp_abstol = prm.qp_reltol = 1e-10         # the defaults are tuned to the SCP QP's curvature (R = 4000); this P is O(1)
    r = emu.qp_solve_dense(P[None], q[None], A[None], b[None], lb[None], ub[None], prm)
    o = oracle.qp_boxed(P, q, A, b, lb, ub, opts=dict(abstol=1e-10, reltol=1e-10, feastol=1e-9))
    assert o["status"] == 0 and (r["status"][0] & ~capi.ST_QP_DRES_FLOOR) == 0, (o["status"], r["status"][0])
    assert np.abs(r["x"][0] - o["x"]).max() < 1e-7
    assert (np.abs(o["x"]) > 0.399).sum() > 5            # the box is active somewhere: the scaled normal matrix is not benign


@pytest.mark.parametrize("fname", FAST_FILES)
def test_dense_and_structured_solvers_agree_with_oracle_iterations(oracle, fname):
    """The kernel's interior-point iteration follows coneqp: same iteration count (+-1) as the oracle run in
    double with the same tolerances and regularisation off where that converges."""
    G = load_golden(fname)
    S = _setup(oracle, G)
    P, q, A, b, lb, ub = _qp_from_golden(oracle, G, S, 0)
    p = params_for(G, qp_abstol=1e-8, qp_reltol=1e-8, qp_feastol=1e-8)
    r = emu.qp_solve_dense(P[None], q[None], A[None], b[None], lb[None], ub[None], p)
    o = oracle.qp_boxed(P, q, A, b, lb, ub, opts=dict(abstol=1e-8, reltol=1e-8, feastol=1e-8))
    assert o["status"] == 0 and (r["status"][0] & ~capi.ST_QP_DRES_FLOOR) == 0
    assert abs(int(r["iters"][0]) - o["iterations"]) <= 1
    assert np.abs(r["x"][0] - o["x"]).max() < 1e-7


@pytest.mark.parametrize("fname", FAST_FILES)
@pytest.mark.parametrize("reverse", [False, True])
def test_scp_kernel_teacher_forced(oracle, fname, reverse):
    """K4 with max_scp_iter = 1 from the reference's own linearisation points: one QP per instance, batch =
    the SCP iterations of the golden step.  u within 1e-6 (bar 1e-5), objective 1e-6, slack 1e-6."""
    G = load_golden(fname)
    S = _setup(oracle, G)
    nit = int(G["scp_iters"])
    emu.config(nt=128, reverse=reverse)
    rep = lambda a: np.repeat(a, nit, axis=0)
    p = params_for(G, max_scp_iter=1)
    r = emu.scp_solve(rep(S["g"]), rep(S["cterm"]), rep(S["H"]), rep(S["qv"]), rep(S["gamma0"]),
                      rep(G["sc_dsafeVehicles"][None]), G["prev_u"][:nit], p)
    for it in range(nit):
        assert np.abs(r["u"][it] - G["x"][it][:-1]).max() < 1e-6, it
        assert abs(r["log"][it, 0, 0] - G["slack"][it]) < 1e-6
        assert abs(r["log"][it, 0, 1] - G["SCP_ObjVal"][it]) <= 1e-6 * max(1.0, abs(G["SCP_ObjVal"][it]))
        assert abs(r["log"][it, 0, 2] - G["QCQP_ObjVal"][it]) <= 1e-6 * max(1.0, abs(G["QCQP_ObjVal"][it]))
        assert bool(r["log"][it, 0, 5]) == bool(G["feasible"][it])
        assert (r["status"][it] & (capi.ST_QP_MAXITER | capi.ST_QP_PIVOT)) == 0


@pytest.mark.parametrize("slots", [0, 1, 3])
def test_scp_kernel_pair_block_scratch_variants(oracle, slots):
    """The normal-matrix pair blocks are formed per warp on the tensor path (mode > 0) or entry by entry (mode 0, horizons
    beyond the accumulator budget); both must give the same iterates."""
    G = load_golden("circle8_hp10_step10.npz")
    S = _setup(oracle, G)
    nit = int(G["scp_iters"])
    rep = lambda a: np.repeat(a, nit, axis=0)
    p = params_for(G, max_scp_iter=1)
    args = (rep(S["g"]), rep(S["cterm"]), rep(S["H"]), rep(S["qv"]), rep(S["gamma0"]), rep(G["sc_dsafeVehicles"][None]),
            G["prev_u"][:nit], p)
    emu.config(nt=128, alpha_slots=-1)
    base = emu.scp_solve(*args)
    emu.config(nt=128, alpha_slots=slots)
    r = emu.scp_solve(*args)
    assert np.abs(r["u"] - base["u"]).max() < 1e-9
    assert (r["ipm_iters"] == base["ipm_iters"]).all()
    for it in range(nit):
        assert np.abs(r["u"][it] - G["x"][it][:-1]).max() < 1e-6


@pytest.mark.parametrize("fname", FAST_FILES)
def test_scp_kernel_free_running(oracle, fname):
    """K4 free-running from the reference's warm start: iteration count and converged u (where the SCP map is
    stable; see test_oracle_golden.test_scp_loop_free_running), trajectories within 1e-4 m."""
    G = load_golden(fname)
    S = _setup(oracle, G)
    r = emu.scp_solve(S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], G["sc_dsafeVehicles"][None], G["u_warm"][None],
                      params_for(G))
    assert bool(r["log"][0, r["scp_iters"][0] - 1, 5]) == bool(G["feasible"][-1])
    if int(G["scp_iters"]) <= 5:
        assert r["scp_iters"][0] == int(G["scp_iters"])
        assert np.abs(r["u"][0] - G["u_final"]).max() < 1e-6
        assert np.abs(r["traj"][0] - G["Traj"]).max() < 1e-4
        assert np.abs(r["U"][0] - G["U"]).max() < 1e-6
    else:
        assert abs(int(r["scp_iters"][0]) - int(G["scp_iters"])) <= 2
    assert (r["status"][0] & (capi.ST_SCP_MAXITER | capi.ST_INFEASIBLE)) == 0


@pytest.mark.parametrize("quantum", [1, 3])
def test_scp_kernel_park_and_resume_is_bit_identical(oracle, quantum, monkeypatch):
    """The work-queue scheduler runs an instance `quantum` SCP iterations at a time and parks it in between (u + five
    scalars in global memory).  The parked / resumed run must reproduce the uninterrupted one bit for bit."""
    G = load_golden("circle8_hp10_step6.npz")        # 12 SCP iterations in the reference's run
    S = _setup(oracle, G)
    args = (S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], G["sc_dsafeVehicles"][None], G["u_warm"][None], params_for(G))
    monkeypatch.delenv("SCPB200_EMU_QUANTUM", raising=False)
    a = emu.scp_solve(*args)
    monkeypatch.setenv("SCPB200_EMU_QUANTUM", str(quantum))
    b = emu.scp_solve(*args)
    assert a["scp_iters"][0] > quantum
    for k in ("u", "traj", "U", "log", "scp_iters", "ipm_iters", "status", "obj", "max_violation"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.parametrize("fname", ["circle8_hp10_step6.npz", "circle3_hp10_step8.npz"])
def test_resumed_invocation_reads_nothing_left_over(oracle, fname, monkeypatch):
    """A parked instance is resumed by whichever CTA pops it, in a working set that last served another instance.
    With every invocation started from a working set full of NaNs (SCPB200_EMU_POISON) the results must still be those
    of the uninterrupted run, bit for bit: nothing an invocation reads is left over from the one before."""
    G = load_golden(fname)
    S = _setup(oracle, G)
    args = (S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], G["sc_dsafeVehicles"][None], G["u_warm"][None], params_for(G))
    monkeypatch.delenv("SCPB200_EMU_QUANTUM", raising=False)
    monkeypatch.delenv("SCPB200_EMU_POISON", raising=False)
    a = emu.scp_solve(*args)
    monkeypatch.setenv("SCPB200_EMU_QUANTUM", "1")
    monkeypatch.setenv("SCPB200_EMU_POISON", "1")
    b = emu.scp_solve(*args)
    assert not np.isnan(b["u"]).any()
    for k in ("u", "traj", "U", "scp_iters", "ipm_iters", "status", "obj", "max_violation"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_plant_step_vs_reference_run(oracle):
    """Clamp + plant integration of main.py:104-109, 164-191 against the reference's own 50-step run: from the
    reference's measured state, actuated command and raw controller output of step i, the next measured state
    (dopri5 at 1e-8 in the reference) and the clamped command."""
    R = load_golden("circle8_hp10_run.npz")
    G0 = load_golden("circle8_hp10_step0.npz")
    n = R["x_measured"].shape[0] - 2            # the last transition is truncated at the end of the simulated timespan
    veh = np.repeat(np.stack([G0["sc_Lf"], G0["sc_Lr"], G0["sc_Q"], G0["sc_Q_final"], G0["sc_R"]], axis=1)[None], n, 0)
    U = R["u_final"][:n].reshape(n, 8, 10).transpose(0, 2, 1)                  # U[b, k, v] = u[v*Hp + k]
    x, ua, umax, uc = emu.plant_step(R["x_measured"][:n], R["u0"][:n], veh, U, float(R["sc_mechanicalSteeringLimit"]),
                                     float(R["sc_lateralAccelerationLimit"]), float(R["sc_duLim"]), float(R["sc_dt"]), 64,
                                     params_for(G0))
    assert np.abs(x - R["x_measured"][1:n + 1]).max() < 2e-7
    assert np.abs(ua - R["u0"][1:n + 1]).max() < 1e-12
    assert np.abs(uc - R["U_clamped"][:n]).max() < 1e-12
    assert np.abs(umax - float(R["sc_mechanicalSteeringLimit"])).max() < 1e-15    # v = 4 m/s: the mechanical limit binds


def _oracle_teacher(oracle, G, S, **kw):
    """Linearisation points and solutions of the oracle's own SCP run with the given extensions."""
    tight = dict(abstol=1e-10, reltol=1e-10, feastol=1e-9)
    O = oracle.scp_optimizer(S["g"][0], S["cterm"][0], S["H"][0], S["qv"][0], float(S["gamma0"][0]), G["sc_dsafeVehicles"],
                             G["u_warm"], dsafeExtra=float(G["sc_dsafeExtra"]), uLim=float(G["sc_uLim"]), opts=tight, **kw)
    u0 = np.array(G["u_warm"], dtype=float).ravel().copy()
    if abs(u0[0]) < 2.220446049250313e-16:
        u0[0] = 2.220446049250313e-16                                   # SCP_controller.py:75-76
    ubars = np.vstack([u0[None], O["u_hist"][:-1]])
    return O, ubars


def test_trust_region_rows_vs_oracle(oracle):
    """BASELINE config 4's extension: |u - ubar|_inf <= rho folded into the box (scpb200_params.trust_radius), teacher-forced
    against the oracle's own trust-region run."""
    G = load_golden("circle8_hp10_step6.npz")
    S = _setup(oracle, G)
    rho = 0.2 * float(G["sc_uLim"])
    O, ubars = _oracle_teacher(oracle, G, S, trust_radius=rho, max_scp_iter=12)
    nit = len(ubars)
    rep = lambda a: np.repeat(a, nit, axis=0)
    r = emu.scp_solve(rep(S["g"]), rep(S["cterm"]), rep(S["H"]), rep(S["qv"]), rep(S["gamma0"]), rep(G["sc_dsafeVehicles"][None]),
                      ubars, params_for(G, max_scp_iter=1, trust_radius=rho))
    assert np.abs(r["u"] - O["u_hist"]).max() < 1e-6
    assert (np.abs(r["u"] - ubars) <= rho + 1e-9).all()
    assert np.abs(r["u"] - ubars).max() > 0.99 * rho                      # the trust region binds on this step


@pytest.mark.parametrize("fname,nt", [("circle8_hp10_step29.npz", 128), ("circle8_hp10_step10.npz", 256), ("circle3_hp10_step8.npz", 128)])
def test_rate_rows_vs_oracle(oracle, fname, nt):
    """The steering-rate rows of scpb200_params.enable_rate_rows (north-star item 2; the reference only clamps after the
    solve, main.py:164-174): |u_v[k] - u_v[k-1]| <= duLim with u_v[-1] = the command being actuated, as a tridiagonal term
    of the normal matrix.  Teacher-forced against the oracle's run with the same rows as dense rows; the bound is chosen so
    that it binds."""
    if fname not in ALL_STEP_FILES:
        pytest.skip("fixture not present")
    emu.config(nt=nt, reverse=False)
    G = load_golden(fname)
    S = _setup(oracle, G)
    nVeh, Hp = S["g"].shape[1], S["g"].shape[2]
    u_prev = np.array(G["u0"], dtype=float).reshape(-1)[:nVeh]
    # unconstrained (in rate) solution of the first QP: how fast does it steer?
    O0, ub0 = _oracle_teacher(oracle, G, S, max_scp_iter=1)
    steps0 = np.abs(np.diff(np.concatenate([u_prev[:, None], O0["u_hist"][0].reshape(nVeh, Hp)], axis=1), axis=1))
    du = 0.5 * steps0.max()
    assert du > 1e-5
    O, ubars = _oracle_teacher(oracle, G, S, max_scp_iter=8, u_prev=u_prev, duLim=du)
    nit = len(ubars)
    rep = lambda a: np.repeat(a, nit, axis=0)
    r = emu.scp_solve(rep(S["g"]), rep(S["cterm"]), rep(S["H"]), rep(S["qv"]), rep(S["gamma0"]), rep(G["sc_dsafeVehicles"][None]),
                      ubars, params_for(G, max_scp_iter=1, enable_rate_rows=1, duLim=du), u_prev=rep(u_prev[None]))
    assert (r["status"] & 3 == 0).all()
    assert np.abs(r["u"] - O["u_hist"]).max() < 1e-6
    steps = np.abs(np.diff(np.concatenate([rep(u_prev[None])[:, :, None], r["u"].reshape(nit, nVeh, Hp)], axis=2), axis=2))
    assert steps.max() <= du + 1e-8
    assert steps.max() > du - 1e-7                                         # the rows bind
    assert np.abs(r["log"][:, 0, 1] - O["log"][:, 1]).max() <= 1e-6 * np.abs(O["log"][:, 1]).max()
    # rows off (the default): the result is that of the reference's QP, bit for bit the same as without the field
    a = emu.scp_solve(S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], G["sc_dsafeVehicles"][None], ubars[:1], params_for(G, max_scp_iter=1))
    b = emu.scp_solve(S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], G["sc_dsafeVehicles"][None], ubars[:1],
                      params_for(G, max_scp_iter=1, enable_rate_rows=0, duLim=du), u_prev=u_prev[None])
    assert np.array_equal(a["u"], b["u"])


def test_rate_rows_free_running_and_parked(oracle, monkeypatch):
    """Free-running SCP loop with rate rows: same iteration count and result as the oracle on a stable step, and the park /
    resume path (warm-start snapshot with the rate rows' slacks and multipliers) is bit-identical to the uninterrupted loop."""
    G = load_golden("circle8_hp10_step10.npz")
    S = _setup(oracle, G)
    nVeh, Hp = S["g"].shape[1], S["g"].shape[2]
    u_prev = np.array(G["u0"], dtype=float).reshape(-1)[:nVeh]
    du = float(G["sc_duLim"]) if "sc_duLim" in G else np.pi / 180 * 6
    tight = dict(abstol=1e-10, reltol=1e-10, feastol=1e-9)
    O = oracle.scp_optimizer(S["g"][0], S["cterm"][0], S["H"][0], S["qv"][0], float(S["gamma0"][0]), G["sc_dsafeVehicles"], G["u_warm"],
                             dsafeExtra=float(G["sc_dsafeExtra"]), uLim=float(G["sc_uLim"]), opts=tight, u_prev=u_prev, duLim=du)
    args = (S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], G["sc_dsafeVehicles"][None], np.array(G["u_warm"], dtype=float).reshape(1, -1))
    P = params_for(G, enable_rate_rows=1, duLim=du)
    r = emu.scp_solve(*args, P, u_prev=u_prev[None])
    assert int(r["scp_iters"][0]) == O["iters"]
    assert np.abs(r["u"][0] - O["u"]).max() < 1e-6
    monkeypatch.setenv("SCPB200_EMU_QUANTUM", "1")
    monkeypatch.setenv("SCPB200_EMU_POISON", "1")
    q = emu.scp_solve(*args, P, u_prev=u_prev[None])
    assert np.array_equal(q["u"], r["u"]) and int(q["ipm_iters"][0]) == int(r["ipm_iters"][0])


def test_obstacle_rows_vs_oracle(oracle):
    """Obstacle rows (SCP_controller.py:106-114, 321-326): two static obstacles next to the paths of the
    3-vehicle scenario, teacher-forced against the oracle's own run with the same obstacles."""
    G = load_golden("circle3_hp10_step8.npz")
    S = _setup(oracle, G)
    nVeh, Hp = int(G["sc_nVeh"]), int(G["sc_Hp"])
    pos = S["cterm"][0].reshape(nVeh, Hp, 2)
    obst = np.stack([np.repeat((pos[0, 4] + [0.8, 0.6])[None], Hp, 0), np.repeat((pos[1, 6] + [-0.7, 0.9])[None], Hp, 0)])
    dso = np.full((nVeh, 2), 1.2)
    O, ubars = _oracle_teacher(oracle, G, S, dsafe_obst=dso, obst=obst, max_scp_iter=8)
    nit = len(ubars)
    rep = lambda a: np.repeat(a, nit, axis=0)
    r = emu.scp_solve(rep(S["g"]), rep(S["cterm"]), rep(S["H"]), rep(S["qv"]), rep(S["gamma0"]), rep(G["sc_dsafeVehicles"][None]),
                      ubars, params_for(G, max_scp_iter=1), dsafe_obst=rep(dso[None]), obst=rep(obst[None]))
    assert np.abs(r["u"] - O["u_hist"]).max() < 1e-6
    base = emu.scp_solve(rep(S["g"]), rep(S["cterm"]), rep(S["H"]), rep(S["qv"]), rep(S["gamma0"]), rep(G["sc_dsafeVehicles"][None]),
                         ubars, params_for(G, max_scp_iter=1))
    assert np.abs(r["u"] - base["u"]).max() > 1e-4                        # the obstacles matter


@pytest.mark.parametrize("fname", FROG_FILES)
@pytest.mark.parametrize("reverse", [False, True])
def test_frog_scenario_obstacle_rows_vs_reference(oracle, fname, reverse):
    """The reference's own obstacle scenario (Scenarios.py:127-146; rows SCP_controller.py:106-114, 321-326): K2 against
    the dense QP the reference logged, K4 teacher-forced against its per-iteration solutions and free-running against its
    iteration count (obstacle_eval_mode = 1 reproduces the nesting of SCP_controller.py:249-263 under which a single
    vehicle's obstacle constraints are never evaluated)."""
    G = load_golden(fname)
    S = _setup(oracle, G)
    emu.config(nt=128, reverse=reverse)
    dso, obst = G["sc_dsafeObstacles"][None], G["obst"][None]
    for it in sorted(int(k.split("_")[1]) for k in G if k.startswith("Aineq_")):
        P, q, A, b, lb, ub = emu.assemble_dense(S["g"], S["cterm"], S["H"], S["qv"], G["prev_u"][it][None],
                                                G["sc_dsafeVehicles"][None], params_for(G), dsafe_obst=dso, obst=obst)
        assert np.abs(P[0] - G[f"P_{it}"]).max() <= 1e-12 * np.abs(G[f"P_{it}"]).max()
        assert np.abs(A[0] - G[f"Aineq_{it}"]).max() <= 1e-11 * np.abs(G[f"Aineq_{it}"]).max()
        assert np.abs(b[0] - G[f"bineq_{it}"]).max() <= 1e-11 * np.abs(G[f"bineq_{it}"]).max()
        np.testing.assert_array_equal(lb[0], G[f"lb_{it}"])
        np.testing.assert_array_equal(ub[0], G[f"ub_{it}"])
    nit = int(G["scp_iters"])
    rep = lambda a: np.repeat(a, nit, axis=0)
    r = emu.scp_solve(rep(S["g"]), rep(S["cterm"]), rep(S["H"]), rep(S["qv"]), rep(S["gamma0"]),
                      rep(G["sc_dsafeVehicles"][None]), G["prev_u"][:nit], params_for(G, max_scp_iter=1, obstacle_eval_mode=1),
                      dsafe_obst=rep(dso), obst=rep(obst))
    for it in range(nit):
        assert np.abs(r["u"][it] - G["x"][it][:-1]).max() < 1e-6, it
        assert abs(r["log"][it, 0, 1] - G["SCP_ObjVal"][it]) <= 1e-6 * max(1.0, abs(G["SCP_ObjVal"][it]))
    r = emu.scp_solve(S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], G["sc_dsafeVehicles"][None], G["u_warm"][None],
                      params_for(G, obstacle_eval_mode=1), dsafe_obst=dso, obst=obst)
    assert r["scp_iters"][0] == nit
    assert np.abs(r["u"][0] - G["u_final"]).max() < 1e-6
    assert np.abs(r["traj"][0] - G["Traj"]).max() < 1e-4


def test_rollout_equals_the_step_by_step_call_sequence():
    """scpb200_mpc_rollout's per-instance logic (set-up inside the solve kernel, loop closure, step accounting) against
    the per-step entry points: setup -> solve -> advance on the linear model, three MPC steps, bit for bit."""
    scen = importlib.import_module("senquential-convex-programming-for-trajectory-planning_b200.scenarios")
    cb = scen.circle_batch(3, nVeh=8, Hp=10, step_lo=5, step_hi=6)
    p = capi.default_params_py()
    p.noise_sigma, p.seed, p.noise_counter = 3e-6, 77, 4
    nsteps = 3
    R = emu.mpc_rollout(cb.x0, cb.u0, cb.veh, cb.poly, cb.dsafe, 10, p, nsteps, scen.MECH_LIMIT, scen.DU_LIM)
    x0, u0, u = cb.x0.copy(), cb.u0.copy(), np.zeros((3, 80))
    for s in range(nsteps):
        q = capi.default_params_py()
        q.noise_sigma, q.seed, q.noise_counter = 3e-6, 77, 4 + s
        S = emu.mpc_setup(x0, u0, cb.veh, cb.poly, 10, q)
        O = emu.scp_solve(S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], cb.dsafe, u, q)
        np.testing.assert_array_equal(R["scp_iters_hist"][:, s], O["scp_iters"])
        np.testing.assert_array_equal(R["status_hist"][:, s], O["status"])
        np.testing.assert_array_equal(R["U_hist"][:, s], O["U"])
        np.testing.assert_array_equal(R["x_hist"][:, s], x0)
        x0, u0 = emu.advance_linear(S["abe"], O["U"], scen.MECH_LIMIT, scen.DU_LIM, x0, u0)
        u = O["u"]
    np.testing.assert_array_equal(R["x0"], x0)
    np.testing.assert_array_equal(R["u0"], u0)
    np.testing.assert_array_equal(R["u"], u)
    np.testing.assert_array_equal(R["x_hist"][:, nsteps], x0)
    assert (R["qp_total"] == R["scp_iters_hist"].sum(axis=1)).all() and (R["qp_total"] >= nsteps).all()
