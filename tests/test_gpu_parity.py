"""Parity tests proper: the CUDA path, called through the C ABI (libscpb200.so via the ctypes layer), against the
oracle and the reference's golden vectors.  Run on a B200 with `pytest -m gpu`.

Tolerances are the north-star's: per SCP iteration (teacher-forced) relative objective <= 1e-6, |u - u*|_inf <= 1e-5,
constraint violation <= 1e-6; free-running trajectories within 1e-4 m.  The tests assert a 10x tighter 1e-6 on u."""
import ctypes as C
import glob
import importlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_setup_inputs, load_golden

pytestmark = pytest.mark.gpu

PKG = "senquential-convex-programming-for-trajectory-planning_b200"
ALL_STEP_FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*_step*.npz")))
STEP_FILES = [f for f in ALL_STEP_FILES if "frog" not in f]          # vehicle-pair rows only
FROG_FILES = [f for f in ALL_STEP_FILES if "frog" in f]              # the reference's obstacle scenario (nVeh = 1, nObst = 22)
NOT50 = [f for f in STEP_FILES if "hp50" not in f]


@pytest.fixture(scope="module")
def mods():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    capi = importlib.import_module(PKG + "._capi")
    batch = importlib.import_module(PKG + ".batch")
    scen = importlib.import_module(PKG + ".scenarios")
    return dict(torch=torch, capi=capi, batch=batch, scen=scen)


def make_batch(mods, G, B=1, **pkw):
    capi, batch = mods["capi"], mods["batch"]
    p = capi.Params()
    capi.load().scpb200_default_params(C.byref(p))
    p.dt, p.uLim, p.dsafeExtra = float(G["sc_dt"]), float(G["sc_uLim"]), float(G["sc_dsafeExtra"])
    for k, v in pkw.items():
        setattr(p, k, v)
    bs = batch.BatchSCP(B, int(G["sc_nVeh"]), int(G["sc_Hp"]), params=p)
    x0, u0, veh, poly = golden_setup_inputs(G)
    rep = lambda a: np.repeat(a, B, axis=0)
    bs.load_inputs(x0=rep(x0), u0=rep(u0), veh=rep(veh), poly=rep(poly), dsafe=rep(G["sc_dsafeVehicles"][None]))
    return bs


def host(t):
    return t.detach().cpu().numpy()


def test_library_loaded_is_the_in_tree_cuda_build(mods):
    lib = mods["capi"].load()
    assert os.path.samefile(mods["capi"].LIB_PATH, os.path.join(os.path.dirname(GOLDEN), "..", PKG, "libscpb200.so"))
    assert lib.scpb200_device_count() >= 1
    assert lib.scpb200_version() == 100


@pytest.mark.parametrize("fname", STEP_FILES)
def test_setup_kernel(mods, oracle, fname):
    """K1 vs the reference's MPCclass / IterClass arrays (golden) and the oracle, 1e-12 relative."""
    G = load_golden(fname)
    Hp, nVeh = int(G["sc_Hp"]), int(G["sc_nVeh"])
    bs = make_batch(mods, G)
    bs.setup()
    O = oracle.mpc_setup(*golden_setup_inputs(G), Hp=Hp, dt=float(G["sc_dt"]))
    assert (host(bs.setup_status) == 0).all()
    for k in ["ref", "g", "cterm", "H", "abe"]:
        assert np.abs(host(getattr(bs, k)) - O[k]).max() <= 1e-12 * np.abs(O[k]).max(), k
    qscale = 2 * 20 * Hp * np.abs(O["g"]).max() * np.abs(O["ref"]).max()
    assert np.abs(host(bs.qv) - O["qv"]).max() <= 1e-13 * qscale
    assert abs(host(bs.gamma0)[0] - O["gamma0"][0]) <= 1e-12 * max(1.0, abs(O["gamma0"][0])) * Hp
    for v in range(nVeh):
        assert np.abs(host(bs.H)[0, v] - G["Phi_0"][:, :, v]).max() <= 1e-12 * np.abs(G["Phi_0"]).max()
        assert np.abs(host(bs.cterm)[0, v].ravel() - G["const_term"][:, v]).max() <= 1e-12 * np.abs(G["const_term"]).max()
        assert np.abs(host(bs.ref)[0, v] - G["RefPts"][:, :, v]).max() <= 1e-12 * np.abs(G["RefPts"]).max()
        assert np.abs(host(bs.abe)[0, v, :36].reshape(6, 6) - G["A"][:, :, v]).max() <= 1e-12


@pytest.mark.parametrize("fname", STEP_FILES)
def test_assemble_dense_kernel(mods, fname):
    """K2 vs the dense P, q, Aineq, bineq, lb, ub the reference logged for the same ubar."""
    G = load_golden(fname)
    torch = mods["torch"]
    bs = make_batch(mods, G)
    bs.setup()
    for it in sorted(int(k.split("_")[1]) for k in G if k.startswith("Aineq_")):
        ubar = torch.as_tensor(G["prev_u"][it][None], device=bs.device)
        D = {k: host(v)[0] for k, v in bs.assemble_dense(ubar).items()}
        assert np.abs(D["P"] - G[f"P_{it}"]).max() <= 1e-12 * np.abs(G[f"P_{it}"]).max()
        assert np.abs(D["A"] - G[f"Aineq_{it}"]).max() <= 1e-11 * np.abs(G[f"Aineq_{it}"]).max()
        assert (D["A"][G[f"Aineq_{it}"] == 0] == 0).all()
        assert np.abs(D["b"] - G[f"bineq_{it}"]).max() <= 1e-11 * np.abs(G[f"bineq_{it}"]).max()
        qscale = 2 * 20 * int(G["sc_Hp"]) * np.abs(host(bs.g)).max() * np.abs(host(bs.ref)).max()
        assert np.abs(D["q"] - G[f"q_{it}"]).max() <= 1e-13 * qscale
        np.testing.assert_array_equal(D["lb"], G[f"lb_{it}"])
        np.testing.assert_array_equal(D["ub"], G[f"ub_{it}"])


@pytest.mark.parametrize("fname", STEP_FILES)
def test_evaluate_and_forward_kernels(mods, fname):
    G = load_golden(fname)
    torch = mods["torch"]
    bs = make_batch(mods, G)
    bs.setup()
    u = torch.as_tensor(G["u_final"][None], device=bs.device)
    ev = {k: (host(v) if v is not None else None) for k, v in bs.evaluate(u, want_ci=True).items()}
    assert bool(ev["feasible"][0]) == bool(G["eval_feasible"])
    assert abs(ev["obj"][0] - float(G["eval_obj"])) <= 1e-9 * max(1.0, abs(float(G["eval_obj"])))
    assert abs(ev["max_violation"][0] - float(G["eval_max_violation"])) < 1e-9
    assert abs(ev["sum_violations"][0] - float(G["eval_sum_violations"])) < 1e-9
    fin = np.isfinite(G["eval_ci"])
    assert (np.isfinite(ev["ci"][0]) == fin).all()
    assert np.abs(ev["ci"][0][fin] - G["eval_ci"][fin]).max() < 1e-9
    traj, U = bs.forward_u(u)
    assert np.abs(host(traj)[0] - G["Traj"]).max() < 1e-10
    assert np.abs(host(U)[0] - G["U"]).max() == 0.0


def _dense_qps(oracle, G, its):
    O = oracle.mpc_setup(*golden_setup_inputs(G), Hp=int(G["sc_Hp"]), dt=float(G["sc_dt"]))
    qps = [oracle.assemble_dense(O["g"][0], O["cterm"][0], O["H"][0], O["qv"][0], G["prev_u"][it], G["sc_dsafeVehicles"],
                                 float(G["sc_dsafeExtra"]), float(G["sc_uLim"])) for it in its]
    return [np.stack([qp[k] for qp in qps]) for k in range(6)]


@pytest.mark.parametrize("fname", NOT50)
def test_dense_qp_entry_vs_extended_precision_minimiser(mods, oracle, fname):
    """K3 (the CVXOPT-replacement entry) on the reference's dense QPs of every SCP iteration of the golden step."""
    G = load_golden(fname)
    torch = mods["torch"]
    its = list(range(int(G["scp_iters"])))
    P, q, A, b, lb, ub = _dense_qps(oracle, G, its)
    dev = torch.device("cuda")
    r = mods["batch"].qp_solve_dense(*[torch.as_tensor(a, device=dev) for a in (P, q, A, b, lb, ub)])
    x, fval, st, zA = host(r["x"]), host(r["fval"]), host(r["status"]), host(r["zA"])
    for j, it in enumerate(its):
        xs = G["x"][it]
        # converged, possibly accepted at the precision floor of the dual residual (the accuracy checks below hold either way)
        assert (st[j] & ~mods["capi"].ST_QP_DRES_FLOOR) == 0, (it, st[j], host(r["iters"])[j])
        assert np.abs(x[j] - xs).max() < 1e-6
        f_ref = 0.5 * xs @ P[j] @ xs + q[j] @ xs
        assert abs(fval[j] - f_ref) <= 1e-6 * max(1.0, abs(f_ref))
        assert max((A[j] @ x[j] - b[j]).max(), (lb[j] - x[j]).max(), (x[j][:-1] - ub[j][:-1]).max()) <= 1e-6
        assert (zA[j] >= 0).all()


@pytest.mark.parametrize("fname", NOT50)
def test_scp_kernel_teacher_forced(mods, fname):
    """K4, one QP per instance from the reference's own linearisation points (batch = the SCP iterations)."""
    G = load_golden(fname)
    capi = mods["capi"]
    nit = int(G["scp_iters"])
    bs = make_batch(mods, G, B=nit, max_scp_iter=1)
    bs.load_inputs(u=G["prev_u"][:nit])
    bs.controller_step()
    u, log, st = host(bs.u), host(bs.log), host(bs.status)
    for it in range(nit):
        assert np.abs(u[it] - G["x"][it][:-1]).max() < 1e-6, it
        assert abs(log[it, 0, 0] - G["slack"][it]) < 1e-6
        assert abs(log[it, 0, 1] - G["SCP_ObjVal"][it]) <= 1e-6 * max(1.0, abs(G["SCP_ObjVal"][it]))
        assert abs(log[it, 0, 2] - G["QCQP_ObjVal"][it]) <= 1e-6 * max(1.0, abs(G["QCQP_ObjVal"][it]))
        assert bool(log[it, 0, 5]) == bool(G["feasible"][it])
        assert (st[it] & (capi.ST_QP_MAXITER | capi.ST_QP_PIVOT)) == 0


def test_scp_kernel_hp50_first_iteration(mods):
    """Hp = 50 (n1 = 401, normal matrix in the L2-resident workspace): first SCP iteration, teacher-forced."""
    G = load_golden("circle8_hp50_step3.npz")
    bs = make_batch(mods, G, B=1, max_scp_iter=1)
    bs.load_inputs(u=G["prev_u"][:1])
    assert not bs.plan()["S_in_shared"]
    bs.controller_step()
    assert np.abs(host(bs.u)[0] - G["x"][0][:-1]).max() < 1e-6
    assert abs(host(bs.log)[0, 0, 1] - G["SCP_ObjVal"][0]) <= 1e-6 * max(1.0, abs(G["SCP_ObjVal"][0]))


def test_launch_shape_follows_dimensions_and_every_cta_width_agrees(mods, monkeypatch):
    """The plan depends on the problem dimensions only: 3 x 128 threads per SM at Hp = 10, one 512-thread CTA where only
    one fits (Hp = 20 shared-resident, Hp = 50 with the normal matrix in the workspace).  Forcing other CTA widths
    (different kernel instantiations, different reduction order) gives the same solution to rounding."""
    shapes = {"circle8_hp10_step6.npz": (128, True), "circle8_hp20_step5.npz": (512, True), "circle8_hp50_step3.npz": (512, False)}
    for fname, (threads, shared) in shapes.items():
        if not os.path.exists(os.path.join(GOLDEN, fname)):
            continue
        G = load_golden(fname)
        nit = min(2, int(G["scp_iters"]))
        sols = {}
        for forced in (0, 128, 256, 512):
            if forced:
                monkeypatch.setenv("SCPB200_THREADS", str(forced))
            else:
                monkeypatch.delenv("SCPB200_THREADS", raising=False)
            bs = make_batch(mods, G, B=nit, max_scp_iter=1)
            bs.load_inputs(u=G["prev_u"][:nit])
            pl = bs.plan()
            if not forced:
                assert (pl["threads"], bool(pl["S_in_shared"])) == (threads, shared), (fname, pl)
            else:
                assert pl["threads"] == forced
            bs.controller_step()
            sols[forced] = host(bs.u).copy()
            for it in range(nit):
                assert np.abs(sols[forced][it] - G["x"][it][:-1]).max() < 1e-6, (fname, forced, it)
        for forced in (128, 256, 512):
            assert np.abs(sols[forced] - sols[0]).max() < 1e-7, (fname, forced)
    monkeypatch.delenv("SCPB200_THREADS", raising=False)


def test_long_horizon_routines_on_short_horizon_goldens(mods, monkeypatch):
    """chol_factor_left / chol_solve_far (the versions of the factorisation and of the triangular sweeps for a factor in the
    L2-resident workspace, Hp = 50 in production) on the Hp = 10 / 20 goldens: SCPB200_FORCE_GLOBAL_S pushes the normal
    matrix out of shared memory; every CTA width incl. a single warp (the rows beyond the current tile then belong to lanes
    8..31 of the same warp) and 64 threads (more rows than threads: the not-prefetched remainder).  Teacher-forced against the
    reference's golden solution (1e-6) and against the shared-memory path (rounding level)."""
    for fname in ("circle8_hp10_step6.npz", "circle8_hp10_step10.npz", "circle8_hp20_step5.npz"):
        if not os.path.exists(os.path.join(GOLDEN, fname)):
            continue
        G = load_golden(fname)
        nit = min(4, int(G["scp_iters"]))
        monkeypatch.delenv("SCPB200_FORCE_GLOBAL_S", raising=False)
        monkeypatch.delenv("SCPB200_THREADS", raising=False)
        bs = make_batch(mods, G, B=nit, max_scp_iter=1)
        bs.load_inputs(u=G["prev_u"][:nit])
        assert bs.plan()["S_in_shared"]
        bs.controller_step()
        u_shared = host(bs.u).copy()
        monkeypatch.setenv("SCPB200_FORCE_GLOBAL_S", "1")
        for forced in (32, 64, 128, 256, 512):
            monkeypatch.setenv("SCPB200_THREADS", str(forced))
            bs = make_batch(mods, G, B=nit, max_scp_iter=1)
            bs.load_inputs(u=G["prev_u"][:nit])
            pl = bs.plan()
            assert not pl["S_in_shared"] and pl["threads"] == forced, (fname, pl)
            bs.controller_step()
            u = host(bs.u)
            assert (host(bs.status) & (mods["capi"].ST_QP_MAXITER | mods["capi"].ST_QP_PIVOT)).max() == 0
            for it in range(nit):
                assert np.abs(u[it] - G["x"][it][:-1]).max() < 1e-6, (fname, forced, it)
            assert np.abs(u - u_shared).max() < 1e-7, (fname, forced)
    monkeypatch.delenv("SCPB200_FORCE_GLOBAL_S", raising=False)
    monkeypatch.delenv("SCPB200_THREADS", raising=False)


@pytest.mark.parametrize("fname", NOT50)
def test_scp_kernel_free_running(mods, fname):
    G = load_golden(fname)
    capi = mods["capi"]
    bs = make_batch(mods, G)
    bs.load_inputs(u=G["u_warm"][None])
    bs.controller_step()
    its = int(host(bs.scp_iters)[0])
    assert bool(host(bs.log)[0, its - 1, 5]) == bool(G["feasible"][-1])
    if int(G["scp_iters"]) <= 5:
        assert its == int(G["scp_iters"])
        assert np.abs(host(bs.u)[0] - G["u_final"]).max() < 1e-6
        assert np.abs(host(bs.traj)[0] - G["Traj"]).max() < 1e-4
        assert np.abs(host(bs.U)[0] - G["U"]).max() < 1e-6
    else:
        assert abs(its - int(G["scp_iters"])) <= 2
    assert (host(bs.status)[0] & (capi.ST_SCP_MAXITER | capi.ST_INFEASIBLE)) == 0


def _oracle_teacher(oracle, G, S, **kw):
    tight = dict(abstol=1e-10, reltol=1e-10, feastol=1e-9)
    O = oracle.scp_optimizer(S["g"][0], S["cterm"][0], S["H"][0], S["qv"][0], float(S["gamma0"][0]), G["sc_dsafeVehicles"],
                             G["u_warm"], dsafeExtra=float(G["sc_dsafeExtra"]), uLim=float(G["sc_uLim"]), opts=tight, **kw)
    u0 = np.array(G["u_warm"], dtype=float).ravel().copy()
    if abs(u0[0]) < 2.220446049250313e-16:
        u0[0] = 2.220446049250313e-16                                   # SCP_controller.py:75-76
    return O, np.vstack([u0[None], O["u_hist"][:-1]])


@pytest.mark.parametrize("fname", ["circle8_hp10_step6.npz", "circle8_hp20_step7.npz"])
def test_trust_region_vs_oracle(mods, oracle, fname):
    """BASELINE config 4's extension (trust_radius folded into the box), teacher-forced against the oracle's own
    trust-region SCP run, then free-running with a raised iteration cap."""
    G = load_golden(fname)
    x0, u0, veh, poly = golden_setup_inputs(G)
    S = oracle.mpc_setup(x0, u0, veh, poly, Hp=int(G["sc_Hp"]), dt=float(G["sc_dt"]))
    rho = 0.2 * float(G["sc_uLim"])
    O, ubars = _oracle_teacher(oracle, G, S, trust_radius=rho, max_scp_iter=12)
    bs = make_batch(mods, G, B=len(ubars), max_scp_iter=1, trust_radius=rho)
    bs.load_inputs(u=ubars)
    bs.controller_step()
    u = host(bs.u)
    assert np.abs(u - O["u_hist"]).max() < 1e-6
    assert (np.abs(u - ubars) <= rho + 1e-9).all()
    bs = make_batch(mods, G, B=1, max_scp_iter=100, trust_radius=rho)
    bs.load_inputs(u=G["u_warm"][None])
    bs.controller_step()
    assert 1 <= int(host(bs.scp_iters)[0]) <= 100 and (host(bs.status)[0] & (mods["capi"].ST_QP_MAXITER | mods["capi"].ST_QP_PIVOT)) == 0


RATE_CASES = [f for f in ("circle8_hp10_step10.npz", "circle8_hp10_step29.npz", "circle3_hp10_step8.npz", "circle8_hp20_step7.npz")
              if f in ALL_STEP_FILES]


@pytest.mark.parametrize("fname", RATE_CASES)
def test_rate_rows_vs_oracle(mods, oracle, fname):
    """Steering-rate rows inside the QP (scpb200_params.enable_rate_rows, north-star item 2; the reference only clamps after
    the solve, main.py:164-174).  K4 carries them as a tridiagonal term of the vehicle blocks; the oracle solves the same QPs
    with the rows as dense rows.  Teacher-forced, bound chosen so that the rows bind: u within 1e-6, objective within 1e-6
    relative, every steering step within the bound; then the free-running loop against the oracle's."""
    G = load_golden(fname)
    x0, u0, veh, poly = golden_setup_inputs(G)
    nVeh, Hp = int(G["sc_nVeh"]), int(G["sc_Hp"])
    S = oracle.mpc_setup(x0, u0, veh, poly, Hp=Hp, dt=float(G["sc_dt"]))
    u_prev = np.asarray(u0, dtype=float).reshape(-1)[:nVeh]
    O0, _ = _oracle_teacher(oracle, G, S, max_scp_iter=1)
    steps0 = np.abs(np.diff(np.concatenate([u_prev[:, None], O0["u_hist"][0].reshape(nVeh, Hp)], axis=1), axis=1))
    du = 0.5 * steps0.max()
    assert du > 1e-5
    O, ubars = _oracle_teacher(oracle, G, S, max_scp_iter=8, u_prev=u_prev, duLim=du)
    nit = len(ubars)
    bs = make_batch(mods, G, B=nit, max_scp_iter=1, enable_rate_rows=1, duLim=du)
    bs.load_inputs(u=ubars)
    bs.controller_step()
    u, log, st = host(bs.u), host(bs.log), host(bs.status)
    assert (st & (mods["capi"].ST_QP_MAXITER | mods["capi"].ST_QP_PIVOT) == 0).all()
    assert np.abs(u - O["u_hist"]).max() < 1e-6
    steps = np.abs(np.diff(np.concatenate([np.repeat(u_prev[None, :, None], nit, axis=0), u.reshape(nit, nVeh, Hp)], axis=2), axis=2))
    assert steps.max() <= du + 1e-8 and steps.max() > du - 1e-7
    assert np.abs(log[:, 0, 1] - O["log"][:, 1]).max() <= 1e-6 * np.abs(O["log"][:, 1]).max()
    print(f"\n[rate rows {fname}] {nit} QPs, du = {du:.3e}: max |u-u*| {np.abs(u - O['u_hist']).max():.2e}, "
          f"largest steering step {steps.max():.6e}")
    # free-running with the reference's own bound (6 deg per step: inactive next to |u| <= 3 deg, so the run equals the
    # run without the rows) and with the binding bound against the oracle
    tight = dict(abstol=1e-10, reltol=1e-10, feastol=1e-9)
    Of = oracle.scp_optimizer(S["g"][0], S["cterm"][0], S["H"][0], S["qv"][0], float(S["gamma0"][0]), G["sc_dsafeVehicles"], G["u_warm"],
                              dsafeExtra=float(G["sc_dsafeExtra"]), uLim=float(G["sc_uLim"]), opts=tight, u_prev=u_prev, duLim=du)
    bs = make_batch(mods, G, B=1, enable_rate_rows=1, duLim=du)
    bs.load_inputs(u=np.asarray(G["u_warm"], dtype=float)[None])
    bs.controller_step()
    if Of["iters"] <= 6:
        assert int(host(bs.scp_iters)[0]) == Of["iters"]
        assert np.abs(host(bs.u)[0] - Of["u"]).max() < 1e-6
    ref_du = float(G["sc_duLim"])
    a = make_batch(mods, G, B=1)
    a.load_inputs(u=np.asarray(G["u_warm"], dtype=float)[None]); a.controller_step()
    b = make_batch(mods, G, B=1, enable_rate_rows=1, duLim=ref_du)
    b.load_inputs(u=np.asarray(G["u_warm"], dtype=float)[None]); b.controller_step()
    assert int(host(a.scp_iters)[0]) == int(host(b.scp_iters)[0])
    assert np.abs(host(a.u) - host(b.u)).max() < 1e-6


def test_rate_rows_contract_and_rollout(mods):
    """enable_rate_rows without u_prev is an argument error of the plain entries (not a silent fall-back); the rollout entry
    (u_prev = the set-up input u0 of each step) reproduces the per-step call sequence; every applied steering step of the
    closed loop respects the bound without the post-solve clamp having to act."""
    torch, capi, batch, scen = mods["torch"], mods["capi"], mods["batch"], mods["scen"]
    B, nVeh, Hp, nsteps = 24, 8, 10, 4
    du = 0.15 * scen.DU_LIM
    cb = scen.circle_batch(B, nVeh=nVeh, Hp=Hp, instance0=0, step_lo=4, step_hi=7)

    def fresh():
        p = capi.Params()
        capi.load().scpb200_default_params(C.byref(p))
        p.enable_rate_rows, p.duLim = 1, du
        bs = batch.BatchSCP(B, nVeh, Hp, params=p)
        bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((B, nVeh * Hp)))
        return bs

    bs = fresh()
    bs.setup()
    rc = bs.lib.scpb200_scp_solve(C.byref(bs.dims), C.byref(bs.params), *[C.c_void_p(t.data_ptr()) for t in (bs.g, bs.cterm, bs.H, bs.qv, bs.gamma0, bs.dsafe)],
                                  None, None, C.c_void_p(bs.u.data_ptr()), C.c_void_p(bs.traj.data_ptr()), C.c_void_p(bs.U.data_ptr()), None,
                                  *[C.c_void_p(t.data_ptr()) for t in (bs.scp_iters, bs.ipm_iters, bs.status, bs.obj, bs.max_violation, bs.ws)], None)
    assert rc == -1 and b"u_prev" in bs.lib.scpb200_last_error()
    a = fresh()
    worst = 0.0
    for s in range(nsteps):
        a.params.noise_counter = s
        u_before = a.u0.clone()
        a.setup(); a.solve()
        worst = max(worst, float((a.U[:, 0, :] - u_before).abs().max()))
        a.advance_linear(scen.MECH_LIMIT, scen.DU_LIM)
    assert worst <= du + 1e-8, worst
    assert worst > 0.5 * du                                                # the bound matters on this workload
    r = fresh()
    r.params.noise_counter = 0
    r.rollout(nsteps, scen.MECH_LIMIT, scen.DU_LIM)
    torch.cuda.synchronize()
    # (bit-identity of the two routes is pinned on the fixed-shape kernel, tests/test_gpu_workloads.py; with rate rows the
    # run-time-dimension instantiations run, whose in-kernel set-up differs from the stand-alone one in the last bit of qv)
    for k in ("scp_iters", "status"):
        assert torch.equal(getattr(a, k), getattr(r, k)), k
    for k in ("u", "U", "x0", "u0"):
        assert float((getattr(a, k) - getattr(r, k)).abs().max()) < 1e-6, k
    print(f"\n[rate rows closed loop] {B} instances x {nsteps} steps: largest applied steering step {worst:.6e} (bound {du:.6e})")


def test_obstacle_rows_vs_oracle(mods, oracle):
    """Obstacle rows (SCP_controller.py:106-114, 321-326) in K4, K2 and the evaluate kernel against the
    oracle with the same two static obstacles."""
    capi, batch = mods["capi"], mods["batch"]
    G = load_golden("circle3_hp10_step8.npz")
    x0, u0, veh, poly = golden_setup_inputs(G)
    nVeh, Hp = int(G["sc_nVeh"]), int(G["sc_Hp"])
    S = oracle.mpc_setup(x0, u0, veh, poly, Hp=Hp, dt=float(G["sc_dt"]))
    pos = S["cterm"][0].reshape(nVeh, Hp, 2)
    obst = np.stack([np.repeat((pos[0, 4] + [0.8, 0.6])[None], Hp, 0), np.repeat((pos[1, 6] + [-0.7, 0.9])[None], Hp, 0)])
    dso = np.full((nVeh, 2), 1.2)
    O, ubars = _oracle_teacher(oracle, G, S, dsafe_obst=dso, obst=obst, max_scp_iter=8)
    nit = len(ubars)
    p = capi.Params()
    capi.load().scpb200_default_params(C.byref(p))
    p.dt, p.uLim, p.dsafeExtra, p.max_scp_iter = float(G["sc_dt"]), float(G["sc_uLim"]), float(G["sc_dsafeExtra"]), 1
    bs = batch.BatchSCP(nit, nVeh, Hp, nObst=2, params=p)
    rep = lambda a: np.repeat(a, nit, axis=0)
    bs.load_inputs(x0=rep(x0), u0=rep(u0), veh=rep(veh), poly=rep(poly), dsafe=rep(G["sc_dsafeVehicles"][None]),
                   dsafe_obst=rep(dso[None]), obst=rep(obst[None]), u=ubars)
    bs.controller_step()
    assert np.abs(host(bs.u) - O["u_hist"]).max() < 1e-6
    # K2 with obstacle rows against the oracle's dense assembly at the same linearisation points
    out = bs.assemble_dense(mods["torch"].as_tensor(ubars).cuda())
    D = oracle.assemble_dense(S["g"][0], S["cterm"][0], S["H"][0], S["qv"][0], ubars[1], G["sc_dsafeVehicles"],
                              float(G["sc_dsafeExtra"]), float(G["sc_uLim"]), dsafe_obst=dso, obst=obst)
    D = D if isinstance(D, dict) else dict(zip(("P", "q", "A", "b", "lb", "ub"), D))
    for k in ("P", "q", "A", "b", "lb", "ub"):
        ref = np.asarray(D[k]).reshape(host(out[k])[1].shape)
        assert np.abs(host(out[k])[1] - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), k


@pytest.mark.parametrize("fname", FROG_FILES)
def test_frog_scenario_obstacle_rows_vs_reference(mods, fname):
    """The reference's own obstacle scenario (Scenarios.py:127-146; rows SCP_controller.py:106-114, 321-326) on the GPU:
    K2 against the dense QP the reference logged, K4 teacher-forced against its per-iteration solutions, and free-running
    against its iteration count and result (obstacle_eval_mode = 1: the nesting of SCP_controller.py:249-263)."""
    capi, batch, torch = mods["capi"], mods["batch"], mods["torch"]
    G = load_golden(fname)
    x0, u0, veh, poly = golden_setup_inputs(G)
    nVeh, Hp, nObst, nit = int(G["sc_nVeh"]), int(G["sc_Hp"]), int(G["sc_nObst"]), int(G["scp_iters"])

    def frog_batch(B, **pkw):
        p = capi.Params()
        capi.load().scpb200_default_params(C.byref(p))
        p.dt, p.uLim, p.dsafeExtra, p.obstacle_eval_mode = float(G["sc_dt"]), float(G["sc_uLim"]), float(G["sc_dsafeExtra"]), 1
        for k, v in pkw.items():
            setattr(p, k, v)
        bs = batch.BatchSCP(B, nVeh, Hp, nObst=nObst, params=p)
        rep = lambda a: np.repeat(a, B, axis=0)
        bs.load_inputs(x0=rep(x0), u0=rep(u0), veh=rep(veh), poly=rep(poly), dsafe=rep(G["sc_dsafeVehicles"][None]),
                       dsafe_obst=rep(G["sc_dsafeObstacles"][None]), obst=rep(G["obst"][None]))
        return bs

    bs = frog_batch(nit, max_scp_iter=1)
    bs.load_inputs(u=G["prev_u"][:nit])
    bs.controller_step()
    u, log = host(bs.u), host(bs.log)
    for it in range(nit):
        assert np.abs(u[it] - G["x"][it][:-1]).max() < 1e-6, it
        assert abs(log[it, 0, 1] - G["SCP_ObjVal"][it]) <= 1e-6 * max(1.0, abs(G["SCP_ObjVal"][it]))
    D = {k: host(v) for k, v in bs.assemble_dense(torch.as_tensor(G["prev_u"][:nit]).cuda()).items()}
    for it in sorted(int(k.split("_")[1]) for k in G if k.startswith("Aineq_")):
        assert np.abs(D["P"][it] - G[f"P_{it}"]).max() <= 1e-12 * np.abs(G[f"P_{it}"]).max()
        assert np.abs(D["A"][it] - G[f"Aineq_{it}"]).max() <= 1e-11 * np.abs(G[f"Aineq_{it}"]).max()
        assert np.abs(D["b"][it] - G[f"bineq_{it}"]).max() <= 1e-11 * np.abs(G[f"bineq_{it}"]).max()
        np.testing.assert_array_equal(D["lb"][it], G[f"lb_{it}"])
        np.testing.assert_array_equal(D["ub"][it], G[f"ub_{it}"])
    bs = frog_batch(1)
    bs.load_inputs(u=G["u_warm"][None])
    bs.controller_step()
    assert int(host(bs.scp_iters)[0]) == nit
    assert np.abs(host(bs.u)[0] - G["u_final"]).max() < 1e-6
    assert np.abs(host(bs.traj)[0] - G["Traj"]).max() < 1e-4


def test_closed_loop_rollout_against_reference_run(mods):
    """The reference's own 50-step closed loop (golden run): at every MPC step feed the reference's measured
    state (x0, u0) and warm start; controller outputs must match where the SCP map is stable, and the number of
    QPs solved over the run must match the reference's 124 within the symmetric-step slack."""
    R = load_golden("circle8_hp10_run.npz")
    G0 = load_golden("circle8_hp10_step0.npz")
    nsteps = R["x0"].shape[0]
    bs = make_batch(mods, G0, B=nsteps)
    veh = np.stack([G0["sc_Lf"], G0["sc_Lr"], G0["sc_Q"], G0["sc_Q_final"], G0["sc_R"]], axis=1)
    warm = np.vstack([np.zeros((1, 80)), R["u_final"][:-1]])
    bs.load_inputs(x0=R["x0"], u0=R["u0"], veh=np.repeat(veh[None], nsteps, 0), u=warm)
    bs.controller_step()
    its, u = host(bs.scp_iters), host(bs.u)
    stable = R["scp_iters"] <= 5
    assert (its[stable] == R["scp_iters"][stable]).all()
    assert np.abs(u[stable] - R["u_final"][stable]).max() < 1e-6
    assert np.abs(host(bs.traj)[stable] - R["Traj"][stable]).max() < 1e-4
    assert abs(int(its.sum()) - int(R["qp_total"])) <= 3
    assert (host(bs.status) & 8 == 0).all()                  # 50/50 feasible


def test_batch_1024_against_oracle_and_invariants(mods, oracle):
    """BASELINE config 2 size (default scenario x1024, perturbed): every instance against the CPU oracle run in
    double with the same tolerances on a subset, and size-independent invariants on the full batch."""
    scen, capi = mods["scen"], mods["capi"]
    B = 1024
    cb = scen.circle_batch(B, nVeh=8, Hp=10, step_lo=5, step_hi=7)
    p = capi.Params()
    capi.load().scpb200_default_params(C.byref(p))
    bs = mods["batch"].BatchSCP(B, 8, 10, params=p)
    bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((B, 80)))
    bs.controller_step()
    u, its, st = host(bs.u), host(bs.scp_iters), host(bs.status)
    assert ((st & (capi.ST_QP_MAXITER | capi.ST_SETUP)) == 0).all()
    assert np.abs(u).max() <= p.uLim + 1e-9                                    # box satisfied
    # evaluate kernel agrees with the solver's own bookkeeping
    ev = bs.evaluate()
    assert np.abs(host(ev["obj"]) - host(bs.obj)).max() <= 1e-9 * max(1.0, np.abs(host(bs.obj)).max())
    # idempotence: re-solving from the converged u stops after one QP for converged instances and keeps u
    conv = (st & capi.ST_SCP_MAXITER) == 0
    u_first = u.copy()
    bs.solve()
    assert (host(bs.scp_iters)[conv] <= 2).all()
    # (the SCP stop test is |delta merit| < 1e-3, so a re-solve may still move u by the SCP tolerance)
    assert np.abs(host(bs.u)[conv] - u_first[conv]).max() < 1e-3
    # subset against the oracle (double precision, same stopping rule), free-running
    idx = np.arange(0, B, 16)
    S = oracle.mpc_setup(cb.x0[idx], cb.u0[idx], cb.veh[idx], cb.poly[idx], Hp=10, dt=0.4)
    O = oracle.scp_controller_batch(S["g"], S["cterm"], S["H"], S["qv"], S["gamma0"], cb.dsafe[idx], np.zeros((len(idx), 80)),
                                    opts=dict(abstol=1e-10, reltol=1e-10, feastol=1e-9), threads=os.cpu_count() or 1)
    same = (O["scp_iters"] == its[idx]) & (O["scp_iters"] <= 5)
    assert same.sum() >= len(idx) // 2
    assert np.abs(u_first[idx][same] - O["u"][same]).max() < 1e-5


def test_results_do_not_depend_on_batch_position(mods):
    """Sharding determinism: an instance's result is bit-identical wherever it sits in a batch."""
    scen = mods["scen"]
    cb = scen.circle_batch(96, step_lo=6, step_hi=7)
    def run(order):
        bs = mods["batch"].BatchSCP(len(order), 8, 10)
        bs.load_inputs(x0=cb.x0[order], u0=cb.u0[order], veh=cb.veh[order], poly=cb.poly[order], dsafe=cb.dsafe[order],
                       u=np.zeros((len(order), 80)))
        bs.controller_step()
        return host(bs.u), host(bs.scp_iters)
    u_a, it_a = run(np.arange(96))
    perm = np.random.default_rng(0).permutation(96)
    u_b, it_b = run(perm)
    np.testing.assert_array_equal(u_a[perm], u_b)
    np.testing.assert_array_equal(it_a[perm], it_b)
    u_c, _ = run(np.arange(48, 96))                          # a "second rank's" shard
    np.testing.assert_array_equal(u_a[48:], u_c)


def test_host_io_arena_round_trip(mods):
    """The host-buffer path (BatchSCP.host_io / upload / download: one contiguous copy each way over the I/O arena)
    gives bit-identical results to per-tensor load_inputs, and the pinned mirror holds every result of the step;
    an obstacle batch lays out its extra inputs in the same arena."""
    torch, scen = mods["torch"], mods["scen"]
    cb = scen.circle_batch(64, step_lo=6, step_hi=7)
    ref = mods["batch"].BatchSCP(64, 8, 10)
    ref.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((64, 80)))
    ref.controller_step()
    ref.advance_linear(scen.MECH_LIMIT, scen.DU_LIM)
    bs = mods["batch"].BatchSCP(64, 8, 10)
    hio = bs.host_io()
    for k in ("x0", "u0", "veh", "poly", "dsafe"):
        getattr(hio, k)[...] = np.asarray(getattr(cb, k)).reshape(getattr(hio, k).shape)
    hio.u[...] = 0.0
    bs.upload(hio)
    bs.controller_step()
    bs.advance_linear(scen.MECH_LIMIT, scen.DU_LIM)
    bs.download(hio)
    torch.cuda.synchronize()
    for k in ("x0", "u0", "u", "U", "traj", "obj", "max_violation", "scp_iters", "ipm_iters", "status"):
        np.testing.assert_array_equal(getattr(hio, k), host(getattr(ref, k)), err_msg=k)
        np.testing.assert_array_equal(getattr(hio, k), host(getattr(bs, k)), err_msg=k)
    assert hio.nbytes_in == sum(host(getattr(bs, k)).nbytes for k in ("veh", "poly", "dsafe", "x0", "u0", "u"))
    assert int(hio.scp_iters.sum()) > 64
    # second step straight from the host mirror (x0, u0, u are the values just read back)
    ref.controller_step(); bs.upload(hio); bs.controller_step(); bs.download(hio); torch.cuda.synchronize()
    np.testing.assert_array_equal(hio.u, host(ref.u))
    ob = mods["batch"].BatchSCP(3, 1, 10, nObst=22)
    assert ob.obst.shape == (3, 22, 10, 2) and ob.dsafe_obst.shape == (3, 1, 22)
    assert ob.obst.data_ptr() - ob.io.data_ptr() == ob.io_layout["obst"][0] and ob.obst.is_contiguous()


def test_work_order_and_ordered_solve(mods):
    """scpb200_work_order gives a permutation sorted by descending work; the ordered solve (longest instances first)
    returns bit-identical results to the natural pull order."""
    torch, scen = mods["torch"], mods["scen"]
    lib = mods["capi"].load()
    rng = np.random.default_rng(3)
    for B in (1, 7, 1024, 5000):
        work = torch.as_tensor(rng.integers(0, 3000, B).astype(np.int32)).cuda()
        order = torch.full((B,), -1, dtype=torch.int32, device="cuda")
        assert lib.scpb200_work_order(B, C.c_void_p(work.data_ptr()), C.c_void_p(order.data_ptr()), None) == 0
        o = host(order).astype(np.int64)
        assert sorted(o.tolist()) == list(range(B))
        w = np.minimum(host(work)[o], 2047)
        assert (np.diff(w) <= 0).all()
    cb = scen.circle_batch(300, step_lo=5, step_hi=7)
    def run(flag):
        bs = mods["batch"].BatchSCP(300, 8, 10)
        bs.schedule_by_previous_work = flag
        bs.load_inputs(x0=cb.x0, u0=cb.u0, veh=cb.veh, poly=cb.poly, dsafe=cb.dsafe, u=np.zeros((300, 80)))
        out = []
        for _ in range(2):
            bs.controller_step()
            out.append((host(bs.u).copy(), host(bs.scp_iters).copy(), host(bs.ipm_iters).copy()))
            bs.advance_linear(scen.MECH_LIMIT, scen.DU_LIM)
        return out
    for a, b in zip(run(True), run(False)):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)


def test_ode_predict_and_linear_advance(mods, oracle):
    G = load_golden("circle8_hp10_step10.npz")
    torch = mods["torch"]
    bs = make_batch(mods, G)
    T = float(G["sc_delay_x"] + G["sc_dt"] + G["sc_delay_u"])
    xm = torch.as_tensor(G["x_measured"][None], device=bs.device)
    ur = torch.as_tensor(G["u_path"][:, -1][None], device=bs.device)
    Y = host(bs.ode_predict(xm, ur, T, steps=10, nsub=16))[0]                   # [nVeh,10,6]
    assert np.abs(np.transpose(Y, (1, 2, 0)) - G["delay_traj"]).max() < 1e-7     # the reference's LSODA is ~1e-8
    O = np.array([oracle.ode_predict(G["x_measured"][v], G["u_path"][v, -1], G["sc_Lf"][v], G["sc_Lr"][v], T) for v in range(8)])
    assert np.abs(Y - O).max() < 1e-10
    assert np.abs(Y[:, -1] - G["x0"]).max() < 1e-7
    bs.setup()
    bs.load_inputs(u=G["u_final"][None])
    traj, U = bs.forward_u()
    bs.U.copy_(U)
    x_before, abe = host(bs.x0)[0].copy(), host(bs.abe)[0]
    bs.advance_linear(float(G["sc_mechanicalSteeringLimit"]), float(G["sc_duLim"]))
    for v in range(8):
        ua = np.clip(G["U"][0, v], -G["sc_mechanicalSteeringLimit"], G["sc_mechanicalSteeringLimit"])
        ua = np.clip(ua, G["u0"][v, 0] - G["sc_duLim"], G["u0"][v, 0] + G["sc_duLim"])
        xn = abe[v, :36].reshape(6, 6) @ x_before[v] + abe[v, 36:42] * ua + abe[v, 42:48]
        assert np.abs(host(bs.x0)[0, v] - xn).max() < 1e-12
        assert abs(host(bs.u0)[0, v] - ua) < 1e-15


def test_plant_step_and_device_resident_closed_loop(mods):
    """(a) scpb200_plant_step against the reference's own run (next measured state, clamped command);
    (b) the whole MPC step on device (delay compensation -> K1 -> K4 -> clamp + plant, no host round trip) run free
    for 50 steps from the scenario's initial state: collision-free, converges to the reference's end state, and equals
    the reference's trajectory up to the first symmetric conflict (where its own solver noise decides the branch)."""
    torch = mods["torch"]
    R = load_golden("circle8_hp10_run.npz")
    G0 = load_golden("circle8_hp10_step0.npz")
    mech, lat, du, dt = (float(R[k]) for k in ("sc_mechanicalSteeringLimit", "sc_lateralAccelerationLimit", "sc_duLim", "sc_dt"))
    n = R["x_measured"].shape[0] - 2            # the last transition is truncated at the end of the simulated timespan
    bs = make_batch(mods, G0, B=n)
    bs.U.copy_(torch.as_tensor(R["u_final"][:n].reshape(n, 8, 10).transpose(0, 2, 1).copy()))
    x = torch.as_tensor(R["x_measured"][:n].copy()).cuda()
    ua = torch.as_tensor(R["u0"][:n].copy()).cuda()
    umax, uc = bs.plant_step(x, ua, mech, lat, du, want_clamped=True)
    assert np.abs(host(x) - R["x_measured"][1:n + 1]).max() < 2e-7
    assert np.abs(host(ua) - R["u0"][1:n + 1]).max() < 1e-12
    assert np.abs(host(uc) - R["U_clamped"][:n]).max() < 1e-12
    # (b) free-running closed loop, 4 copies of the scenario in one batch (identical results expected)
    B = 4
    bs = make_batch(mods, G0, B=B)
    x = torch.as_tensor(np.repeat(R["sc_x_init"][None], B, 0).copy()).cuda()
    ua = torch.as_tensor(np.repeat(R["sc_u_init"][None], B, 0).copy()).cuda()
    delay = float(R["sc_delay_x"]) + dt + float(R["sc_delay_u"])
    path, feas = [host(x).copy()], []
    for i in range(50):
        bs.mpc_step(x, ua, mech, lat, du, delay)
        path.append(host(x).copy())
        feas.append((host(bs.status) & 8) == 0)
    path = np.stack(path)                                                       # [51, B, nVeh, 6]
    assert np.abs(path - path[:, :1]).max() == 0.0                              # copies agree bit for bit
    ref = R["vehiclePath_step_ends"].transpose(2, 1, 0)                         # [51, nVeh, 6]
    assert np.abs(path[:6, 0] - ref[:6]).max() < 1e-4                           # before the symmetric conflict: 1e-4 m
    assert np.all(feas)
    pos = path[:, 0, :, :2]
    dmin = min(np.linalg.norm(pos[:, a] - pos[:, b], axis=1).min() for a in range(8) for b in range(a + 1, 8))
    assert dmin > 3.0                                                           # reference run: 3.0661 m
    goal = R["sc_poly"][:, 1, :]
    assert np.linalg.norm(pos[-1] - goal, axis=1).max() < np.linalg.norm(ref[-1, :, :2] - goal, axis=1).max() + 1.0


def test_batch_rollout_in_the_reference_result_format(mods, tmp_path):
    """a batched device-resident rollout recorded under the keys / shapes main.py:213-224 dumps, against
    the reference's own run up to its first symmetric conflict (MPC step 6), and read back the way draw_video.py does."""
    import importlib
    results = importlib.import_module(PKG + ".results")
    R = load_golden("circle8_hp10_run.npz")
    G0 = load_golden("circle8_hp10_step0.npz")
    B, Nsim, nref = 2, 9, 6
    bs = make_batch(mods, G0, B=B)
    ro = results.BatchRollout(bs, np.repeat(R["sc_x_init"][None], B, 0), np.repeat(R["sc_u_init"][None], B, 0),
                              mech_limit=float(R["sc_mechanicalSteeringLimit"]), lat_acc_limit=float(R["sc_lateralAccelerationLimit"]),
                              duLim=float(R["sc_duLim"]), delay_u=float(R["sc_delay_u"]), tick_length=float(R["sc_tick_length"]),
                              Nsim=Nsim)
    assert (ro.tps, ro.tdu, ro.ticks_total) == (int(R["sc_ticks_per_sim"]), int(R["sc_ticks_delay_u"]), Nsim * 40)
    ro.run()
    A, A1 = ro.result_arrays(0), ro.result_arrays(1)
    for k in results.RESULT_KEYS[:9]:
        np.testing.assert_array_equal(A[k], A1[k])                              # identical scenarios, identical records
    nt = nref * ro.tps
    assert np.abs(A["vehiclePathFullRes"][:, :, :nt + 1:10] - R["vehiclePath_every10"][:, :, :nt // 10 + 1]).max() < 1e-4
    assert np.abs(A["controlPathFullRes"][:, :nt + 1:10] - R["controlPath_every10"][:, :nt // 10 + 1]).max() < 1e-6
    assert np.abs(A["trajectoryPredictions"][..., :nref] - np.moveaxis(R["Traj"][:nref], 0, -1)).max() < 1e-4
    assert np.abs(A["controlPredictions"][..., :nref] - np.moveaxis(R["U_clamped"][:nref], 0, -1)).max() < 1e-5
    assert np.abs(A["ReferenceTrajectory"][..., :nref] - np.moveaxis(R["RefPts"][:nref], 0, -1)).max() < 1e-4
    assert np.abs(A["initial_pos"][..., :nref] - np.moveaxis(R["x0"][:nref, :, :2], 0, -1).transpose(1, 0, 2)).max() < 1e-4
    np.testing.assert_allclose(A["evaluations_obj_value"][:nref], R["evaluations_obj_value"][:nref], rtol=1e-5, atol=1e-9)
    assert (ro.scp_iters[0, :nref] == R["scp_iters"][:nref]).all()
    # the actuator path is piecewise constant: step i's first clamped command from tick (i+1)*40 + 4 on
    cp = A["controlPathFullRes"]
    for i in range(Nsim - 2):
        seg = cp[:, (i + 1) * 40 + 4:(i + 2) * 40 + 4]
        assert (seg == A["controlPredictions"][0, :, i][:, None]).all()
    assert not np.isnan(A["vehiclePathFullRes"]).any() and np.isnan(cp).sum() == 0
    # file round trip in the reader's convention (draw_video.py:42-56)
    path = str(tmp_path / "Circle_num_8_control_SCP.json")
    ro.dump_json(0, path)
    with open(path) as f:
        import json
        assert tuple(json.load(f).keys()) == results.RESULT_KEYS
    L = results.load_result(path, nx=6, nVeh=8, nObst=0, Hp=10, Nsim=Nsim, ticks_total=ro.ticks_total)
    for k in ("vehiclePathFullRes", "controlPathFullRes", "controlPredictions", "trajectoryPredictions", "ReferenceTrajectory",
              "MPC_delay_compensation_trajectory"):
        np.testing.assert_array_equal(L[k], A[k])
    assert L["initial_pos"].shape == (1, 2, 8, Nsim) and L["evaluations_obj_value"].shape == (Nsim, 1)


def test_noise_is_keyed_per_instance_and_reproducible(mods, oracle):
    G = load_golden("circle8_hp10_step10.npz")
    bs = make_batch(mods, G, B=4, noise_sigma=3e-6, seed=99, instance0=5, noise_counter=2)
    bs.setup()
    x0, u0, veh, poly = (np.repeat(a, 4, axis=0) for a in golden_setup_inputs(G))
    O = oracle.mpc_setup(x0, u0, veh, poly, Hp=10, dt=0.4, noise_sigma=3e-6, seed=99, instance0=5, noise_counter=2)
    assert np.abs(host(bs.cterm) - O["cterm"]).max() <= 1e-12 * np.abs(O["cterm"]).max()
    assert np.abs(host(bs.cterm)[0] - host(bs.cterm)[1]).max() > 1e-9


def test_error_codes_not_exceptions_from_c(mods):
    capi = mods["capi"]
    lib = capi.load()
    d = capi.Dims(1, 0, 10, 0, 2)
    n = C.c_size_t(0)
    assert lib.scpb200_workspace_bytes(C.byref(d), C.byref(n)) == -1
    assert b"bad dims" in lib.scpb200_last_error()
    p = capi.Params()
    lib.scpb200_default_params(C.byref(p))
    d = capi.Dims(1, 8, 10, 0, 2)
    assert lib.scpb200_mpc_setup(C.byref(d), C.byref(p), *([None] * 13)) == -1


class _Model:
    nx, nu, ny, is_noise = 6, 1, 2, False


class _Scenario:
    """The fields of the reference's Scenario object that the path reads, from a golden record."""

    def __init__(self, G):
        self.model = _Model()
        self.nVeh, self.Hp, self.Hu, self.nObst = int(G["sc_nVeh"]), int(G["sc_Hp"]), int(G["sc_Hp"]), 0
        self.dt, self.dsafeExtra = float(G["sc_dt"]), float(G["sc_dsafeExtra"])
        self.tick_length, self.delay_x, self.delay_u = float(G["sc_tick_length"]), float(G["sc_delay_x"]), float(G["sc_delay_u"])
        self.mechanicalSteeringLimit, self.duLim = float(G["sc_mechanicalSteeringLimit"]), float(G["sc_duLim"])
        self.Lf, self.Lr, self.Q, self.Q_final, self.R = (list(G[k]) for k in ("sc_Lf", "sc_Lr", "sc_Q", "sc_Q_final", "sc_R"))
        self.referenceTrajectories = [p for p in G["sc_poly"]]
        self.dsafeVehicles = G["sc_dsafeVehicles"]
        self.obstacles, self.dsafeObstacles = [], np.zeros((self.nVeh, 0))


def test_reference_call_surface_closed_loop(mods):
    """main.py:123-134 with the drop-in classes, fed the reference's own measured states for all 50 MPC steps of
    the default run: IterClass (delay compensation + reference sampling), SCPcontroller(...).SCP_controller(Iter),
    warm-started from the previous controllerOutput exactly as main.py does."""
    facade_iter = importlib.import_module(PKG + ".MPC_Iter")
    facade_scp = importlib.import_module(PKG + ".SCP_controller")
    R = load_golden("circle8_hp10_run.npz")
    sc = _Scenario(load_golden("circle8_hp10_step0.npz"))
    outputs = []
    nsteps = R["x_measured"].shape[0]
    for i in range(nsteps):
        Iter = facade_iter.IterClass(sc, R["x_measured"][i], R["u_path"][i], np.zeros((0, 2)), np.full((1, sc.nVeh), sc.mechanicalSteeringLimit))
        assert np.abs(Iter.x0 - R["x0"][i]).max() < 1e-7                        # the reference's LSODA is ~1e-8
        assert np.abs(Iter.ReferenceTrajectoryPoints - R["RefPts"][i]).max() < 1e-6
        # teacher-forced warm start: the reference's previous solution (the SCP map amplifies 1e-14 at symmetric steps)
        prev = [] if i == 0 else {"u": R["u_final"][i - 1].reshape(-1, 1)}
        ctl = facade_scp.SCPcontroller(sc, Iter, prev)
        U, traj, out = ctl.SCP_controller(Iter)
        outputs.append(out)
        assert U.shape == (sc.Hp, sc.nVeh) and traj.shape == (sc.Hp, 2, sc.nVeh) and out["u"].shape == (sc.nVeh * sc.Hp, 1)
        if R["scp_iters"][i] <= 5:
            assert len(out["optimization_log"]["slack"]) == R["scp_iters"][i]
            assert np.abs(out["u"].ravel() - R["u_final"][i]).max() < 2e-6
            assert np.abs(traj - R["Traj"][i]).max() < 1e-4
        feas = ctl.QCQP_evaluate(out["u"])[0]
        assert feas == bool(R["feasible_last"][i])
        evo = ctl.evaluateInOriginalProblem(U, traj, {"ignoreQCQPcheck": True})
        if R["scp_iters"][i] <= 5:
            # evaluations_obj_value is computed by the reference on the CLAMPED U (main.py:164-174, :202)
            evc = ctl.evaluateInOriginalProblem(R["U_clamped"][i], R["Traj"][i], {})
            assert abs(evc["predictionObjectiveValue"] - R["evaluations_obj_value"][i]) <= 1e-6 * max(1.0, abs(R["evaluations_obj_value"][i]))
        assert "constraintValuesVehicle" in evo
    assert sum(len(o["optimization_log"]["slack"]) for o in outputs) in range(120, 129)      # 124 QPs in the reference run


def test_mpcclass_attributes_match_reference(mods):
    facade_iter = importlib.import_module(PKG + ".MPC_Iter")
    G = load_golden("circle8_hp10_step10.npz")
    sc = _Scenario(G)

    class _It:
        pass
    It = _It()
    It.x0, It.u0 = G["x0"], G["u0"]
    mpc = facade_iter.MPCclass(sc, It)
    for name, ref in (("Mathcal_A", G["Mathcal_A"]), ("Mathcal_B", G["Mathcal_B"]), ("Phi_0", G["Phi_0"])):
        got = getattr(mpc, name)
        assert got.shape == ref.shape, name
        assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max(), name
    assert np.abs(mpc.Mathcal_C[:, 0, :] - G["Mathcal_C"]).max() <= 1e-11 * np.abs(G["Mathcal_C"]).max()
    assert np.abs(mpc.const_term[:, 0, :] - G["const_term"]).max() <= 1e-11 * np.abs(G["const_term"]).max()
    assert np.abs(mpc.gamma_0[0] - G["gamma_0"]).max() <= 1e-9 * max(1.0, np.abs(G["gamma_0"]).max())
    assert np.abs(mpc.A[:, :, 0, :] - G["A"]).max() < 1e-12 and mpc.A.shape == (6, 6, sc.Hp, sc.nVeh)
    assert np.abs(mpc.B[:, 0, 3, :] - G["B"]).max() < 1e-12 and np.abs(mpc.E[:, 5, :] - G["E"]).max() < 1e-12
