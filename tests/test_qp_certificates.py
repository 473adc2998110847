"""Independent certificates for the QP solve — CPU suite.

The reference delegates the solve (SCP_controller.py:135-150) to a third-party solver that cannot be installed here,
and ships no solver fixtures, so the oracle's coneqp restatement (oracle/scp_oracle.c) is certified from OUTSIDE:

  * every golden QP solution (x, multipliers — stored by oracle/make_golden.py) satisfies the KKT conditions of the QP
    the REFERENCE posed, evaluated in NumPy longdouble with data rebuilt in NumPy from the reference's own MPCclass
    arrays; the residuals bound the distance to the exact minimiser (P is strongly convex on u);
  * closed-form known-answer QPs (box-only, no rows, single active row, omega-active) through the oracle and through
    the kernels' interior-point source (the emulator build; the real kernels run them in tests/test_gpu_workloads.py);
  * Philox4x32-10 against Random123's known-answer vectors.
"""
import glob
import importlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from qp_cases import (KAT_CASES, PHILOX_KAT, kkt_certificate, minimiser_distance_bound, noise_pair_py, philox4x32_10_py,
                      qp_from_reference_arrays)

capi = importlib.import_module("senquential-convex-programming-for-trajectory-planning_b200._capi")
STEP_FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "circle*_step*.npz")))


# ------------------------------------------------------------------------------------------------ golden QPs
@pytest.mark.parametrize("fname", STEP_FILES)
def test_numpy_qp_data_equals_what_the_reference_logged(fname):
    """The NumPy restatement of the QP data (used by the certificate below for the iterations whose dense matrices are
    not stored) against the dense P, q, Aineq, bineq, lb, ub the reference itself logged."""
    G = load_golden(fname)
    for it in sorted(int(k.split("_")[1]) for k in G if k.startswith("Aineq_")):
        P, q, A, b, lb, ub = qp_from_reference_arrays(G, G["prev_u"][it])
        assert np.abs(P - G[f"P_{it}"]).max() <= 1e-12 * np.abs(P).max()
        assert np.abs(A - G[f"Aineq_{it}"]).max() <= 1e-11 * np.abs(A).max()
        assert (A[G[f"Aineq_{it}"] == 0] == 0).all()
        assert np.abs(b - G[f"bineq_{it}"]).max() <= 1e-11 * np.abs(b).max()
        assert np.abs(q - G[f"q_{it}"]).max() <= 1e-12 * np.abs(q).max()
        np.testing.assert_array_equal(lb, G[f"lb_{it}"])
        np.testing.assert_array_equal(ub, G[f"ub_{it}"])


@pytest.mark.parametrize("fname", STEP_FILES)
def test_every_golden_qp_solution_is_a_kkt_point_in_longdouble(fname):
    """Stationarity, primal feasibility, dual sign and complementarity of every golden (x, z) in 80-bit arithmetic,
    independent of scp_oracle.c; tolerance 1e-10 relative to the natural scale of each residual (|q| for the gradient:
    the slack weight alone is 1e5; max(1, |f|) for the gap), and the distance to the exact minimiser they imply."""
    G = load_golden(fname)
    assert "zA" in G, "golden fixture without multipliers: regenerate with oracle/make_golden.py"
    nVeh, Hp = int(G["sc_nVeh"]), int(G["sc_Hp"])
    worst = dict(stat=0.0, primal=0.0, dual=0.0, gap=0.0, bound=0.0)
    for it in range(int(G["scp_iters"])):
        if f"Aineq_{it}" in G:
            P, q, A, b, lb, ub = (G[f"{k}_{it}"] for k in ("P", "q", "Aineq", "bineq", "lb", "ub"))
        else:
            P, q, A, b, lb, ub = qp_from_reference_arrays(G, G["prev_u"][it])
        x = G["x"][it]
        cert = kkt_certificate(P, q, A, b, lb, ub, x, G["zA"][it], G["zub"][it], G["zlb"][it])
        fval = abs(float(0.5 * x @ P @ x + q @ x))
        qs = max(1.0, float(np.abs(q).max()))
        assert cert["stationarity"] <= 1e-10 * qs, (it, cert)
        assert cert["primal"] <= 1e-10, (it, cert)
        assert cert["dual"] == 0.0, (it, cert)
        assert cert["gap"] <= 1e-10 * max(1.0, fval), (it, cert)
        lam_min = 2.0 * min(np.linalg.eigvalsh(G["Phi_0"][:, :, v])[0] for v in range(nVeh))
        radius = 2.0 * float(G["sc_uLim"]) * np.sqrt(nVeh * Hp)
        amax = float(np.sqrt((np.asarray(A)[:, :-1] ** 2).sum(axis=1)).max())
        bound = minimiser_distance_bound(cert, lam_min, radius, amax, float(q[-1]), A.shape[0] + 2 * len(q))
        worst = dict(stat=max(worst["stat"], cert["stationarity"] / qs), primal=max(worst["primal"], cert["primal"]),
                     dual=max(worst["dual"], cert["dual"]), gap=max(worst["gap"], cert["gap"] / max(1.0, fval)),
                     bound=max(worst["bound"], bound))
        assert bound <= 5e-6, (it, bound, cert)                   # inside the north-star's 1e-5 on u (a rigorous, not a tight, bound)
    print(f"\n[{fname}] worst over {int(G['scp_iters'])} QPs: stationarity/|q| {worst['stat']:.1e}, primal {worst['primal']:.1e}, "
          f"gap/|f| {worst['gap']:.1e}, certified |u - u*|_2 <= {worst['bound']:.1e}")


# ------------------------------------------------------------------------------------------------ known-answer QPs
@pytest.mark.parametrize("case", KAT_CASES, ids=lambda f: f.__name__)
def test_known_answer_qps_oracle(oracle, case):
    K = case()
    for opts in (dict(abstol=1e-10, reltol=1e-10, feastol=1e-10, maxiters=100), None):
        r = oracle.qp_boxed(K["P"], K["q"], K["A"], K["b"], K["lb"], K["ub"], opts=opts)
        assert r["status"] == 0, K["name"]
        assert np.abs(r["x"] - K["x"]).max() < 1e-7 * max(1.0, np.abs(K["x"]).max()), (K["name"], np.abs(r["x"] - K["x"]).max())
        assert abs(r["fval"] - K["fval"]) <= 1e-8 * max(1.0, abs(K["fval"]))
        if "zA" in K:
            assert np.abs(r["zA"] - K["zA"]).max() < 1e-6
        cert = kkt_certificate(K["P"], K["q"], K["A"], K["b"], K["lb"], K["ub"], r["x"], r["zA"], r["zub"], r["zlb"])
        assert cert["stationarity"] <= 1e-7 * max(1.0, np.abs(K["q"]).max()) and cert["primal"] <= 1e-9 and cert["dual"] == 0.0


@pytest.mark.parametrize("case", KAT_CASES, ids=lambda f: f.__name__)
@pytest.mark.parametrize("nt", [32, 128])
def test_known_answer_qps_kernel_source(case, nt):
    """The kernels' interior-point source (ipm_core.cuh compiled for the host by tests/emu) on the closed-form QPs."""
    from emu import emu
    K = case()
    emu.config(nt=nt, reverse=False)
    r = emu.qp_solve_dense(K["P"][None], K["q"][None], K["A"][None], K["b"][None], K["lb"][None], K["ub"][None],
                           capi.default_params_py())
    assert (int(r["status"][0]) & ~capi.ST_QP_DRES_FLOOR) == 0, (K["name"], r["status"], r["iters"])
    assert np.abs(r["x"][0] - K["x"]).max() < 1e-7 * max(1.0, np.abs(K["x"]).max()), (K["name"], np.abs(r["x"][0] - K["x"]).max())
    assert abs(r["fval"][0] - K["fval"]) <= 1e-8 * max(1.0, abs(K["fval"]))
    if "zA" in K:
        assert np.abs(r["zA"][0] - K["zA"]).max() < 1e-6


# ------------------------------------------------------------------------------------------------ Philox
def test_philox4x32_10_known_answer_vectors(oracle):
    """Random123's kat_vectors for philox4x32-10: the plain-Python restatement of the published algorithm, the oracle's C
    and (through the noise pairs) the kernels' source all produce them."""
    for ctr, key, out in PHILOX_KAT:
        assert philox4x32_10_py(ctr, key) == out
        assert tuple(int(v) for v in oracle.philox4x32_10(ctr, key)) == out


def test_noise_pairs_host_implementations_agree(oracle):
    for args in [(0, 0, 0, 0), (12345, 7, 3, 9), (0xFEDCBA9876543210, 1 << 31, 7, 0xFFFFFFFF)]:
        a, b = oracle.noise_pair(*args), noise_pair_py(*args)
        assert abs(a[0] - b[0]) < 1e-15 and abs(a[1] - b[1]) < 1e-15
    assert noise_pair_py(1, 2, 3, 4, stream=0) != noise_pair_py(1, 2, 3, 4, stream=1)


def test_documented_drop_in_imports_work_verbatim():
    """the import lines INTEGRATION.md documents, executed as written (in a fresh interpreter)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("from scp_b200.SCP_controller import SCPcontroller\n"
            "from scp_b200.MPC_Iter import IterClass, MPCclass\n"
            "import scp_b200, importlib, sys\n"
            "real = importlib.import_module('senquential-convex-programming-for-trajectory-planning_b200.MPC_Iter')\n"
            "assert sys.modules['scp_b200.MPC_Iter'] is real and real.IterClass is IterClass\n"
            "assert scp_b200.batch.BatchSCP and scp_b200.SCP_controller.SCPcontroller is SCPcontroller\n"
            "try:\n    scp_b200.no_such_module\n    raise SystemExit(1)\nexcept AttributeError:\n    pass\n"
            "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr
