"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/scpb200.h declares,
agrees with the Python mirror of its structs, and fails loudly (no fallback) without a CUDA device."""
import ctypes as C
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

PKG = "senquential-convex-programming-for-trajectory-planning_b200"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scpb200.h")


@pytest.fixture(scope="module")
def capi():
    build = importlib.import_module(PKG + ".build")
    build.build()                                   # nvcc cross-compiles sm_100a without a GPU
    return importlib.import_module(PKG + "._capi")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scpb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(capi):
    lib = capi.load()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/scpb200.h but not exported"
    assert set(syms) == set(capi.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert lib.scpb200_version() == 100


def test_library_is_sm100a_only_and_has_no_torch_dependency(capi):
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)
    ldd = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd


def test_struct_layout_and_defaults_match_c(capi):
    lib = capi.load()
    p = capi.Params()
    lib.scpb200_default_params(C.byref(p))
    q = capi.default_params_py()
    for name, _ in capi.Params._fields_:
        assert getattr(p, name) == getattr(q, name), name
    assert abs(p.uLim - 0.05235987755982989) < 1e-18
    assert p.constraint_tol == 2 * 2.1 * 1e-3                    # Config.py:18
    assert C.sizeof(capi.Dims) == 20


def test_no_cpu_fallback_without_a_device(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = capi.load()
    assert lib.scpb200_device_count() == 0
    d = capi.Dims(1, 8, 10, 0, 2)
    n = C.c_size_t(0)
    rc = lib.scpb200_workspace_bytes(C.byref(d), C.byref(n))
    assert rc == -2 and lib.scpb200_last_error()                 # SCPB200_ERR_CUDA, with a message
    batch = importlib.import_module(PKG + ".batch")
    with pytest.raises(capi.Scpb200Error):
        batch.BatchSCP(1, 8, 10)
    with pytest.raises(capi.Scpb200Error):
        batch.qp_solve_dense(*[torch.zeros(1, 2, 2, dtype=torch.float64)] * 6)


def test_argument_errors_are_codes(capi):
    lib = capi.load()
    n = C.c_size_t(0)
    assert lib.scpb200_workspace_bytes(C.byref(capi.Dims(1, 0, 10, 0, 2)), C.byref(n)) == -1
    assert b"bad dims" in lib.scpb200_last_error()
    assert lib.scpb200_workspace_bytes(None, C.byref(n)) == -1


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package (or bench.py's product arm) may
    import, link or execute oracle/ or tests/emu."""
    pkg_dir = os.path.join(ROOT, PKG)
    banned = [r"import\s+oracle", r"from\s+oracle", r"oracle[/.]", r"libscp_oracle", r"\borc_\w+", r"libscpb200_emu",
              r"from\s+emu", r"import\s+emu", r"emu\.cpp", r"\bemu_\w+\s*\("]
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                for pat in banned:
                    assert not re.search(pat, txt), f"{f} matches {pat}"


def test_scenario_generator_reproduces_reference_scenario():
    scen = importlib.import_module(PKG + ".scenarios")
    G = dict(np.load(os.path.join(ROOT, "tests", "golden", "circle8_hp10_step0.npz")))
    r = scen.reference_circle()
    assert np.abs(r.x0[0] - G["sc_x_init"]).max() == 0.0
    assert np.abs(r.poly[0] - G["sc_poly"]).max() == 0.0
    assert np.abs(r.dsafe[0] - G["sc_dsafeVehicles"]).max() < 1e-15
    a = scen.circle_batch(4, instance0=10)
    b = scen.circle_batch(8, instance0=8)
    np.testing.assert_array_equal(a.x0, b.x0[2:6])               # an instance does not depend on the sharding
