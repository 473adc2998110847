import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc


def golden_setup_inputs(G):
    """(x0[1,nVeh,6], u0[1,nVeh], veh[1,nVeh,5], poly[1,nVeh,nPts,2]) of one golden step record."""
    nVeh = int(G["sc_nVeh"])
    veh = np.stack([G["sc_Lf"], G["sc_Lr"], G["sc_Q"], G["sc_Q_final"], G["sc_R"]], axis=1)
    return G["x0"][None], G["u0"].reshape(1, nVeh), veh[None], G["sc_poly"][None]
