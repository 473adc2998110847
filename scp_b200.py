"""Importable alias of the package directory `senquential-convex-programming-for-trajectory-planning_b200/`
(the driver-mandated name contains '-' and cannot appear in an `import` statement).

    import scp_b200
    scp_b200.batch.BatchSCP(...)        # the batched controller stage on a B200
    scp_b200.SCP_controller.SCPcontroller / scp_b200.MPC_Iter.IterClass, MPCclass   # the reference's call surface
"""
import importlib
import sys

_PKG = "senquential-convex-programming-for-trajectory-planning_b200"
_pkg = importlib.import_module(_PKG)


def __getattr__(name):
    mod = importlib.import_module(f"{_PKG}.{name}")
    setattr(sys.modules[__name__], name, mod)
    return mod


__version__ = _pkg.__version__
