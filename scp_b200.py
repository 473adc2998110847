"""Importable alias of the package directory `senquential-convex-programming-for-trajectory-planning_b200/`
(the driver-mandated name contains '-' and cannot appear in an `import` statement).

    from scp_b200.SCP_controller import SCPcontroller          # the reference's call surface (main.py:13)
    from scp_b200.MPC_Iter import IterClass, MPCclass           # (main.py:15)
    import scp_b200
    scp_b200.batch.BatchSCP(...)                                # the batched controller stage on a B200

`scp_b200.<sub>` IS the module `<package>.<sub>` (one module object under two names), so module-level state such as
the loaded shared library or the engine cache is never duplicated.
"""
import importlib
import importlib.abc
import importlib.machinery
import sys

_PKG = "senquential-convex-programming-for-trajectory-planning_b200"
_pkg = importlib.import_module(_PKG)

# a package without files of its own: sub-modules are resolved by the finder below
__path__ = []


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """`scp_b200.X` -> the already imported (or now imported) module `<package>.X`."""

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(__name__ + "."):
            return None
        try:
            real = importlib.import_module(_PKG + fullname[len(__name__):])
        except ModuleNotFoundError as e:
            if e.name and e.name.startswith(_PKG):
                return None                                   # no such sub-module: the import system raises
            raise
        spec = importlib.machinery.ModuleSpec(fullname, self, is_package=hasattr(real, "__path__"))
        spec._scp_real = real
        return spec

    def create_module(self, spec):
        real = spec._scp_real
        self._real_spec = real.__spec__
        return real

    def exec_module(self, module):
        module.__spec__ = self._real_spec                     # the import system has just overwritten it with the alias spec


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    try:
        return importlib.import_module(f"{__name__}.{name}")
    except ModuleNotFoundError:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}") from None


__version__ = _pkg.__version__
