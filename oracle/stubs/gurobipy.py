"""Empty stand-in for MIQP.py:10-11 (`import gurobipy`, `from gurobipy import GRB`); MIQP is out of scope."""
class GRB:  # noqa: D401
    pass
