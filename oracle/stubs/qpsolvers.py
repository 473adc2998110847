"""Empty stand-in for the unused `from qpsolvers import solve_qp` at SCP_controller.py:5."""
def solve_qp(*a, **k):
    raise NotImplementedError("qpsolvers stub: the reference never calls this")
