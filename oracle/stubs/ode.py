"""Empty stand-in for the unused `import ode` at MPC_Iter.py:4 (test infrastructure only)."""
