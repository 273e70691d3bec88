"""Name kept for import compatibility with sopa/src/solvers/rk_parametric_order2stage2.py; the class lives in rk_parametric.py."""
from .rk_parametric import RKParametricSolver, RKOrder2Stage2  # noqa: F401
