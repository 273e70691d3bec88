"""Name kept for import compatibility with sopa/src/solvers/rk_parametric_order3stage3.py; the class lives in rk_parametric.py."""
from .rk_parametric import RKParametricSolver, RKOrder3Stage3  # noqa: F401
