"""Name kept for import compatibility with sopa/src/solvers/euler.py; the class lives in rk_parametric.py."""
from .rk_parametric import RKParametricSolver, Euler  # noqa: F401
