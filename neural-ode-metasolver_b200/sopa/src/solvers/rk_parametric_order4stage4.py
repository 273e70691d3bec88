"""Name kept for import compatibility with sopa/src/solvers/rk_parametric_order4stage4.py; the class lives in rk_parametric.py."""
from .rk_parametric import RKParametricSolver, RKOrder4Stage4  # noqa: F401
