"""CPU oracle for the meta-solver ODE-block hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU with plain torch/numpy ops, the algorithm of the
reference's fixed-step parametrized Runge-Kutta ODE block (juliagusak/neural-ode-metasolver,
`sopa/src/solvers/*`, `sopa/src/models/odenet_{cifar10,mnist}/layers.py`).  Every function
cites the reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import it -- as the checker or the reported baseline, never as the product path.
The product (`neural-ode-metasolver_b200/`) never imports `oracle`.

Parity pin: the oracle is checked bit-for-bit (fp32) against golden vectors produced by
running the *real* reference from /root/reference (tests/golden/make_golden.py; the
reference is Python and cannot travel to the GPU box, so its outputs are committed as
fixtures).  The arithmetic itself lives in PyTorch (pinned by the reference only in prose,
README.md:9-11 "pytorch==1.7"; this image has torch 2.11) -- conv2d / gelu / group_norm are
called here exactly as the reference calls them.
"""
from .detrand import det_uniform, det_normal  # noqa: F401
from .tableau import butcher_tableau  # noqa: F401
from .rk import (make_time_grid, integrate, ode_block_forward, rhs_preact, rhs_preact_gn, rhs_postact, rhs_postact_gn,  # noqa: F401
                 rhs_mnist, RhsCounter)
