"""metasolver_b200: B200-native fixed-step parametrized Runge-Kutta neural-ODE block.

Drop-in for the hot path of juliagusak/neural-ode-metasolver behind the reference's own `sopa`
API: `metasolver_b200.sopa.src.solvers.utils.create_solver`, solver objects carrying u / v and the
step count, `MetaODEBlock`, the standalone / switching / solver-ensembling regimes.  All device
arithmetic of an ODE block runs in hand-written sm_100a CUDA kernels reached through the C ABI in
include/metasolver_b200.h (ctypes binding: _cabi.py).  There is no cuDNN, Triton or CPU path.
"""
from . import _cabi  # noqa: F401
from .ops import (ode_block_integrate, input_grad_only, set_default_engine, launch_count,  # noqa: F401
                  profile_enable, profile_read, profile_read_executed, set_option, get_option, pool_fc, cross_entropy,
                  set_library_fallback, library_fallback_allowed)
from .graphs import GraphedStep, HostFedLoop  # noqa: F401
from .train_ops import attack_step, FusedSGD, CyclicLR, augment_normalize  # noqa: F401

__version__ = "0.1.0"
