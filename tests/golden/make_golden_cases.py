"""Case tables and deterministic inputs shared by make_golden.py (reference side) and the tests."""
import numpy as np

from oracle.detrand import det_uniform, det_normal

WG_STRIDE = 17

# ----------------------------------------------------------------------------- case tables
TABLEAU_CASES = [
    # method, parameterization, u0, v0
    ("euler", None, -1, -1),
    ("rk2", "u", 0.5, -1), ("rk2", "u", 1.0, -1), ("rk2", "u", 0.3, -1), ("rk2", "u", 2 / 3., -1),
    ("rk2", "u", 0.05, -1), ("rk2", "u", 1.5, -1), ("rk2", "u", 1e-5, -1), ("rk2", "u", 0.5125, -1),
    ("rk3", "uv", 1 / 3., 2 / 3.), ("rk3", "uv", 0.5, 1.0), ("rk3", "uv", 0.4, 0.4), ("rk3", "uv", 1.0, 1.0),
    ("rk3", "uv", 0.25, 0.8),
    ("rk4", "u1", 0.2, -1), ("rk4", "u2", 1 / 3., -1), ("rk4", "u3", 0.1, -1), ("rk4", "u2", 0.45, -1),
    ("rk4", "uv", 1 / 3., 2 / 3.), ("rk4", "uv", 0.5, 0.7), ("rk4", "uv", 0.3, 0.3), ("rk4", "uv", 0.6, 0.9),
    ("rk4", "u1", 1.2, -1),
]

# name, C, H, W, B, rhs kind, solver tuple (method, param, n_steps, step_size, u0, v0)
ODE_CASES = [
    ("c64_rk2_u05_n8", 64, 8, 32, 2, "preact", ("rk2", "u", 8, -1, 0.5, -1)),
    ("c64_rk2_u03_n3", 64, 8, 32, 2, "preact", ("rk2", "u", 3, -1, 0.3, -1)),
    ("c64_rk2_u1_n10", 64, 4, 32, 1, "preact", ("rk2", "u", 10, -1, 1.0, -1)),
    ("c64_euler_n2", 64, 8, 32, 2, "preact", ("euler", None, 2, -1, -1, -1)),
    ("c64_rk3_n2", 64, 8, 32, 2, "preact", ("rk3", "uv", 2, -1, 1 / 3., 2 / 3.)),
    ("c64_rk4u2_n2", 64, 8, 32, 2, "preact", ("rk4", "u2", 2, -1, 1 / 3., -1)),
    ("c64_rk4uv_n1", 64, 8, 32, 2, "preact", ("rk4", "uv", 1, -1, 1 / 3., 2 / 3.)),
    ("c64_rk4u1_n1", 64, 4, 32, 1, "preact", ("rk4", "u1", 1, -1, 0.2, -1)),
    ("c64_rk4u3_n1", 64, 4, 32, 1, "preact", ("rk4", "u3", 1, -1, 0.1, -1)),
    ("c64_rk2_step03", 64, 4, 32, 1, "preact", ("rk2", "u", -1, 0.3, 0.5, -1)),
    ("c128_rk2_u05_n8", 128, 8, 16, 2, "preact", ("rk2", "u", 8, -1, 0.5, -1)),
    ("c128_rk4u2_n1", 128, 16, 16, 1, "preact", ("rk4", "u2", 1, -1, 1 / 3., -1)),
    ("c16_rk2_u05_n2_odd", 16, 5, 7, 3, "preact", ("rk2", "u", 2, -1, 0.5, -1)),
    ("c64_post_rk2_n2", 64, 8, 32, 2, "postact", ("rk2", "u", 2, -1, 0.5, -1)),
    ("c128_post_rk2_n8", 128, 8, 16, 2, "postact", ("rk2", "u", 8, -1, 0.5, -1)),
    ("c64_post_rk4u2_n2", 64, 4, 32, 1, "postact", ("rk4", "u2", 2, -1, 1 / 3., -1)),
]


def conv_w(c_out, c_in, seed, k=3):
    bound = 1.0 / np.sqrt(c_in * k * k)          # nn.Conv2d default init range
    return det_uniform((c_out, c_in, k, k), seed, -bound, bound)


def ode_case_inputs(C, H, W, B, seed=0):
    x = det_normal((B, C, H, W), 11 + seed)
    w1 = conv_w(C, C, 21 + seed)
    w2 = conv_w(C, C, 31 + seed)
    r = det_normal((B, C, H, W), 41 + seed)
    return x, w1, w2, r



REGIME_SOLVERS = [("rk2", "u", 4, -1, 0.3, -1), ("rk2", "u", 4, -1, 0.5, -1), ("rk2", "u", 2, -1, 2 / 3., -1),
                  ("rk4", "u2", 2, -1, 1 / 3., -1)]

# BASELINE config 3: solver ensembling / model ensembling over 4 RK2 u values (and RK4), 8 steps each
C3_RK2_SOLVERS = [("rk2", "u", 8, -1, 0.3, -1), ("rk2", "u", 8, -1, 0.5, -1), ("rk2", "u", 8, -1, 2 / 3., -1),
                  ("rk2", "u", 8, -1, 1.0, -1)]
C3_RK4_SOLVERS = [("rk4", "u2", 8, -1, 1 / 3., -1), ("rk4", "uv", 8, -1, 1 / 3., 2 / 3.)]
C3_WEIGHTS = [0.4, 0.3, 0.2, 0.1]

# gradients w.r.t. the solver parameters u (and v): name, C, H, W, B, rhs kind, solver tuple
SOLVER_GRAD_CASES = [
    ("sg_rk2_u05_n4", 64, 8, 32, 2, "preact", ("rk2", "u", 4, -1, 0.5, -1)),
    ("sg_rk2_u03_n3", 64, 8, 32, 2, "preact", ("rk2", "u", 3, -1, 0.3, -1)),
    ("sg_rk3_n2", 64, 8, 32, 2, "preact", ("rk3", "uv", 2, -1, 0.3, 0.7)),
    ("sg_rk4u2_n2", 64, 8, 32, 2, "preact", ("rk4", "u2", 2, -1, 0.3, -1)),
    ("sg_rk4uv_n1", 64, 8, 32, 2, "preact", ("rk4", "uv", 1, -1, 0.3, 0.7)),
    ("sg_rk4u1_n1", 64, 4, 32, 1, "preact", ("rk4", "u1", 1, -1, 0.2, -1)),
    ("sg_rk4u3_n1", 64, 4, 32, 1, "preact", ("rk4", "u3", 1, -1, 0.1, -1)),
    ("sg_c128_rk2_n2", 128, 8, 16, 2, "preact", ("rk2", "u", 2, -1, 0.6, -1)),
    ("sg_post_rk2_n2", 64, 8, 32, 2, "postact", ("rk2", "u", 2, -1, 0.5, -1)),
    ("sg_rk2_clamped", 64, 4, 32, 1, "preact", ("rk2", "u", 2, -1, 1.5, -1)),   # u > 1 is clamped: zero gradient
]

# integrate() with interior output times: name, C, H, W, B, solver tuple, output times
MULTITIME_CASES = [
    ("mt_rk2_n4", 64, 8, 32, 2, ("rk2", "u", 4, -1, 0.5, -1), [0.0, 0.3, 0.55, 1.0]),
    ("mt_rk4_n3_gridpoint", 64, 8, 32, 2, ("rk4", "u2", 3, -1, 1 / 3., -1), [0.0, 1 / 3., 0.9, 1.0]),
    ("mt_rk2_two_in_one_step", 64, 4, 32, 1, ("rk2", "u", 2, -1, 0.7, -1), [0.0, 0.1, 0.4, 0.5, 1.0]),
]

# the same for the time-dependent MNIST right-hand side (trained weights): tag, solver tuple
MNIST_SOLVER_GRAD_CASES = [
    ("msg_rk2_u05_n4", ("rk2", "u", 4, -1, 0.5, -1)),
    ("msg_rk2_u03_n2", ("rk2", "u", 2, -1, 0.3, -1)),
    ("msg_rk3_n2", ("rk3", "uv", 2, -1, 0.3, 0.7)),
    ("msg_rk4uv_n1", ("rk4", "uv", 1, -1, 0.3, 0.7)),
    ("msg_rk4u3_n2", ("rk4", "u3", 2, -1, 0.1, -1)),
]

# CIFAR pre-activation right-hand side with a per-sample normalisation inside the ODE block:
# name, C, H, W, B, normalisation key (cifar10/utils.py:26-36), num_groups, solver tuple
GN_CASES = [
    ("gn_c64_rk2_n3", 64, 8, 32, 2, "GN", 32, ("rk2", "u", 3, -1, 0.5, -1)),
    ("gn_c128_rk2_n2", 128, 16, 16, 2, "GN", 32, ("rk2", "u", 2, -1, 0.3, -1)),
    ("ln_c64_rk4_n1", 64, 8, 32, 2, "LN", 32, ("rk4", "u2", 1, -1, 1 / 3., -1)),
    ("in_c64_euler_n2", 64, 8, 32, 2, "IN", 32, ("euler", None, 2, -1, -1, -1)),
    ("gn_c16_odd_rk2_n2", 16, 5, 7, 3, "GN", 4, ("rk2", "u", 2, -1, 0.5, -1)),
]


# the same normalisations inside the POST-activation right-hand side (BasicBlock2)
GN_POST_CASES = [
    ("pgn_c64_rk2_n3", 64, 8, 32, 2, "GN", 32, ("rk2", "u", 3, -1, 0.5, -1)),
    ("pln_c128_rk2_n2", 128, 16, 16, 2, "LN", 32, ("rk2", "u", 2, -1, 0.3, -1)),
    ("pin_c64_rk4_n1", 64, 8, 32, 2, "IN", 32, ("rk4", "u2", 1, -1, 1 / 3., -1)),
    ("pgn_c16_odd_euler_n2", 16, 5, 7, 3, "GN", 4, ("euler", None, 2, -1, -1, -1)),
]


def gn_affine(C, k):
    """deterministic, non-trivial GroupNorm weight / bias of norm layer k"""
    return det_uniform((C,), 61 + k, 0.5, 1.5), det_uniform((C,), 71 + k, -0.3, 0.3)
