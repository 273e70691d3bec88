#!/usr/bin/env python
"""Golden for SURVEY 8 a12: the smoothing / ensemble samplers of the REAL reference (sopa/src/solvers/utils.py:60-117:
noise_params, sample_solver_by_noising_params, create_solver_ensemble_by_noising_params) under fixed torch seeds --
the drawn u / v and the tableaus rebuilt from them, for float32 AND float64 solvers (the draw is a float32 tensor either
way, which changes the clamp epsilon and the arithmetic dtype of a float64 solver's tableau).

    python tests/golden/make_golden_noise.py      ->  tests/golden/noise_samplers.npz
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [  # name, create_solver args (without dtype/device), noise kwargs, seed
    ("rk2_normal", ("rk2", "u", 8, -1, 0.5, -1), dict(std=0.0125, bernoulli_p=1.0, noise_type="normal"), 11),
    ("rk2_cauchy", ("rk2", "u", 8, -1, 0.5, -1), dict(std=0.01, bernoulli_p=1.0, noise_type="cauchy"), 12),
    ("rk2_bern", ("rk2", "u", 4, -1, 2 / 3., -1), dict(std=0.05, bernoulli_p=0.5, noise_type="normal"), 13),
    ("rk2_minerr", ("rk2", "u", 4, -1, 0.5, -1), dict(std=0.02, bernoulli_p=1.0, noise_type="normal", minimize_rk2_error=True), 14),
    ("rk4uv_normal", ("rk4", "uv", 2, -1, 1 / 3., 2 / 3.), dict(std=0.02, bernoulli_p=1.0, noise_type="normal"), 15),
    ("rk3uv_cauchy", ("rk3", "uv", 2, -1, 1 / 3., 2 / 3.), dict(std=0.01, bernoulli_p=1.0, noise_type="cauchy"), 16),
]
N_DRAWS = 6


def tab(s):
    c, w, b = s.build_ButcherTableau(return_tableau=True)
    flat = [float(x) for x in c] + [float(x) for x in b]
    for row in w:
        flat += [float(x) for x in np.atleast_1d(row.detach().numpy() if torch.is_tensor(row) else row)]
    return np.asarray(flat, dtype=np.float64)


def main():
    sys.path.insert(0, "/root/reference")
    from sopa.src.solvers.utils import (create_solver, sample_solver_by_noising_params,
                                        create_solver_ensemble_by_noising_params)
    out = {}
    for name, args, kw, seed in CASES:
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            solver = create_solver(*args, dt, "cpu")
            solver.freeze_params()
            torch.manual_seed(seed)
            us, vs, tabs = [], [], []
            with contextlib.redirect_stdout(io.StringIO()):
                for _ in range(N_DRAWS):
                    s2 = sample_solver_by_noising_params(solver, **kw)
                    us.append(float(s2.u))
                    vs.append(float(s2.v) if s2.v is not None else np.nan)
                    tabs.append(tab(s2))
                ens = create_solver_ensemble_by_noising_params(solver, ensemble_size=4, kwargs_noise=kw)
            out["%s_%s_u" % (name, tag)] = np.asarray(us, dtype=np.float64)
            out["%s_%s_v" % (name, tag)] = np.asarray(vs, dtype=np.float64)
            out["%s_%s_tab" % (name, tag)] = np.stack(tabs)
            out["%s_%s_ens_u" % (name, tag)] = np.asarray([float(e.u) for e in ens], dtype=np.float64)
            out["%s_%s_ens_tab" % (name, tag)] = np.stack([tab(e) for e in ens])
            out["%s_%s_udtype" % (name, tag)] = np.asarray([str(s2.u.dtype)])
    np.savez(os.path.join(HERE, "noise_samplers.npz"), **out)
    print("wrote noise_samplers.npz:", len(out), "arrays;", {k: out[k] for k in ("rk2_normal_f64_u", "rk2_normal_f64_udtype")})


if __name__ == "__main__":
    main()
