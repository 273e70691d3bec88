"""Solver factory, solver smoothing and solver-ensemble samplers -- the API of
sopa/src/solvers/utils.py:13-117, host-side scalar work (torch CPU RNG, like the reference)."""
import copy

import numpy as np
import torch
from torch.distributions.cauchy import Cauchy
from torch.distributions.normal import Normal

from .rk_parametric import Euler, RKOrder2Stage2, RKOrder3Stage3, RKOrder4Stage4

_SOLVERS = {'euler': Euler, 'rk2': RKOrder2Stage2, 'rk3': RKOrder3Stage3, 'rk4': RKOrder4Stage4}


def create_solver(method, parameterization, n_steps, step_size, u0, v0, dtype, device):
    """Same signature and -1 -> None convention as sopa/src/solvers/utils.py:13-57."""
    n_steps = None if n_steps == -1 else n_steps
    step_size = None if step_size == -1 else step_size
    if dtype == torch.float64:
        u0, v0 = np.float64(u0), np.float64(v0)
    elif dtype == torch.float32:
        u0, v0 = np.float32(u0), np.float32(v0)
    cls = _SOLVERS.get(method)
    if cls is None:
        return None                                   # the reference falls through and returns None
    return cls(n_steps=n_steps, step_size=step_size, parameterization=parameterization,
               u0=u0, v0=v0, dtype=dtype, device=device)


def sample_noise(mu, sigma, noise_type='cauchy', size=1, device='cpu', minimize_rk2_error=False):
    """utils.py:60-72: draws from N / Cauchy centred at mu (or at 2/3 when minimize_rk2_error)."""
    if minimize_rk2_error:
        mu, sigma = 2 / 3., 2 / 3. * sigma
    if noise_type == 'cauchy':
        dist = Cauchy(torch.tensor([mu]), torch.tensor([sigma]))
    elif noise_type == 'normal':
        dist = Normal(torch.tensor([mu]), torch.tensor([sigma]))
    return torch.tensor([dist.sample() for _ in range(size)], device=device)


def noise_params(mean_u, mean_v=None, std=0.01, bernoulli_p=1.0, noise_type='cauchy', minimize_rk2_error=False):
    """utils.py:75-98: with probability p replace u (and v) by a noisy draw; u outside +-2 std -> mean."""
    gate = torch.distributions.Bernoulli(torch.tensor([bernoulli_p], dtype=torch.float32))
    v = None
    device = mean_u.device
    if gate.sample():
        std = torch.abs(torch.tensor(std, device=device))
        u = sample_noise(mean_u, std, noise_type=noise_type, size=1, device=device,
                         minimize_rk2_error=minimize_rk2_error)
        if u <= mean_u - 2 * std or u >= mean_u + 2 * std:
            u = mean_u
        if mean_v is not None:
            v = sample_noise(mean_v, std, noise_type=noise_type, size=1, device=device,
                             minimize_rk2_error=minimize_rk2_error)
    else:
        u = mean_u
        if mean_v is not None:
            v = mean_v
    return u, v


def sample_solver_by_noising_params(solver, std=0.01, bernoulli_p=1., noise_type='cauchy', minimize_rk2_error=False):
    """utils.py:100-110"""
    new_solver = copy.deepcopy(solver)
    new_solver.u, new_solver.v = noise_params(mean_u=new_solver.u0, mean_v=new_solver.v0, std=std,
                                              bernoulli_p=bernoulli_p, noise_type=noise_type,
                                              minimize_rk2_error=minimize_rk2_error)
    new_solver.build_ButcherTableau()
    print(new_solver.u, new_solver.v)
    return new_solver


def create_solver_ensemble_by_noising_params(solver, ensemble_size=1, kwargs_noise={}):
    """utils.py:112-117"""
    return [solver] + [sample_solver_by_noising_params(solver, **kwargs_noise) for _ in range(1, ensemble_size)]
