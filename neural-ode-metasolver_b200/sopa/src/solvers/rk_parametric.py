"""Solver objects of the `sopa` API (same names, attributes and error behaviour as the reference's
sopa/src/solvers/{rk_parametric,rk_parametric_order2stage2,..order3stage3,..order4stage4,euler}.py).

Design difference: a solver here is a HOST-side object.  u, v and the derived Butcher scalars are
1-element CPU tensors / numpy scalars computed with exactly the reference's fp32 (or fp64)
operation order, so no device kernels and no device->host syncs are spent on them
(the reference keeps them on `device`: rk_parametric_order2stage2.py:37-49 and syncs in
rk_parametric.py:95,109,117-119).  `integrate` hands the whole steps x stages loop to the fused
CUDA path (ops.ode_block_integrate); it does not evaluate `rhs_func` from Python.
"""
import abc

import numpy as np
import torch
import torch.nn as nn

from ....ops import ode_block_integrate, ode_block_integrate_gn, ode_block_integrate_mnist, ode_block_integrate_stacked
from .... import _cabi


def _np_dtype(dtype):
    if dtype == torch.float64:
        return np.float64
    if dtype == torch.float32:
        return np.float32
    raise ValueError("solver dtype must be torch.float32 or torch.float64")


def _clamp_eps(dtype):
    # fp64 parameters are clamped with the fp32 machine epsilon, fp32 ones with the fp16 epsilon
    # (rk_parametric_order2stage2.py:56-60)
    return float(torch.finfo(torch.float32).eps if dtype == torch.float64 else torch.finfo(torch.float16).eps)


class RKParametricSolver(object, metaclass=abc.ABCMeta):
    """Fixed-grid explicit RK integrator with a u/v-parametrized tableau (rk_parametric.py:5-113)."""

    n_stages = None
    method = None

    def __init__(self, n_steps=None, step_size=None, grid_constructor=None):
        given = [a for a in (n_steps, step_size, grid_constructor) if a is not None]
        if len(given) >= 2:
            raise ValueError("n_steps, step_size and grid_constructor are pairwise exclusive arguments.")
        if n_steps is not None:
            self.grid_constructor = self._grid_constructor_from_n_steps(n_steps)
        elif step_size is not None:
            self.grid_constructor = self._grid_constructor_from_step_size(step_size)
        elif grid_constructor is not None:
            self.grid_constructor = grid_constructor
        else:
            self.grid_constructor = lambda t: t

    # ---- time grids (rk_parametric.py:23-47); always evaluated on the host ----
    def _grid_constructor_from_step_size(self, step_size):
        step = float(step_size)

        def make(t):
            count = torch.ceil((t[-1] - t[0]) / step + 1).item()
            grid = torch.arange(0, count).to(t) * step + t[0]
            if grid[-1] > t[-1]:
                grid[-1] = t[-1]
            return grid
        return make

    def _grid_constructor_from_n_steps(self, n_steps):
        def make(t):
            # default-dtype CPU linspace, then cast -- the reference's exact recipe (rk_parametric.py:43)
            return torch.linspace(float(t[0]), float(t[-1]), int(n_steps) + 1).to(t)
        return make

    # ---- parameters -> tableau ----
    def _work_dtype(self):
        """The reference computes the tableau -- and picks the clamp epsilon -- in the dtype of the CURRENT `u` tensor
        (`_make_params_valid` tests `self.u.dtype`, rk_parametric_order2stage2.py:56-60), not in the dtype the solver was
        created with: after `noise_params` / `sample_solver_by_noising_params` (solvers/utils.py:60-105) u is a float32
        tensor even on a float64 solver."""
        u = getattr(self, "u", None)
        if isinstance(u, torch.Tensor) and u.dtype in (torch.float32, torch.float64):
            return u.dtype
        return self.dtype

    def _param_np(self, p):
        """current value of u / v as a numpy scalar of the solver dtype"""
        return _np_dtype(self._work_dtype())(p.detach().cpu().reshape(-1)[0].item())

    @abc.abstractmethod
    def _tableau_np(self):
        """-> (c list, b list, w lower-triangular rows, valid (u_, v_)) as numpy scalars"""

    def build_ButcherTableau(self, return_tableau=False):
        c, b, w, (u_, v_) = self._tableau_np()
        mk = lambda val: torch.tensor((float(val),), dtype=self._work_dtype())
        self.u_ = None if u_ is None else mk(u_)
        self.v_ = None if v_ is None else mk(v_)
        s = len(c)
        for i in range(s):
            setattr(self, "c%d" % (i + 1), mk(c[i]))
            setattr(self, "b%d" % (i + 1), mk(b[i]))
            for j in range(i + 1):
                setattr(self, "w%d%d" % (i + 1, j + 1), mk(w[i][j]))
        self._host_tableau = dict(stages=s, c=[float(x) for x in c], b=[float(x) for x in b],
                                  w=[[float(w[i][j]) if j < i else 0.0 for j in range(s)] for i in range(s)])
        if return_tableau:
            return self._collect_ButcherTableau()

    def _collect_ButcherTableau(self):
        t = self._host_tableau
        s = t["stages"]
        dt = self._work_dtype()      # torch.tensor([self.c1, self.c2]) of the reference inherits the scalars' dtype
        c = torch.tensor(t["c"], dtype=dt)
        w = [torch.tensor(t["w"][i][:i + 1], dtype=dt) for i in range(s)]
        b = torch.tensor(t["b"], dtype=dt)
        return c, w, b

    def host_tableau(self):
        return self._host_tableau

    # ---- requires_grad plumbing (API surface; every shipped script freezes: train_and_attack.py:413-414) ----
    def freeze_params(self):
        for p in (self.u, self.v):
            if p is not None:
                p.requires_grad = False
        self.build_ButcherTableau()

    def unfreeze_params(self):
        for p in (self.u, self.v):
            if p is not None:
                p.requires_grad = True
        self.build_ButcherTableau()

    # ---- differentiable tableau (only evaluated when u / v require grad) ----
    def _tableau_torch(self, u, v, eps):
        """-> (b list, w lower-triangular rows, c list) as float64 torch scalars, differentiable w.r.t. u, v: the same
        closed forms (and the same clamps, hence the same zero gradient outside the valid range) as `_tableau_np`."""
        raise NotImplementedError

    def tableau_coef(self):
        """Host float64 tensor [b_1..b_4 | w_11..w_44 | c_1..c_4] (_cabi.TABLEAU_GRAD_DOUBLES) connected to self.u /
        self.v by autograd -- the handle through which the fused backward returns dL/du, dL/dv (ops._OdeBlockFn)."""
        M = _cabi.MSB_MAX_STAGES
        u = self.u.to(torch.float64).reshape(()) if self.u is not None else None
        v = self.v.to(torch.float64).reshape(()) if self.v is not None else None
        b, w, c = self._tableau_torch(u, v, _clamp_eps(self._work_dtype()))
        zero = torch.zeros((), dtype=torch.float64)
        as_t = lambda t: t if torch.is_tensor(t) else torch.tensor(float(t), dtype=torch.float64)
        flat = [as_t(b[i]) if i < len(b) else zero for i in range(M)]
        for i in range(M):
            for j in range(M):
                flat.append(as_t(w[i][j]) if (i < len(w) and j < i) else zero)
        flat += [as_t(c[i]) if i < len(c) else zero for i in range(M)]
        return torch.stack([t.reshape(()).to(torch.float64) for t in flat])

    def _params_need_grad(self):
        return any(p is not None and torch.is_tensor(p) and p.requires_grad for p in (self.u, self.v))

    @property
    @abc.abstractmethod
    def order(self):
        pass

    # ---- integration ----
    def host_time_grid(self, t):
        t_host = t.detach().to("cpu", torch.float32)
        grid = self.grid_constructor(t_host)
        assert grid[0] == t_host[0] and grid[-1] == t_host[-1]        # rk_parametric.py:95
        return grid.to(torch.float32)

    def integrate(self, rhs_func, x, t):
        """-> Tensor[len(t), B, C, H, W]; differentiable w.r.t. x, rhs_func's parameters (and u, v when unfrozen).

        With more than two output times the solution at an interior t[j] is the linear interpolation between the two
        grid points that bracket it (rk_parametric.py:108-123): the fused path integrates grid segment by grid segment
        and interpolates with the reference's formula."""
        if len(t) == 2:
            return torch.stack((x, self.integrate_end(rhs_func, x, t)))
        spec, grid = self._fused_args(rhs_func, t)
        t_out = t.detach().to("cpu", torch.float32)
        solution = [x]
        j, start, y = 1, 0, x                                          # y = state at grid[start]
        for i in range(len(grid) - 1):
            t0, t1 = grid[i], grid[i + 1]
            if j < len(t_out) and t1 >= t_out[j]:
                if i > start:
                    y = self._integrate_grid(rhs_func, spec, y, grid[start:i + 1])
                y1 = self._integrate_grid(rhs_func, spec, y, grid[i:i + 2])
                while j < len(t_out) and t1 >= t_out[j]:
                    solution.append(self._linear_interp(t0, t1, y, y1, t_out[j]))
                    j += 1
                y, start = y1, i + 1
        return torch.stack(solution)

    @staticmethod
    def _linear_interp(t0, t1, y0, y1, t):
        # rk_parametric.py:116-123
        if t == t0:
            return y0
        if t == t1:
            return y1
        t0, t1, t = t0.to(y0[0]), t1.to(y0[0]), t.to(y0[0])           # 0-dim tensors on y's device: true division
        slope = (y1 - y0) / (t1 - t0)
        return y0 + slope * (t - t0)

    def integrate_end(self, rhs_func, x, t):
        """The end state y(t[-1]) alone, (B, C, H, W): what MetaODEBlock keeps of integrate()'s result
        (`y[-1]`, cifar10/layers.py:207) -- without materialising the stacked [x, y] copy."""
        if len(t) != 2:
            raise NotImplementedError("metasolver_b200: integrate_end() takes t = [t0, t1]; use integrate() for "
                                      "intermediate output times")
        spec, grid = self._fused_args(rhs_func, t)
        return self._integrate_grid(rhs_func, spec, x, grid)

    def _integrate_grid(self, rhs_func, spec, x, grid):
        """One fused call over the (sub-)grid `grid` (host fp32 tensor of >= 2 points) -> state at grid[-1]."""
        coef = self.tableau_coef() if (self._params_need_grad() and torch.is_grad_enabled()) else None
        if spec["rhs_kind"] == _cabi.RHS_MNIST_GN_T:
            y = ode_block_integrate_mnist(x, spec["params"], self._host_tableau, grid.tolist(), spec["groups"],
                                          spec["eps"], tableau_coef=coef)
        elif spec["rhs_kind"] in (_cabi.RHS_PREACT_GN, _cabi.RHS_POSTACT_GN):
            if coef is not None:
                raise NotImplementedError("metasolver_b200: gradients w.r.t. the solver parameters are not implemented for "
                                          "the GroupNorm CIFAR right-hand side; call freeze_params()")
            y = ode_block_integrate_gn(x, spec["params"], self._host_tableau, grid.tolist(), spec["groups"], spec["eps"],
                                       act=spec["act"], engine=spec.get("engine"), rhs_kind=spec["rhs_kind"])
        else:
            y = ode_block_integrate(x, spec["w1"], spec["w2"], self._host_tableau, grid.tolist(),
                                    rhs_kind=spec["rhs_kind"], act=spec["act"], engine=spec.get("engine"),
                                    tableau_coef=coef)
        rhs_func.nfe += self.n_stages * (len(grid) - 1)               # cifar10/layers.py:149
        return y

    def _fused_args(self, rhs_func, t):
        """Checks shared by integrate() and integrate_stacked(); -> (rhs spec, host time grid)."""
        spec = getattr(rhs_func, "fused_rhs_spec", None)
        if spec is None:
            raise NotImplementedError("metasolver_b200: %s is not a right-hand side the fused CUDA path knows; "
                                      "there is no unfused fallback" % type(rhs_func).__name__)
        return spec(), self.host_time_grid(t)

    def print_is_requires_grad(self):
        print('\nIs requires grad? (RK solver)')
        for name, p in self.__dict__.items():
            if hasattr(p, 'requires_grad'):
                print(name, p.requires_grad)


def can_stack(solvers, rhs_func, t):
    """True when `solvers` can share one set of launches on a stacked solver axis: same stage count,
    bit-identical time grids, a CIFAR-family right-hand side, at most MSB_MAX_SOLVERS of them."""
    if not (2 <= len(solvers) <= _cabi.MSB_MAX_SOLVERS):
        return False
    spec = getattr(rhs_func, "fused_rhs_spec", None)
    if spec is None or spec()["rhs_kind"] in (_cabi.RHS_MNIST_GN_T, _cabi.RHS_PREACT_GN, _cabi.RHS_POSTACT_GN):
        return False
    if any(s.n_stages != solvers[0].n_stages for s in solvers):
        return False
    g0 = solvers[0].host_time_grid(t)
    return all(torch.equal(s.host_time_grid(t), g0) for s in solvers[1:])


def integrate_stacked(solvers, rhs_func, x, t, replicate=True):
    """Stacked solver axis: every solver's stage evaluations run in the SAME kernel launches.

    replicate=True  (solver ensembling, cifar10/layers.py:198-203): x is (B,C,H,W); every solver integrates
                    the same x -> Tensor[K,B,C,H,W] of end states, slice k bit-identical to
                    solvers[k].integrate(rhs_func, x, t)[-1].
    replicate=False (model ensembling, fgsm.py:135-143, with the K model copies folded into the batch):
                    x is (K*B,C,H,W), slice k along dim 0 is integrated by solvers[k] -> Tensor[K*B,C,H,W]."""
    if not can_stack(solvers, rhs_func, t):
        raise ValueError("metasolver_b200: these solvers cannot share a stacked solver axis "
                         "(need equal stage counts and identical time grids, 2..%d solvers)" % _cabi.MSB_MAX_SOLVERS)
    spec, grid = solvers[0]._fused_args(rhs_func, t)
    for s in solvers[1:]:
        s._fused_args(rhs_func, t)
    tabs = [s.host_tableau() for s in solvers]
    kw = dict(rhs_kind=spec["rhs_kind"], act=spec["act"], engine=spec.get("engine"))
    if torch.is_grad_enabled() and any(s._params_need_grad() for s in solvers):
        # unfrozen u / v (unfreeze_params()): one row of Butcher coefficients per solver, built differentiably on the host;
        # the backward pass reduces dL/d(b, w) per slice and autograd chains each row to its solver's parameters
        kw["tableau_coef"] = torch.stack([s.tableau_coef() if s._params_need_grad()
                                          else torch.zeros(_cabi.TABLEAU_GRAD_DOUBLES, dtype=torch.float64) for s in solvers])
    if replicate:
        y = ode_block_integrate_stacked(x, spec["w1"], spec["w2"], tabs, grid.tolist(), **kw)
    else:
        y = ode_block_integrate(x, spec["w1"], spec["w2"], tabs, grid.tolist(), **kw)
    rhs_func.nfe += sum(s.n_stages for s in solvers) * (len(grid) - 1)
    return y


def _init_params(self, parameterization, u0, v0, dtype, device, with_v):
    self.dtype = dtype
    self.device = device          # kept for API compatibility; solver scalars live on the host
    self.parameterization = parameterization
    self.u = nn.Parameter(torch.tensor((u0,), dtype=dtype))
    self.u0 = torch.tensor((u0,), dtype=dtype)
    if with_v:
        self.v = nn.Parameter(torch.tensor((v0,), dtype=dtype))
        self.v0 = torch.tensor((v0,), dtype=dtype)
    else:
        self.v = None
        self.v0 = None


def _separate(u_, v_, eps, T):
    # keep the two nodes distinct (rk_parametric_order3stage3.py:64-68, order4stage4.py:150-156)
    if u_ == v_:
        if u_ < T(1.0 - eps):
            v_ = u_ + T(eps)
        else:
            u_ = v_ - T(eps)
    return u_, v_


class Euler(RKParametricSolver):
    n_stages, method = 1, "euler"

    def __init__(self, parameterization=None, u0=None, v0=None, dtype=None, device=None, **kwargs):
        super().__init__(**kwargs)
        self.dtype, self.device, self.parameterization = dtype, device, None
        self.u = self.u0 = self.v = self.v0 = None
        self.build_ButcherTableau()

    def _tableau_np(self):
        T = _np_dtype(self._work_dtype())
        return [T(0)], [T(1)], [[T(0)]], (None, None)

    def _tableau_torch(self, u, v, eps):
        return [torch.ones((), dtype=torch.float64)], [[]], [0.]

    def freeze_params(self):
        pass

    def unfreeze_params(self):
        pass

    @property
    def order(self):
        return 1


class RKOrder2Stage2(RKParametricSolver):
    """c2 = u, b2 = 1/(2u), b1 = 1 - b2, w21 = u  (rk_parametric_order2stage2.py:37-49)."""
    n_stages, method = 2, "rk2"

    def __init__(self, parameterization='u', u0=None, v0=None, dtype=None, device=None, **kwargs):
        super().__init__(**kwargs)
        if parameterization != 'u':
            raise ValueError('Unknown parameterization for RKOrder2Stage2 solver')
        _init_params(self, parameterization, u0, v0, dtype, device, with_v=False)
        self.build_ButcherTableau()

    def _tableau_np(self):
        T = _np_dtype(self._work_dtype())
        eps = _clamp_eps(self._work_dtype())
        u_ = min(max(self._param_np(self.u), T(eps)), T(1.0))
        b2 = T(1.0) / (T(2) * u_)
        b1 = T(1.0) - b2
        return [T(0), u_], [b1, b2], [[T(0)], [u_, T(0)]], (u_, None)

    def _tableau_torch(self, u, v, eps):
        u_ = torch.clamp(u, eps, 1.)                                   # order2stage2.py:52-53
        b2 = 1. / (2 * u_)
        return [1. - b2, b2], [[], [u_]], [0., u_]

    @property
    def order(self):
        return 2


class RKOrder3Stage3(RKParametricSolver):
    """Two-parameter 3-stage family (rk_parametric_order3stage3.py:25-44)."""
    n_stages, method = 3, "rk3"

    def __init__(self, parameterization='uv', u0=1 / 3., v0=2 / 3., dtype=None, device=None, **kwargs):
        super().__init__(**kwargs)
        if parameterization != 'uv':
            raise ValueError('Unknown parameterization for RKOrder3Stage3 solver')
        _init_params(self, parameterization, u0, v0, dtype, device, with_v=True)
        self.build_ButcherTableau()

    def _tableau_np(self):
        T = _np_dtype(self._work_dtype())
        eps = _clamp_eps(self._work_dtype())
        u_ = min(max(self._param_np(self.u), T(eps)), T(1.0))
        v_ = min(max(self._param_np(self.v), T(eps)), T(1.0))
        u_, v_ = _separate(u_, v_, eps, T)
        d = v_ - u_
        b2 = (T(2.0) - T(3.0) * v_) / (T(6.0) * u_ * (-d))
        b3 = (T(2.0) - T(3.0) * u_) / (T(6.0) * v_ * d)
        b1 = T(1.0) - b2 - b3
        w32 = v_ * (v_ - u_) / (u_ * (T(2.0) - T(3.0) * u_))
        w31 = v_ - w32
        return [T(0), u_, v_], [b1, b2, b3], [[T(0)], [u_, T(0)], [w31, w32, T(0)]], (u_, v_)

    def _tableau_torch(self, u, v, eps):
        u_, v_ = torch.clamp(u, eps, 1.), torch.clamp(v, eps, 1.)     # order3stage3.py:47-68
        if float(u_.detach()) == float(v_.detach()):
            if float(u_.detach()) < 1. - eps:
                v_ = u_ + eps
            else:
                u_ = v_ - eps
        d = v_ - u_
        b2 = (2. - 3. * v_) / (6. * u_ * (-d))
        b3 = (2. - 3. * u_) / (6. * v_ * d)
        b1 = 1. - b2 - b3
        w32 = v_ * (v_ - u_) / (u_ * (2. - 3. * u_))
        return [b1, b2, b3], [[], [u_], [v_ - w32, w32]], [0., u_, v_]

    @property
    def order(self):
        return 3


class RKOrder4Stage4(RKParametricSolver):
    """One-parameter families u1/u2/u3 and the two-parameter family uv (rk_parametric_order4stage4.py:40-156)."""
    n_stages, method = 4, "rk4"

    def __init__(self, parameterization='u2', u0=1 / 3., v0=2 / 3., dtype=torch.float64, device='cpu', **kwargs):
        super().__init__(**kwargs)
        if parameterization not in ('u1', 'u2', 'u3', 'uv'):
            raise ValueError('Unknown parameterization for RKOrder4Stage4 solver')
        _init_params(self, parameterization, u0, v0, dtype, device, with_v=(parameterization == 'uv'))
        self.build_ButcherTableau()

    def _tableau_np(self):
        T = _np_dtype(self._work_dtype())
        eps = _clamp_eps(self._work_dtype())
        u = self._param_np(self.u)
        kind = self.parameterization
        one, half = T(1.0), T(0.5)
        if self.v is not None:
            if u < half:
                u_ = min(max(u, T(eps)), T(0.5 - eps))
            else:
                u_ = min(max(u, T(0.5 + eps)), T(1.0 - eps))
            v_ = min(max(self._param_np(self.v), T(eps)), T(1.0 - eps))
            u_, v_ = _separate(u_, v_, eps, T)
        else:
            u_, v_ = min(max(u, T(eps)), T(1.0 - eps)), None
        sixth, two3 = T(1 / 6.), T(2 / 3.)
        if kind == 'u1':
            c2, c3 = half, T(0)
            b1, b2, b3, b4 = sixth - u_, two3, u_, sixth
        elif kind == 'u2':
            c2, c3 = half, half
            b1, b2, b3, b4 = sixth, two3 - u_, u_, sixth
        elif kind == 'u3':
            c2, c3 = one, half
            b1, b2, b3, b4 = sixth, sixth - u_, two3, u_
        else:
            c2, c3 = u_, v_
            su, sv, d = one - u_, one - v_, v_ - u_
            b2 = (T(2.0) * v_ - one) / (T(12) * u_ * su * d)
            b3 = (one - T(2) * u_) / (T(12) * v_ * sv * d)
            b4 = (T(6.0) * u_ * v_ + T(3.0) - T(4.0) * u_ - T(4.0) * v_) / (T(12) * su * sv)
            b1 = one - b2 - b3 - b4
        c4 = one
        # remaining order conditions: 2x2 linear system for (w32, w42), Cramer's rule (order4stage4.py:94-113)
        w43 = b3 * (one - c3) / b4
        a00, a01, a10, a11 = b3 * c3 * c2, b4 * c4 * c2, b3, b4
        r0 = T(0.125) - b4 * c4 * c3 * w43
        r1 = b2 * (one - c2)
        det = a00 * a11 - a01 * a10
        w32 = (r0 * a11 - r1 * a01) / det
        w42 = (a00 * r1 - a10 * r0) / det
        w41 = c4 - (w42 + w43)
        w31 = c3 - w32
        z = T(0)
        return ([z, c2, c3, c4], [b1, b2, b3, b4],
                [[z], [c2, z], [w31, w32, z], [w41, w42, w43, z]], (u_, v_))

    def _tableau_torch(self, u, v, eps):
        kind = self.parameterization
        T = lambda x: torch.tensor(float(x), dtype=torch.float64)
        if v is not None:                                              # order4stage4.py:127-156
            u_ = torch.clamp(u, eps, 0.5 - eps) if float(u.detach()) < 0.5 else torch.clamp(u, 0.5 + eps, 1. - eps)
            v_ = torch.clamp(v, eps, 1. - eps)
            if float(u_.detach()) == float(v_.detach()):
                if float(u_.detach()) < 1. - eps:
                    v_ = u_ + eps
                else:
                    u_ = v_ - eps
        else:
            u_, v_ = torch.clamp(u, eps, 1. - eps), None
        one, half, sixth, two3 = T(1.), T(0.5), T(1 / 6.), T(2 / 3.)
        if kind == 'u1':
            c2, c3 = half, T(0.)
            b1, b2, b3, b4 = sixth - u_, two3, u_, sixth
        elif kind == 'u2':
            c2, c3 = half, half
            b1, b2, b3, b4 = sixth, two3 - u_, u_, sixth
        elif kind == 'u3':
            c2, c3 = one, half
            b1, b2, b3, b4 = sixth, sixth - u_, two3, u_
        else:
            c2, c3 = u_, v_
            su, sv, d = one - u_, one - v_, v_ - u_
            b2 = (2. * v_ - one) / (12 * u_ * su * d)
            b3 = (one - 2 * u_) / (12 * v_ * sv * d)
            b4 = (6. * u_ * v_ + 3. - 4. * u_ - 4. * v_) / (12 * su * sv)
            b1 = one - b2 - b3 - b4
        c4 = one
        w43 = b3 * (one - c3) / b4
        a00, a01, a10, a11 = b3 * c3 * c2, b4 * c4 * c2, b3, b4
        r0 = 0.125 - b4 * c4 * c3 * w43
        r1 = b2 * (one - c2)
        det = a00 * a11 - a01 * a10
        w32 = (r0 * a11 - r1 * a01) / det
        w42 = (a00 * r1 - a10 * r0) / det
        w41 = c4 - (w42 + w43)
        w31 = c3 - w32
        return [b1, b2, b3, b4], [[], [c2], [w31, w32], [w41, w42, w43]], [0., c2, c3, c4]

    @property
    def order(self):
        return 4
