"""Oracle: u/v-parametrized Butcher tableaus (Euler, RK2, RK3, RK4 families).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, with 1-element torch CPU tensors so every rounding step equals the reference's,
  * RK2  `sopa/src/solvers/rk_parametric_order2stage2.py:37-62`
  * RK3  `sopa/src/solvers/rk_parametric_order3stage3.py:25-68`
  * RK4  `sopa/src/solvers/rk_parametric_order4stage4.py:40-156`
  * Euler `sopa/src/solvers/euler.py:23-33`
  * driver order `_make_params_valid -> c -> b -> w`  `sopa/src/solvers/rk_parametric.py:68-75`

Returned tableau: dict(stages=s, c=[s], b=[s], w=[[...lower-triangular s x s...]]) of python
floats that are exactly the fp32 (or fp64) values the reference holds in its 1-elem tensors.
"""
import torch


def _eps_for(dtype):
    # order2stage2.py:56-60: fp64 params are clamped with fp32 eps, fp32 params with fp16 eps
    if dtype == torch.float64:
        return torch.finfo(torch.float32).eps
    if dtype == torch.float32:
        return torch.finfo(torch.float16).eps
    raise ValueError("oracle tableau: dtype must be float32 or float64")


def _t(val, dtype):
    return torch.tensor((val,), dtype=dtype)


def _valid_uv_pair(u_, v_, eps):
    # order3stage3.py:64-68 / order4stage4.py:150-156: keep u != v
    if u_ == v_:
        if u_ < 1.0 - eps:
            v_ = u_ + eps
        else:
            u_ = v_ - eps
    return u_, v_


def _rk2(u, dtype):
    eps = _eps_for(dtype)
    u_ = torch.clamp(u, eps, 1.0)                       # order2stage2.py:52-53
    c2 = u_.clone()                                     # :39
    b2 = 1.0 / (2 * u_)                                 # :43
    b1 = 1.0 - b2                                       # :44
    w21 = c2                                            # :48
    z = _t(0.0, dtype)
    return [z, c2], [b1, b2], [[z, z], [w21, z]]


def _rk3(u, v, dtype):
    eps = _eps_for(dtype)
    u_ = torch.clamp(u, eps, 1.0)                       # order3stage3.py:47-52
    v_ = torch.clamp(v, eps, 1.0)
    u_, v_ = _valid_uv_pair(u_, v_, eps)
    c2, c3 = u_.clone(), v_.clone()                     # :26-28
    v_sub_u = v_ - u_                                   # :32
    b2 = (2.0 - 3.0 * v_) / (6.0 * u_ * (-v_sub_u))     # :34
    b3 = (2.0 - 3.0 * u_) / (6.0 * v_ * v_sub_u)        # :35
    b1 = 1.0 - b2 - b3                                  # :36
    w32 = v_ * (v_ - u_) / (u_ * (2.0 - 3.0 * u_))      # :40
    w31 = c3 - w32                                      # :41
    w21 = c2                                            # :42
    z = _t(0.0, dtype)
    return [z, c2, c3], [b1, b2, b3], [[z, z, z], [w21, z, z], [w31, w32, z]]


def _rk4(param, u, v, dtype):
    eps = _eps_for(dtype)
    # order4stage4.py:127-156
    if v is not None:
        if u < 0.5:
            u_ = torch.clamp(u, eps, 0.5 - eps)
        else:
            u_ = torch.clamp(u, 0.5 + eps, 1.0 - eps)
        v_ = torch.clamp(v, eps, 1.0 - eps)
        u_, v_ = _valid_uv_pair(u_, v_, eps)
    else:
        u_ = torch.clamp(u, eps, 1.0 - eps)
        v_ = None
    z = _t(0.0, dtype)
    c1, c4 = z, _t(1.0, dtype)
    # order4stage4.py:40-59
    if param == "u1":
        c2, c3 = _t(0.5, dtype), _t(0.0, dtype)
    elif param == "u2":
        c2, c3 = _t(0.5, dtype), _t(0.5, dtype)
    elif param == "u3":
        c2, c3 = _t(1.0, dtype), _t(0.5, dtype)
    elif param == "uv":
        c2, c3 = u_.clone(), v_.clone()
    else:
        raise ValueError("oracle tableau: unknown rk4 parameterization %r" % (param,))
    # order4stage4.py:64-91
    if param == "u1":
        b1 = _t(1 / 6.0, dtype) - u_
        b2 = _t(2 / 3.0, dtype)
        b3 = u_.clone()
        b4 = _t(1 / 6.0, dtype)
    elif param == "u2":
        b1 = _t(1 / 6.0, dtype)
        b2 = _t(2 / 3.0, dtype) - u_
        b3 = u_.clone()
        b4 = _t(1 / 6.0, dtype)
    elif param == "u3":
        b1 = _t(1 / 6.0, dtype)
        b2 = _t(1 / 6.0, dtype) - u_
        b3 = _t(2 / 3.0, dtype)
        b4 = u_.clone()
    else:
        sub_u = 1.0 - u_
        sub_v = 1.0 - v_
        v_sub_u = v_ - u_
        b2 = (2.0 * v_ - 1.0) / (12 * u_ * sub_u * v_sub_u)
        b3 = (1.0 - 2 * u_) / (12 * v_ * sub_v * v_sub_u)
        b4 = (6.0 * u_ * v_ + 3.0 - 4.0 * u_ - 4.0 * v_) / (12 * sub_u * sub_v)
        b1 = 1.0 - b2 - b3 - b4
    # order4stage4.py:94-124 (Cramer's rule on the 2x2 order conditions)
    w43 = b3 * (1 - c3) / b4
    A00 = b3 * c3 * c2
    A01 = b4 * c4 * c2
    A10 = b3
    A11 = b4
    B0 = 0.125 - b4 * c4 * c3 * w43
    B1 = b2 * (1 - c2)
    detA = A00 * A11 - A01 * A10
    detA0 = B0 * A11 - B1 * A01
    detA1 = A00 * B1 - A10 * B0
    w32 = detA0 / detA
    w42 = detA1 / detA
    w41 = c4 - (w42 + w43)
    w31 = c3 - w32
    w21 = c2
    return ([c1, c2, c3, c4], [b1, b2, b3, b4],
            [[z, z, z, z], [w21, z, z, z], [w31, w32, z, z], [w41, w42, w43, z]])


def butcher_tableau_tensors(method, parameterization, u=None, v=None, dtype=torch.float32):
    """-> dict(stages, c, b, w) of 1-element TENSORS still connected to `u`, `v` by autograd: what the reference's solver
    holds after unfreeze_params() (order2stage2.py:104-109); used to restate the gradients w.r.t. u and v."""
    if method == "euler":
        c, b, w = [_t(0.0, dtype)], [_t(1.0, dtype)], [[_t(0.0, dtype)]]
    elif method == "rk2":
        c, b, w = _rk2(u, dtype)
    elif method == "rk3":
        c, b, w = _rk3(u, v, dtype)
    elif method == "rk4":
        c, b, w = _rk4(parameterization, u, v if parameterization == "uv" else None, dtype)
    else:
        raise ValueError("oracle tableau: unknown method %r" % (method,))
    return dict(stages=len(c), c=c, b=b, w=w)


def butcher_tableau(method, parameterization, u0=None, v0=None, dtype=torch.float32):
    """-> dict(stages, c, b, w) of python floats (exact images of the reference's tensors).

    `u0`,`v0` may be python/numpy scalars or 1-elem tensors (the *current* u, v of a solver).
    """
    def as_t(x):
        if x is None:
            return None
        if torch.is_tensor(x):
            return x.detach().to("cpu", dtype).reshape(1)
        return torch.tensor((x,), dtype=dtype)

    u, v = as_t(u0), as_t(v0)
    if method == "euler":                                 # euler.py:23-33
        c, b, w = [_t(0.0, dtype)], [_t(1.0, dtype)], [[_t(0.0, dtype)]]
    elif method == "rk2":
        if parameterization != "u":
            raise ValueError("Unknown parameterization for RKOrder2Stage2 solver")
        c, b, w = _rk2(u, dtype)
    elif method == "rk3":
        if parameterization != "uv":
            raise ValueError("Unknown parameterization for RKOrder3Stage3 solver")
        c, b, w = _rk3(u, v, dtype)
    elif method == "rk4":
        c, b, w = _rk4(parameterization, u, v if parameterization == "uv" else None, dtype)
    else:
        raise ValueError("oracle tableau: unknown method %r" % (method,))
    f = lambda x: float(x.item())
    return dict(stages=len(c), c=[f(x) for x in c], b=[f(x) for x in b],
                w=[[f(x) for x in row] for row in w])
