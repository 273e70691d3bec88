"""Oracle: fixed-grid parametrized Runge-Kutta ODE block (forward; backward via torch autograd).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates with plain torch CPU ops, keeping the reference's exact fp32 evaluation order:
  * time grid            `sopa/src/solvers/rk_parametric.py:23-47`
  * integrate loop       `sopa/src/solvers/rk_parametric.py:89-123`
  * _make_step           `rk_parametric_order2stage2.py:81-93`, `..order3stage3.py:88-103`,
                         `..order4stage4.py:175-192`, `euler.py:58-68`
  * CIFAR RHS modules    `sopa/src/models/odenet_cifar10/layers.py:108-121` (BasicBlock2),
                         `:148-161` (PreBasicBlock2)
  * MNIST RHS            `sopa/src/models/odenet_mnist/layers.py:158-171`, ConcatConv2d `:250-253`
  * regime dispatch      `sopa/src/models/odenet_cifar10/layers.py:173-207`
The reference differentiates by ordinary autograd over the unrolled graph (no adjoint), so the
oracle's gradients are simply `torch.autograd` through these functions.
"""
import numpy as np
import torch
import torch.nn.functional as F


class RhsCounter:
    """`rhs_func.nfe += 1` side effect (cifar10/layers.py:149, mnist/layers.py:159)."""

    def __init__(self):
        self.nfe = 0


def make_time_grid(t, n_steps=None, step_size=None):
    """rk_parametric.py:23-47.  `t` is a 1-D tensor of output times (already `type_as(x)`)."""
    if n_steps is not None:
        # :38-45 -- linspace runs with the *default* dtype on the CPU, then `.to(t)`
        return torch.linspace(float(t[0]), float(t[-1]), int(n_steps) + 1).to(t)
    if step_size is not None:
        # :25-33
        start_time, end_time = t[0], t[-1]
        step_size = float(step_size)
        n = torch.ceil((end_time - start_time) / step_size + 1).item()
        t_infer = torch.arange(0, n).to(t) * step_size + start_time
        if t_infer[-1] > t[-1]:
            t_infer[-1] = t[-1]
        return t_infer
    return t                                               # :20  (grid == requested times)


def _rk_step(tab, rhs, x, t, dt):
    """One step; returns dy.  Scalars enter as 1-elem tensors exactly as in the reference."""
    s = tab["stages"]
    # python floats (frozen solver) or 1-elem tensors connected to u / v (tableau.butcher_tableau_tensors)
    sc = lambda v: v if torch.is_tensor(v) else torch.tensor((v,), dtype=x.dtype)
    c, b, w = tab["c"], tab["b"], tab["w"]
    k = []
    for i in range(s):
        ti = t if i == 0 else t + sc(c[i]) * dt            # _get_t
        if i == 0:
            xi = x
        else:
            acc = k[0] * sc(w[i][0])
            for j in range(1, i):
                acc = acc + k[j] * sc(w[i][j])
            # RK2 writes `k1 * w21 * dt`, RK3/4 write `(k1*w31 + k2*w32) * dt`: same order.
            xi = x + acc * dt
        k.append(rhs(ti, xi))
    acc = k[0] * sc(b[0])
    for j in range(1, s):
        acc = acc + k[j] * sc(b[j])
    return acc * dt


def integrate(tab, rhs, x, t, n_steps=None, step_size=None, grid=None):
    """rk_parametric.py:89-113 -> stack of solutions at the times `t` (shape [len(t), *x.shape])."""
    t = t.type_as(x[0])
    time_grid = make_time_grid(t, n_steps, step_size) if grid is None else grid(t)
    assert time_grid[0] == t[0] and time_grid[-1] == t[-1]
    time_grid = time_grid.to(x[0])
    solution = [x]
    j = 1
    y0 = x
    for t0, t1 in zip(time_grid[:-1], time_grid[1:]):
        dy = _rk_step(tab, rhs, y0, t0, t1 - t0)
        y1 = y0 + dy
        while j < len(t) and t1 >= t[j]:
            solution.append(_linear_interp(t0, t1, y0, y1, t[j]))
            j += 1
        y0 = y1
    return torch.stack(solution)


def _linear_interp(t0, t1, y0, y1, t):
    # rk_parametric.py:116-123
    if t == t0:
        return y0
    if t == t1:
        return y1
    t0, t1, t = t0.to(y0[0]), t1.to(y0[0]), t.to(y0[0])
    slope = (y1 - y0) / (t1 - t0)
    return y0 + slope * (t - t0)


# ----------------------------------------------------------------------------- RHS families
def _act(name):
    if name == "gelu":
        return F.gelu                                      # exact-erf GeLU, cifar10/utils.py:67-68
    if name == "relu":
        return F.relu
    raise ValueError(name)


def rhs_preact(w1, w2, act="gelu", counter=None):
    """PreBasicBlock2 with NF (Identity) norm: conv2(act(conv1(act(x)))); cifar10/layers.py:148-161."""
    a = _act(act)

    def f(t, x):
        if counter is not None:
            counter.nfe += 1
        out = a(x)
        out = F.conv2d(out, w1, None, 1, 1)
        out = a(out)
        out = F.conv2d(out, w2, None, 1, 1)
        return out
    return f


def rhs_preact_gn(p, groups, eps=1e-5, act="gelu", counter=None, instance_norm=False):
    """PreBasicBlock2 with a per-sample normalisation ('GN' / 'LN' / 'IN' of cifar10/utils.py:26-36, all group norms):
    conv2(act(GN2(conv1(act(GN1(x)))))); cifar10/layers.py:148-161.  p: dict(norm{1,2}_{w,b}, conv{1,2}_w).
    instance_norm: 'IN' is nn.InstanceNorm2d (no affine parameters) -- mathematically one group per channel, but torch
    evaluates it with a different kernel, so the bit-exact restatement calls F.instance_norm."""
    a = _act(act)
    if instance_norm:
        norm = lambda z, k: F.instance_norm(z, eps=eps)
    else:
        norm = lambda z, k: F.group_norm(z, groups, p["norm%d_w" % k], p["norm%d_b" % k], eps)

    def f(t, x):
        if counter is not None:
            counter.nfe += 1
        out = a(norm(x, 1))
        out = F.conv2d(out, p["conv1_w"], None, 1, 1)
        out = a(norm(out, 2))
        out = F.conv2d(out, p["conv2_w"], None, 1, 1)
        return out
    return f


def rhs_postact_gn(p, groups, eps=1e-5, act="gelu", counter=None, instance_norm=False):
    """BasicBlock2 with a per-sample normalisation ('GN' / 'LN' / 'IN' of cifar10/utils.py:26-36):
    act(GN2(conv2(act(GN1(conv1(x)))))); cifar10/layers.py:108-121.  p: dict(norm{1,2}_{w,b}, conv{1,2}_w)."""
    a = _act(act)
    if instance_norm:
        norm = lambda z, k: F.instance_norm(z, eps=eps)
    else:
        norm = lambda z, k: F.group_norm(z, groups, p["norm%d_w" % k], p["norm%d_b" % k], eps)

    def f(t, x):
        if counter is not None:
            counter.nfe += 1
        out = F.conv2d(x, p["conv1_w"], None, 1, 1)
        out = a(norm(out, 1))
        out = F.conv2d(out, p["conv2_w"], None, 1, 1)
        out = a(norm(out, 2))
        return out
    return f


def rhs_postact(w1, w2, act="gelu", counter=None):
    """BasicBlock2 with NF norm: act(conv2(act(conv1(x)))); cifar10/layers.py:108-121."""
    a = _act(act)

    def f(t, x):
        if counter is not None:
            counter.nfe += 1
        out = F.conv2d(x, w1, None, 1, 1)
        out = a(out)
        out = F.conv2d(out, w2, None, 1, 1)
        out = a(out)
        return out
    return f


def rhs_mnist(p, counter=None):
    """MNIST ODEfunc: GN-ReLU-ConcatConv-GN-ReLU-ConcatConv-GN; mnist/layers.py:158-171, 250-253.

    `p` = dict(norm{1,2,3}_{w,b}: [64], conv{1,2}_w: [64,65,3,3], conv{1,2}_b: [64]).
    GroupNorm(min(32,dim), dim), eps 1e-5 (mnist/layers.py:208-209).
    """
    def cconv(t, x, w, b):
        tt = torch.ones_like(x[:, :1, :, :]) * t
        return F.conv2d(torch.cat([tt, x], 1), w, b, 1, 1)

    def f(t, x):
        if counter is not None:
            counter.nfe += 1
        g = min(32, x.shape[1])
        out = F.group_norm(x, g, p["norm1_w"], p["norm1_b"], 1e-5)
        out = F.relu(out)
        out = cconv(t, out, p["conv1_w"], p["conv1_b"])
        out = F.group_norm(out, g, p["norm2_w"], p["norm2_b"], 1e-5)
        out = F.relu(out)
        out = cconv(t, out, p["conv2_w"], p["conv2_b"])
        out = F.group_norm(out, g, p["norm3_w"], p["norm3_b"], 1e-5)
        return out
    return f


# ----------------------------------------------------------------------------- regimes
def ode_block_forward(x, rhs, tableaus, grids, mode="standalone", switch_probs=None,
                      ensemble_prob=1.0, ensemble_weights=None, t=None, record=None):
    """MetaODEBlock.forward, cifar10/layers.py:173-207 (identical in mnist/layers.py:16-50).

    `tableaus[i]` is a butcher_tableau dict, `grids[i]` a dict(n_steps=..)/(step_size=..).
    Uses the same host RNGs as the reference (numpy global for 'switch', torch CPU for the
    ensemble coin flip) so that seeding both sides identically reproduces the same draws.
    """
    if t is None:
        t = torch.tensor([0, 1]).float()
    n = len(tableaus)
    run = lambda i: integrate(tableaus[i], rhs, x, t, **grids[i])
    if mode == "standalone":
        y = run(0)
    elif mode == "switch":
        probs = switch_probs if switch_probs is not None else [1.0 / n for _ in range(n)]
        sid = np.random.choice(range(n), p=probs)
        if record is not None:
            record["switch_solver_id"] = sid
        y = run(sid)
    elif mode == "ensemble":
        coin = torch.bernoulli(torch.tensor((1,)), ensemble_prob)
        if record is not None:
            record["ensemble_coin_flip"] = coin
        if coin:
            ws = ensemble_weights if ensemble_weights is not None else [1.0 / n for _ in range(n)]
            for i, wi in enumerate(ws[:n]):
                if i == 0:
                    y = wi * run(i)
                else:
                    y += wi * run(i)
        else:
            y = run(0)
    else:
        raise ValueError(mode)
    return y[-1, :, :, :, :]
