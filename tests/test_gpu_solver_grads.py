"""GPU: gradients w.r.t. the solver parameters u, v (unfreeze_params()).

The device reduces dL/db_i = sum_n dt <gbar, k_i> and dL/dw_ij = sum_n dt <xbar_i, k_j>; the host chains them through the
closed-form tableau.  Checks:
  * the coefficient gradients against the fp64 oracle (differentiated through its own tableau tensors): <= 1e-4 of the
    largest coefficient gradient (no cancellation in these);
  * dL/du, dL/dv against the REAL reference's fp64 golden.  du = sum_i dL/dcoef_i * dcoef_i/du is a heavily cancelling
    sum (changing u keeps the order of the method: the reference's own fp32 run is 0.2-20 % off its fp64 run), so the
    tolerance is 1e-4 of the UN-cancelled magnitude sum_i |dL/dcoef_i * dcoef_i/du| plus the reference's own
    fp32-vs-fp64 deviation;
  * a clamped parameter (u > 1) gets exactly zero gradient; frozen solvers and weight / input gradients are unchanged;
  * K unfrozen solvers on a stacked solver axis get the gradients they get when integrated one by one.
"""
import os
import sys
from argparse import Namespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden, max_rel, ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_cases as cases  # noqa: E402

pytestmark = pytest.mark.gpu


def _oracle_coef_grads(case):
    from oracle import integrate, rhs_preact, rhs_postact
    from oracle.tableau import butcher_tableau_tensors
    name, C, H, W, B, kind, sv = case
    method, param, n_steps, step_size, u0, v0 = sv
    dt = torch.float64
    x, w1, w2, r = [torch.from_numpy(a).to(dt) for a in cases.ode_case_inputs(C, H, W, B)]
    u = torch.tensor((u0,), dtype=dt, requires_grad=True)
    v = torch.tensor((v0,), dtype=dt, requires_grad=True) if v0 != -1 else None
    tab = butcher_tableau_tensors(method, param, u, v, dt)
    S = tab["stages"]
    # detach the coefficients from u, v so that each gets its own gradient
    leaf = dict(stages=S, c=tab["c"], b=[t.detach().clone().requires_grad_(True) for t in tab["b"]],
                w=[[t.detach().clone().requires_grad_(True) for t in row] for row in tab["w"]])
    y = integrate(leaf, (rhs_preact if kind == "preact" else rhs_postact)(w1, w2), x, torch.tensor([0., 1.]), n_steps=n_steps)[-1]
    (y * r).sum().backward()
    gb = [float(t.grad) if t.grad is not None else 0.0 for t in leaf["b"]]
    gw = [[float(leaf["w"][i][j].grad) if (j < i and leaf["w"][i][j].grad is not None) else 0.0 for j in range(S)] for i in range(S)]
    return gb, gw


@pytest.mark.parametrize("case", cases.SOLVER_GRAD_CASES, ids=[c[0] for c in cases.SOLVER_GRAD_CASES])
def test_solver_parameter_gradients_vs_reference(case):
    import metasolver_b200  # noqa: F401
    from metasolver_b200 import _cabi
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2, BasicBlock2
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    name, C, H, W, B, kind, sv = case
    g = golden("solver_grads.npz")
    x, w1, w2, r = [torch.from_numpy(a).cuda() for a in cases.ode_case_inputs(C, H, W, B)]
    cls = PreBasicBlock2 if kind == "preact" else BasicBlock2
    blk = MetaODEBlock(cls(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
    with torch.no_grad():
        blk.rhs_func.conv1.weight.copy_(w1)
        blk.rhs_func.conv2.weight.copy_(w2)
    solver = create_solver(*sv, torch.float32, "cuda")
    solver.unfreeze_params()
    coefs = []
    orig = solver.tableau_coef
    solver.tableau_coef = lambda: (coefs.append(orig()), coefs[-1].retain_grad(), coefs[-1])[2]
    x.requires_grad_(True)
    y = blk(x, [solver], Namespace(solver_mode="standalone"))
    (y * r).sum().backward()
    torch.cuda.synchronize()
    # forward / input gradient are the frozen-solver ones
    assert max_rel(y.detach().cpu().numpy().reshape(-1)[::7], g["%s_f32_y" % name]) <= 1e-4
    assert max_rel(x.grad.cpu().numpy().reshape(-1)[::7], g["%s_f32_gx" % name]) <= 1e-4
    assert blk.rhs_func.conv1.weight.grad is not None
    # coefficient gradients vs the fp64 oracle
    M = _cabi.MSB_MAX_STAGES
    gc = coefs[0].grad.numpy()
    gb, gw = _oracle_coef_grads(case)
    S = len(gb)
    ref = np.zeros(_cabi.TABLEAU_GRAD_DOUBLES)               # [b | w | c]; c does not enter an autonomous right-hand side
    ref[:S] = gb
    for i in range(S):
        for j in range(i):
            ref[M + i * M + j] = gw[i][j]
    assert np.abs(gc - ref).max() <= 1e-4 * np.abs(ref).max(), (name, gc, ref)
    # u, v vs the reference's fp64 golden, yardstick = the reference's own fp32 deviation
    coef2 = orig()
    for p, key in ((solver.u, "du"), (solver.v, "dv")):
        if p is None:
            continue
        f64, f32 = float(g["%s_f64_%s" % (name, key)][0]), float(g["%s_f32_%s" % (name, key)][0])
        got = float(p.grad.reshape(-1)[0])
        jac = [torch.autograd.grad(coef2[i], p, retain_graph=True, allow_unused=True)[0] for i in range(coef2.numel())]
        jac = np.array([0.0 if j is None else float(j.reshape(-1)[0]) for j in jac])
        mag = float(np.abs(gc * jac).sum())
        assert abs(float(np.dot(gc, jac)) - got) <= 1e-6 * mag + 1e-12                       # host chain rule (u.grad is fp32)
        tol = 1e-4 * mag + 3.0 * abs(f32 - f64) + 1e-9
        assert abs(got - f64) <= tol, (name, key, got, f64, f32, mag)
        if name == "sg_rk2_clamped":
            assert got == 0.0


def test_frozen_solver_and_stacked_axis_behaviour():
    import metasolver_b200  # noqa: F401
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.solvers.rk_parametric import integrate_stacked
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
    from metasolver_b200.sopa.src.models.odenet_mnist.layers import MetaODEBlock as MnistBlock
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    torch.manual_seed(0)
    blk = MetaODEBlock(PreBasicBlock2(64, norm_layer=Identity, act_layer=F.gelu)).cuda()
    x = torch.randn(2, 64, 8, 32, device="cuda", requires_grad=True)
    s = create_solver("rk2", "u", 2, -1, 0.5, -1, torch.float32, "cuda")
    s.freeze_params()
    blk(x, [s], Namespace(solver_mode="standalone")).sum().backward()
    assert s.u.grad is None                                    # frozen: no solver gradient is formed
    s.unfreeze_params()
    with torch.no_grad():                                      # unfrozen but no grad mode: plain forward
        blk(x, [s], Namespace(solver_mode="standalone"))
    # unfrozen solvers share a stacked solver axis as frozen ones do (per-slice reductions of the coefficient gradients)
    from metasolver_b200.sopa.src.solvers.rk_parametric import can_stack
    s2 = create_solver("rk2", "u", 2, -1, 0.7, -1, torch.float32, "cuda")
    s2.freeze_params()
    assert can_stack([s, s2], blk.rhs_func, blk.integration_time)
    s.freeze_params()
    assert can_stack([s, s2], blk.rhs_func, blk.integration_time)


@pytest.mark.parametrize("method,param,uv", [("rk2", "u", [(0.5, -1), (0.3, -1), (0.8, -1)]),
                                             ("rk4", "uv", [(0.3, 0.7), (0.25, 0.6)])])
def test_solver_parameter_gradients_on_a_stacked_solver_axis(method, param, uv):
    """Solver ensembling (cifar10/layers.py:198-203) with unfrozen solvers: all K solvers in the same launches, each u / v
    receiving the gradient reduced over its own slice == the gradient of the solver integrated alone (which the test
    above pins to the reference)."""
    import metasolver_b200  # noqa: F401
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.solvers.rk_parametric import integrate_stacked
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    C, H, W, B = 64, 8, 32, 2
    x0, w1, w2, r = [torch.from_numpy(a).cuda() for a in cases.ode_case_inputs(C, H, W, B)]
    blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
    with torch.no_grad():
        blk.rhs_func.conv1.weight.copy_(w1)
        blk.rhs_func.conv2.weight.copy_(w2)
    K = len(uv)
    weights = [1.0, -0.5, 2.0][:K]                       # a different loss weight per solver: slices must not mix

    def make():
        ss = [create_solver(method, param, 2, -1, u, v, torch.float32, "cuda") for u, v in uv]
        for s in ss:
            s.unfreeze_params()
        return ss
    # one by one
    alone = make()
    xa = x0.clone().requires_grad_(True)
    blk.zero_grad()
    loss = sum(wk * (s.integrate(blk.rhs_func, xa, blk.integration_time)[-1] * r).sum() for wk, s in zip(weights, alone))
    loss.backward()
    gw_alone = blk.rhs_func.conv1.weight.grad.clone()
    # stacked
    stacked = make()
    stacked[-1].freeze_params()                          # a frozen member among unfrozen ones gets no gradient
    xs = x0.clone().requires_grad_(True)
    blk.zero_grad()
    l0 = metasolver_b200.launch_count()
    ys = integrate_stacked(stacked, blk.rhs_func, xs, blk.integration_time, replicate=True)
    assert tuple(ys.shape) == (K, B, C, H, W)
    sum(wk * (ys[k] * r).sum() for k, wk in enumerate(weights)).backward()
    torch.cuda.synchronize()
    assert max_rel(xs.grad.cpu().numpy(), xa.grad.cpu().numpy()) <= 1e-5
    assert max_rel(blk.rhs_func.conv1.weight.grad.cpu().numpy(), gw_alone.cpu().numpy()) <= 1e-5
    for k in range(K - 1):
        for name in ("u", "v"):
            pa, ps = getattr(alone[k], name), getattr(stacked[k], name)
            if pa is None:
                continue
            ga, gs = float(pa.grad.reshape(-1)[0]), float(ps.grad.reshape(-1)[0])
            assert abs(ga - gs) <= 1e-5 * max(abs(ga), 1e-3) + 1e-7, (k, name, ga, gs)
    assert stacked[-1].u.grad is None


@pytest.mark.parametrize("case", cases.MNIST_SOLVER_GRAD_CASES, ids=[c[0] for c in cases.MNIST_SOLVER_GRAD_CASES])
def test_solver_parameter_gradients_mnist_vs_reference(case):
    """Time-dependent MNIST right-hand side (trained weights): dL/du includes the node terms dL/dc_i = dt <dP, dP/dt>
    (t_i = t_n + c_i dt enters both time-concatenated convolutions).  vs the REAL reference's fp64 golden."""
    import metasolver_b200  # noqa: F401
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_mnist.layers import MetaODEBlock as MnistBlock
    from oracle import det_normal
    tag, sv = case
    g = golden("solver_grads_mnist.npz")
    w = golden("mnist_odeblock_weights.npz")
    feat = golden("mnist_odeblock.npz")["feat"]
    blk = MnistBlock().cuda()
    rf = blk.rhs_func
    with torch.no_grad():
        for i in (1, 2, 3):
            getattr(rf, "norm%d" % i).weight.copy_(torch.from_numpy(w["norm%d_w" % i]))
            getattr(rf, "norm%d" % i).bias.copy_(torch.from_numpy(w["norm%d_b" % i]))
        for i in (1, 2):
            getattr(rf, "conv%d" % i)._layer.weight.copy_(torch.from_numpy(w["conv%d_w" % i]))
            getattr(rf, "conv%d" % i)._layer.bias.copy_(torch.from_numpy(w["conv%d_b" % i]))
    solver = create_solver(*sv, torch.float32, "cuda")
    solver.unfreeze_params()
    x = torch.from_numpy(feat).cuda().requires_grad_(True)
    y = blk(x, [solver], Namespace(solver_mode="standalone"))
    r = torch.from_numpy(det_normal(tuple(y.shape), 77)).cuda()
    (y * r).sum().backward()
    # ReLU masks can flip for pre-activations within rounding distance of zero (see test_gpu_odeblock.py): 2e-3 bound
    assert max_rel(x.grad.cpu().numpy(), g[tag + "_f32_gx"]) <= 2e-3
    assert rf.conv1._layer.weight.grad is not None
    for p, key in ((solver.u, "du"), (solver.v, "dv")):
        if p is None:
            continue
        f64, f32 = float(g["%s_f64_%s" % (tag, key)][0]), float(g["%s_f32_%s" % (tag, key)][0])
        got = float(p.grad.reshape(-1)[0])
        assert abs(got - f64) <= 2e-3 * abs(f64) + 3.0 * abs(f32 - f64), (tag, key, got, f64, f32)


@pytest.mark.parametrize("tag,svs,opts", [
    ("standalone_rk2", [("rk2", "u", 4, -1, 0.5, -1)], Namespace(solver_mode="standalone")),
    ("ensemble_rk2x2", [("rk2", "u", 2, -1, 0.3, -1), ("rk2", "u", 2, -1, 1.0, -1)],
     Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=[0.25, 0.75]))])
def test_mnist_steady_state_regulariser_vs_reference(tag, svs, opts):
    """`loss_options.ss_loss` (odenet_mnist/layers.py:53-93,117-122): one more unit of time integrated from the block
    output on the fused path; value, nfe and the gradients of (sum(y r) + 0.1 ss_loss) vs the REAL reference."""
    import metasolver_b200  # noqa: F401
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_mnist.layers import MetaODEBlock as MnistBlock, MetaNODE
    from oracle import det_normal
    g = golden("ss_loss_mnist.npz")
    w = golden("mnist_odeblock_weights.npz")
    feat = golden("mnist_odeblock.npz")["feat"]
    blk = MnistBlock().cuda()
    rf = blk.rhs_func
    with torch.no_grad():
        for i in (1, 2, 3):
            getattr(rf, "norm%d" % i).weight.copy_(torch.from_numpy(w["norm%d_w" % i]))
            getattr(rf, "norm%d" % i).bias.copy_(torch.from_numpy(w["norm%d_b" % i]))
        for i in (1, 2):
            getattr(rf, "conv%d" % i)._layer.weight.copy_(torch.from_numpy(w["conv%d_w" % i]))
            getattr(rf, "conv%d" % i)._layer.bias.copy_(torch.from_numpy(w["conv%d_b" % i]))
    solvers = [create_solver(*s, torch.float32, "cuda") for s in svs]
    for s in solvers:
        s.freeze_params()
    x = torch.from_numpy(feat).cuda().requires_grad_(True)
    torch.manual_seed(0)
    y = blk(x, solvers, opts)
    ss = blk.ss_loss(y, solvers, opts)
    r = torch.from_numpy(det_normal(tuple(y.shape), 77)).cuda()
    ((y * r).sum() + 0.1 * ss).backward()
    assert rf.nfe == int(g[tag + "_nfe"])
    assert abs(float(ss) - float(g[tag + "_ss"])) <= 1e-4 * abs(float(g[tag + "_ss"]))
    assert max_rel(y.detach().cpu().numpy(), g[tag + "_y"]) <= 1e-4
    assert max_rel(x.grad.cpu().numpy(), g[tag + "_gx"]) <= 2e-3            # ReLU-mask flips, see test_gpu_odeblock.py
    assert max_rel(rf.conv1._layer.weight.grad.cpu().numpy().reshape(-1)[::cases.WG_STRIDE], g[tag + "_gconv1_w"]) <= 5e-4
    assert max_rel(rf.norm3.bias.grad.cpu().numpy(), g[tag + "_gnorm3_b"]) <= 5e-4
    # whole-model wiring: MetaNODE accumulates the regulariser when loss_options.ss_loss is set
    model = MetaNODE().cuda().eval()
    out = model(torch.rand(4, 1, 28, 28, device="cuda"), solvers, opts, Namespace(ss_loss=True))
    assert out.shape == (4, 10) and float(model.get_ss_loss()) > 0
    model(torch.rand(4, 1, 28, 28, device="cuda"), solvers, opts, Namespace(ss_loss=False))
    assert model.get_ss_loss() == 0
