"""GPU parity tests: the CUDA ODE-block path (through the sopa API -> C ABI) against
(a) golden vectors produced by the real reference and (b) the CPU oracle on seeded inputs.
Tolerance: max|d|/max|ref| <= 1e-4 (BASELINE.json north_star), on outputs and on gradients."""
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
TOL = 1e-4


def _mods():
    import metasolver_b200  # noqa: F401
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2, BasicBlock2
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    return metasolver_b200, create_solver, MetaODEBlock, PreBasicBlock2, BasicBlock2, Identity


def _engines(C, H, W):
    import metasolver_b200
    from metasolver_b200 import _cabi
    eng = ["simt"]
    if _cabi.lib().msb_shape_supports_tcgen05(C, H, W):
        eng.append("tcgen05")
    return eng


def _run_block(case, engine, dev="cuda"):
    msb, create_solver, MetaODEBlock, PreBasicBlock2, BasicBlock2, Identity = _mods()
    name, C, H, W, B, kind, sv = case
    x, w1, w2, r = [torch.from_numpy(a).to(dev) for a in cases.ode_case_inputs(C, H, W, B)]
    cls = PreBasicBlock2 if kind == "preact" else BasicBlock2
    blk = MetaODEBlock(cls(C, norm_layer=Identity, act_layer=F.gelu)).to(dev)
    with torch.no_grad():
        blk.rhs_func.conv1.weight.copy_(w1)
        blk.rhs_func.conv2.weight.copy_(w2)
    solver = create_solver(*sv, torch.float32, dev)
    solver.freeze_params()
    msb.set_default_engine(engine)
    try:
        x.requires_grad_(True)
        y = blk(x, [solver], Namespace(solver_mode="standalone"))
        (y * r).sum().backward()
    finally:
        msb.set_default_engine("auto")
    torch.cuda.synchronize()
    return (y.detach().cpu().numpy(), x.grad.cpu().numpy(), blk.rhs_func.conv1.weight.grad.cpu().numpy(),
            blk.rhs_func.conv2.weight.grad.cpu().numpy(), blk.rhs_func.nfe)


PRE_CASES = list(cases.ODE_CASES)     # pre-activation (PreBasicBlock2) and post-activation (BasicBlock2) families


@pytest.mark.parametrize("case", PRE_CASES, ids=[c[0] for c in PRE_CASES])
def test_ode_block_vs_reference_golden(case):
    name, C, H, W, B, kind, sv = case
    g = golden("ode_%s.npz" % name)
    for engine in _engines(C, H, W):
        y, gx, gw1, gw2, nfe = _run_block(case, engine)
        assert nfe == int(g["nfe"])
        assert max_rel(y, g["y"]) <= TOL, (engine, max_rel(y, g["y"]))
        assert max_rel(gx, g["gx"]) <= TOL, (engine, max_rel(gx, g["gx"]))
        assert max_rel(gw1.reshape(-1)[::cases.WG_STRIDE], g["gw1"]) <= TOL, engine
        assert max_rel(gw2.reshape(-1)[::cases.WG_STRIDE], g["gw2"]) <= TOL, engine


def test_unsupported_configurations_raise():
    msb, create_solver, MetaODEBlock, PreBasicBlock2, BasicBlock2, Identity = _mods()
    x = torch.zeros(1, 64, 4, 32, device="cuda")
    solver = create_solver("rk2", "u", 2, -1, 0.5, -1, torch.float32, "cuda")
    solver.freeze_params()
    opts = Namespace(solver_mode="standalone")
    with pytest.raises(NotImplementedError):        # normalisation other than NF inside an ODE block
        MetaODEBlock(PreBasicBlock2(64, norm_layer=torch.nn.BatchNorm2d, act_layer=F.gelu)).cuda()(x, [solver], opts)
    with pytest.raises(NotImplementedError):        # activation the fused epilogue does not know
        MetaODEBlock(PreBasicBlock2(64, norm_layer=Identity, act_layer=F.softsign)).cuda()(x, [solver], opts)
    blk = MetaODEBlock(PreBasicBlock2(64, norm_layer=Identity, act_layer=F.gelu)).cuda()
    solver.unfreeze_params()                        # d/du IS implemented for the CIFAR right-hand sides
    y = blk(x.clone().requires_grad_(True), [solver], opts)   # (tests/test_gpu_solver_grads.py)
    assert y.requires_grad
    solver.freeze_params()
    with pytest.raises(NotImplementedError):        # integrate_end keeps the two-point contract of MetaODEBlock
        solver.integrate_end(blk.rhs_func, x, torch.tensor([0., 0.5, 1.]))


@pytest.mark.parametrize("C,H,W,B", [(64, 32, 32, 3), (128, 16, 16, 5), (64, 12, 32, 2), (32, 6, 6, 4)])
def test_single_conv_and_wgrad_vs_cpu(C, H, W, B):
    """Each GEMM kernel alone, through the C ABI, against fp64 CPU convolution of the same operands."""
    from metasolver_b200 import ops
    from oracle import det_normal, det_uniform
    x = torch.from_numpy(det_normal((B, C, H, W), 5)).cuda().contiguous(memory_format=torch.channels_last)
    go = torch.from_numpy(det_normal((B, C, H, W), 6)).cuda().contiguous(memory_format=torch.channels_last)
    w = torch.from_numpy(det_uniform((C, C, 3, 3), 7, -0.05, 0.05)).cuda()
    xs, _ = ops.act_split(x)
    gs, _ = ops.act_split(go)
    # what the kernels see: hi + lo
    xeff = (xs[:, :, 0].float() + xs[:, :, 1].float()).permute(0, 3, 1, 2).double().cpu()
    geff = (gs[:, :, 0].float() + gs[:, :, 1].float()).permute(0, 3, 1, 2).double().cpu()
    assert max_rel(xeff.numpy(), x.cpu().numpy()) < 2 ** -16
    wd = w.double().cpu()
    ref_f = F.conv2d(xeff, wd, None, 1, 1)
    ref_t = F.conv_transpose2d(geff, wd, None, 1, 1)
    xr = xeff.clone().requires_grad_(True)
    wr = wd.clone().requires_grad_(True)
    (F.conv2d(xr, wr, None, 1, 1) * geff).sum().backward()
    for engine in _engines(C, H, W):
        out = ops.conv3x3(xs, w, False, engine)
        assert max_rel(out.cpu().numpy(), ref_f.numpy()) < 2e-5, engine
        out_t = ops.conv3x3(gs, w, True, engine)
        assert max_rel(out_t.cpu().numpy(), ref_t.numpy()) < 2e-5, engine
        gw = ops.wgrad3x3(gs, xs, engine)
        assert max_rel(gw.cpu().numpy(), wr.grad.numpy()) < 2e-5, engine


def test_regimes_vs_reference_golden():
    msb, create_solver, MetaODEBlock, PreBasicBlock2, BasicBlock2, Identity = _mods()
    g = golden("regimes.npz")
    x, w1, w2, r = [torch.from_numpy(a).cuda() for a in cases.ode_case_inputs(64, 8, 32, 2)]
    blk = MetaODEBlock(PreBasicBlock2(64, norm_layer=Identity, act_layer=F.gelu)).cuda()
    with torch.no_grad():
        blk.rhs_func.conv1.weight.copy_(w1)
        blk.rhs_func.conv2.weight.copy_(w2)
    solvers = [create_solver(*sv, torch.float32, "cuda") for sv in cases.REGIME_SOLVERS]
    for s in solvers:
        s.freeze_params()
    with torch.no_grad():
        np.random.seed(123)
        ids = []
        for rep in range(3):
            opts = Namespace(solver_mode="switch", switch_probs=[0.1, 0.2, 0.3, 0.4])
            y = blk(x, solvers, opts)
            ids.append(opts.switch_solver_id)
            assert max_rel(y.cpu().numpy(), g["switch_y%d" % rep]) <= TOL
        assert ids == list(g["switch_ids"])
        torch.manual_seed(5)
        opts = Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=[0.4, 0.3, 0.2, 0.1])
        assert max_rel(blk(x, solvers, opts).cpu().numpy(), g["ens_weighted_y"]) <= TOL
        opts = Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=None)
        assert max_rel(blk(x, solvers, opts).cpu().numpy(), g["ens_uniform_y"]) <= TOL
        opts = Namespace(solver_mode="ensemble", ensemble_prob=0.0, ensemble_weights=None)
        assert max_rel(blk(x, solvers, opts).cpu().numpy(), g["ens_tails_y"]) <= TOL


def test_premetanode10_whole_model_vs_reference_golden():
    """Published config (NF + GeLU, RK2 u=.5, 8 steps): logits, ODE-block outputs, gradients, argmax."""
    msb, create_solver, *_ = _mods()
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import premetanode10, MetaODEBlock
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    from oracle import det_uniform
    from oracle.models import det_premetanode10_params, CIFAR_MEAN, CIFAR_STD
    g = golden("premetanode10.npz")
    model = premetanode10((Identity,) * 3, (lambda x: x,) * 3, (F.gelu,) * 3, in_planes=64, is_odenet=True)
    model.load_state_dict(det_premetanode10_params())
    model = model.cuda().eval()
    torch.backends.cudnn.allow_tf32 = False      # the non-ODE 5 % stays on PyTorch; keep it fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    img = torch.from_numpy(det_uniform((4, 3, 32, 32), 900, 0.0, 1.0))
    mean = torch.tensor(CIFAR_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CIFAR_STD).view(1, 3, 1, 1)
    x = ((img - mean) / std).cuda().requires_grad_(True)
    solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda")
    solver.freeze_params()
    taps = {}
    for n, m in model.named_modules():
        if isinstance(m, MetaODEBlock):
            m.register_forward_hook(lambda mod, i, o, n=n: taps.__setitem__(n, o.detach().cpu().numpy()))
    logits = model(x, [solver], Namespace(solver_mode="standalone"))
    loss = F.cross_entropy(logits, torch.tensor([3, 1, 4, 1]).cuda())
    loss.backward()
    assert max_rel(logits.detach().cpu().numpy(), g["logits"]) <= TOL
    assert (logits.argmax(1).cpu().numpy() == g["logits"].argmax(1)).all()
    for k, v in taps.items():
        assert max_rel(v, g["odeblock_" + k]) <= TOL, k
    assert max_rel(x.grad.cpu().numpy(), g["gx"]) <= TOL
    params = dict(model.named_parameters())
    for k in g.files:
        if k.startswith("g_"):
            got = params[k[2:]].grad.cpu().numpy().reshape(-1)[::cases.WG_STRIDE]
            assert max_rel(got, g[k]) <= TOL, (k, max_rel(got, g[k]))
    assert model.nfe == 32


def test_mnist_ode_block_trained_weights_vs_reference_golden():
    """BASELINE config 1: the MNIST ODE block (GN / ReLU / time-concatenated convs) with the trained weights
    shipped by the reference: forward and backward against the reference's own outputs / gradients."""
    import metasolver_b200 as msb
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_mnist.layers import MetaODEBlock, MetaNODE
    w = golden("mnist_odeblock_weights.npz")
    g = golden("mnist_odeblock.npz")
    blk = MetaODEBlock()
    rf = blk.rhs_func
    with torch.no_grad():
        for i in (1, 2, 3):
            getattr(rf, "norm%d" % i).weight.copy_(torch.from_numpy(w["norm%d_w" % i]))
            getattr(rf, "norm%d" % i).bias.copy_(torch.from_numpy(w["norm%d_b" % i]))
        for i in (1, 2):
            getattr(rf, "conv%d" % i)._layer.weight.copy_(torch.from_numpy(w["conv%d_w" % i]))
            getattr(rf, "conv%d" % i)._layer.bias.copy_(torch.from_numpy(w["conv%d_b" % i]))
    blk = blk.cuda()
    x = torch.from_numpy(g["feat"]).cuda()
    assert set(k for k in blk.state_dict()) == {
        "rhs_func.norm1.weight", "rhs_func.norm1.bias", "rhs_func.conv1._layer.weight", "rhs_func.conv1._layer.bias",
        "rhs_func.norm2.weight", "rhs_func.norm2.bias", "rhs_func.conv2._layer.weight", "rhs_func.conv2._layer.bias",
        "rhs_func.norm3.weight", "rhs_func.norm3.bias"}
    for tag, sv in (("rk2_u05_n8", ("rk2", "u", 8, -1, 0.5, -1)), ("rk4_u2_n2", ("rk4", "u2", 2, -1, 1 / 3., -1)),
                    ("euler_n4", ("euler", None, 4, -1, -1, -1))):
        solver = create_solver(*sv, torch.float32, "cuda")
        solver.freeze_params()
        with torch.no_grad():
            y = blk(x, [solver], Namespace(solver_mode="standalone"))
        assert max_rel(y.cpu().numpy(), g[tag + "_y"]) <= TOL, (tag, max_rel(y.cpu().numpy(), g[tag + "_y"]))
        # backward (fused discretize-then-optimize pass) against the reference's autograd
        from oracle import det_normal
        blk.zero_grad()
        xg = x.clone().requires_grad_(True)
        yg = blk(xg, [solver], Namespace(solver_mode="standalone"))
        assert torch.equal(yg.detach(), y)                       # tape-recording forward == inference forward
        r = torch.from_numpy(det_normal(tuple(yg.shape), 77)).cuda()
        (yg * r).sum().backward()
        # ReLU makes the gradient discontinuous in the forward values: a pre-activation within rounding
        # distance of zero can take the other side of the mask here than in the reference (observed: one element
        # of one sample for rk2_u05_n8).  The affected sample is bounded separately; every other sample -- and,
        # when no flip occurs, every parameter gradient -- must meet the 1e-4 bar.
        ga, gb = xg.grad.cpu().numpy().astype(np.float64), g[tag + "_gx"].astype(np.float64)
        per_sample = np.abs(ga - gb).reshape(ga.shape[0], -1).max(1) / np.abs(gb).max()
        flipped = per_sample > TOL
        assert flipped.sum() <= 1 and per_sample.max() <= 2e-3, (tag, per_sample)
        ptol = TOL if not flipped.any() else 5e-4
        got = {"gconv1_w": rf.conv1._layer.weight.grad.cpu().numpy().reshape(-1)[::cases.WG_STRIDE],
               "gconv2_b": rf.conv2._layer.bias.grad.cpu().numpy(), "gnorm1_w": rf.norm1.weight.grad.cpu().numpy(),
               "gnorm3_b": rf.norm3.bias.grad.cpu().numpy()}
        for k, v in got.items():
            assert max_rel(v, g[tag + "_" + k]) <= ptol, (tag, k, max_rel(v, g[tag + "_" + k]))
        # every parameter gradient against the CPU oracle on the same inputs
        import oracle
        po = {k: torch.from_numpy(w[k]).clone().requires_grad_(True) for k in w.files}
        xo = torch.from_numpy(g["feat"]).clone().requires_grad_(True)
        uv = {"rk2_u05_n8": ("rk2", "u", np.float32(0.5), None), "rk4_u2_n2": ("rk4", "u2", np.float32(1 / 3.), None),
              "euler_n4": ("euler", None, None, None)}[tag]
        tab = oracle.butcher_tableau(*uv)
        yo = oracle.integrate(tab, oracle.rhs_mnist(po), xo, torch.tensor([0, 1]).float(), n_steps=sv[2])[-1]
        (yo * r.cpu()).sum().backward()
        names = {"norm1_w": rf.norm1.weight, "norm1_b": rf.norm1.bias, "norm2_w": rf.norm2.weight, "norm2_b": rf.norm2.bias,
                 "norm3_w": rf.norm3.weight, "norm3_b": rf.norm3.bias, "conv1_w": rf.conv1._layer.weight,
                 "conv1_b": rf.conv1._layer.bias, "conv2_w": rf.conv2._layer.weight, "conv2_b": rf.conv2._layer.bias}
        for k, prm in names.items():
            assert max_rel(prm.grad.cpu().numpy(), po[k].grad.numpy()) <= ptol, (tag, k)
        # input-gradient-only mode (attacks)
        xg2 = x.clone().requires_grad_(True)
        with msb.input_grad_only():
            gx2, = torch.autograd.grad((blk(xg2, [solver], Namespace(solver_mode="standalone")) * r).sum(), [xg2])
        assert torch.equal(gx2, xg.grad)
    # full model wiring (stem / head are PyTorch): shapes only
    model = MetaNODE().cuda().eval()
    with torch.no_grad():
        out = model(torch.rand(4, 1, 28, 28, device="cuda"), [solver], Namespace(solver_mode="standalone"))
    assert out.shape == (4, 10)


@pytest.mark.parametrize("case", cases.MULTITIME_CASES, ids=[c[0] for c in cases.MULTITIME_CASES])
def test_integrate_with_interior_output_times(case):
    """solver.integrate(rhs, x, t) with len(t) > 2 (rk_parametric.py:104-123): grid segments run through the fused path,
    interior times are interpolated with the reference's formula; outputs, gradients and nfe vs the reference golden."""
    msb, create_solver, MetaODEBlock, PreBasicBlock2, BasicBlock2, Identity = _mods()
    name, C, H, W, B, sv, times = case
    g = golden("multitime.npz")
    x, w1, w2, r = [torch.from_numpy(a).cuda() for a in cases.ode_case_inputs(C, H, W, B)]
    rhs = PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu).cuda()
    with torch.no_grad():
        rhs.conv1.weight.copy_(w1)
        rhs.conv2.weight.copy_(w2)
    solver = create_solver(*sv, torch.float32, "cuda")
    solver.freeze_params()
    x.requires_grad_(True)
    ys = solver.integrate(rhs, x, torch.tensor(times))
    assert ys.shape[0] == len(times) and torch.equal(ys[0], x)
    sum(((k + 1.0) * ys[k] * r).sum() for k in range(1, len(times))).backward()
    assert rhs.nfe == int(g[name + "_nfe"])
    assert max_rel(ys.detach().cpu().numpy()[1:, :, ::3], g[name + "_y"]) <= TOL
    assert max_rel(x.grad.cpu().numpy(), g[name + "_gx"]) <= TOL
    assert max_rel(rhs.conv1.weight.grad.cpu().numpy().reshape(-1)[::cases.WG_STRIDE], g[name + "_gw1"]) <= TOL


def test_forward_under_no_grad_records_no_tape():
    """An evaluation forward of a TRAINABLE model under torch.no_grad() must not allocate the backward tape
    (n_steps * stages * 16 B per state element): the decision is taken by the caller of Function.apply, where grad
    mode is visible (ctx.needs_input_grad stays True for parameters under no_grad)."""
    from argparse import Namespace
    import torch.nn.functional as F
    import metasolver_b200  # noqa: F401
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    blk = MetaODEBlock(PreBasicBlock2(64, norm_layer=Identity, act_layer=F.gelu)).cuda()
    assert all(p.requires_grad for p in blk.parameters())
    solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda")
    solver.freeze_params()
    x = torch.randn(32, 64, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last)
    state = x.numel() * 4
    tape = 8 * 2 * 4 * state                                   # what a training forward records
    opts = Namespace(solver_mode="standalone")

    def peak(fn):
        torch.cuda.synchronize(); torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        fn()
        torch.cuda.synchronize()
        return torch.cuda.max_memory_allocated() - base
    with torch.no_grad():
        p_eval = peak(lambda: blk(x, [solver], opts))
    p_train = peak(lambda: blk(x, [solver], opts))
    assert p_train >= tape and p_eval < tape // 4, (p_eval, p_train, tape)
    with torch.no_grad():
        y_eval = blk(x, [solver], opts)
    assert torch.equal(y_eval, blk(x, [solver], opts).detach())   # same kernels, same numbers


def test_mnist_fused_single_launch_solve_matches_multi_launch_path():
    """mnist_fused.cu: the whole MNIST ODE-block solve in ONE persistent tcgen05 launch (state / stage derivatives in
    registers, operands in shared + tensor memory) against the round-1 multi-launch SIMT path: outputs, the recorded
    tape (through the gradients the unchanged backward derives from it) and the launch count."""
    import metasolver_b200 as msb
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_mnist.layers import MetaODEBlock
    torch.manual_seed(4)
    blk = MetaODEBlock().cuda()
    with torch.no_grad():
        for prm in blk.parameters():
            if prm.dim() == 1:
                prm.add_(0.1 * torch.randn_like(prm))
    d = msb.get_option("mnist_fused")
    try:
        for sv, B in ((("rk2", "u", 8, -1, 0.5, -1), 128), (("rk4", "uv", 3, -1, 1 / 3., 2 / 3.), 5), (("rk3", "uv", 2, -1, 0.4, 0.7), 300),
                      (("euler", None, 5, -1, -1, -1), 1)):
            solver = create_solver(*sv, torch.float32, "cuda")
            solver.freeze_params()
            x = torch.randn(B, 64, 6, 6, device="cuda")
            r = torch.randn(B, 64, 6, 6, device="cuda")
            res = {}
            for fused in (1, 0):
                msb.set_option("mnist_fused", fused)
                xg = x.clone().requires_grad_(True)
                blk.zero_grad()
                torch.cuda.synchronize()
                l0 = msb.launch_count()
                with torch.no_grad():
                    y_inf = blk(x, [solver], Namespace(solver_mode="standalone"))
                launches = msb.launch_count() - l0
                y = blk(xg, [solver], Namespace(solver_mode="standalone"))
                (y * r).sum().backward()
                res[fused] = [y_inf, y.detach(), xg.grad.clone()] + [p.grad.clone() for p in blk.parameters()] + [launches]
            assert torch.equal(res[1][0], res[1][1])                      # inference forward == tape-recording forward
            assert res[1][-1] <= 6 and res[0][-1] > 4 * res[1][-1], (res[1][-1], res[0][-1])   # ONE solve launch (+ 4 weight-pack / tap-map launches) against ~5 per stage evaluation
            for k, (a, b) in enumerate(zip(res[1][:-1], res[0][:-1])):
                a, b = a.cpu().numpy().astype(np.float64), b.cpu().numpy().astype(np.float64)
                if k < 2:                                   # outputs: different GEMM engines, same algorithm
                    assert max_rel(a, b) <= 2e-5, (sv, k, max_rel(a, b))
                elif k == 2:
                    # Input gradients pass through ReLU masks.  A pre-activation within rounding distance of zero can fall on
                    # the other side in the two forwards (different GEMM engines: ~1e-6 relative differences; ~70 000
                    # masked elements per image over 16 evaluations), which perturbs the whole gradient of THAT image by
                    # ~1e-3 (see the golden test above: one of 8 samples there).  Samples are independent: most must agree
                    # tightly, the flipped ones loosely.
                    per = np.abs(a - b).reshape(a.shape[0], -1).max(1) / np.abs(b).max()
                    # measured (scripts/diag_mnist_fused.py): no flip -> every gradient agrees to ~4e-6; 3-4 % of the images flip
                    assert np.median(per) <= 2e-5 and (per <= 5e-5).mean() >= 0.9 and per.max() <= 0.1, (sv, np.sort(per)[-5:])
                else:
                    # parameter gradients: sums over all images incl. the flipped ones (random weights, zero-mean GroupNorm
                    # outputs: flips are far more frequent here than with the trained weights of the golden test, which
                    # holds every parameter gradient of this path to 1e-4 / 5e-4 against the reference)
                    assert max_rel(a, b) <= 1e-2, (sv, k, max_rel(a, b))
    finally:
        msb.set_option("mnist_fused", d)
