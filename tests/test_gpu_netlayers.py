"""GPU parity of the non-ODE layers of premetanode10 now running on the library's own kernels:
the stem convolution (3 -> 64, + activation) and the strided pre-activation residual block (3x3 stride 2,
3x3, 1x1 stride-2 shortcut) -- forward, input gradient and weight gradients -- against a CPU restatement with
torch fp32 ops (cifar10/layers.py:77-81, 411-413) and against golden gradients of the real reference."""
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


def _rel(a, b):
    return max_rel(a.detach().cpu().numpy(), b.detach().numpy())


@pytest.mark.parametrize("B,H,W,C,act", [(3, 32, 32, 64, "gelu"), (2, 5, 7, 64, "relu"), (1, 9, 40, 128, "gelu")])
def test_stem_vs_cpu(B, H, W, C, act):
    import metasolver_b200 as msb
    from metasolver_b200 import _cabi, ops
    from oracle import det_normal
    x = torch.from_numpy(det_normal((B, 3, H, W), 5))
    w = torch.from_numpy(cases.conv_w(C, 3, 6))
    r = torch.from_numpy(det_normal((B, C, H, W), 7))
    fn = F.gelu if act == "gelu" else F.relu
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = fn(F.conv2d(xr, wr, None, 1, 1))
    (yr * r).sum().backward()
    xg, wg = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    before = msb.launch_count()
    y = ops.stem_conv_act(xg, wg, _cabi.ACT_GELU_ERF if act == "gelu" else _cabi.ACT_RELU)
    (y * r.cuda()).sum().backward()
    torch.cuda.synchronize()
    assert msb.launch_count() - before == 4        # fwd, wgrad + reduce, dgrad
    assert _rel(y, yr) <= 1e-5
    assert _rel(xg.grad, xr.grad) <= 1e-5
    assert _rel(wg.grad, wr.grad) <= 1e-5
    # input-gradient-only mode (attacks): no weight gradient is formed
    xg2, wg2 = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    with msb.input_grad_only():
        y2 = ops.stem_conv_act(xg2, wg2, _cabi.ACT_GELU_ERF if act == "gelu" else _cabi.ACT_RELU)
        gx, = torch.autograd.grad((y2 * r.cuda()).sum(), [xg2])
    assert torch.equal(gx, xg.grad)


@pytest.mark.parametrize("B,H,W,Ci,engine", [(2, 32, 32, 64, "tcgen05"), (2, 32, 32, 64, "simt"), (3, 6, 10, 8, "simt"),
                                             (1, 16, 64, 64, "auto")])
def test_strided_residual_block_vs_cpu(B, H, W, Ci, engine):
    import metasolver_b200 as msb
    from metasolver_b200 import ops
    from oracle import det_normal
    from oracle.models import _pre_basic_block
    Co = 2 * Ci
    x = torch.from_numpy(det_normal((B, Ci, H, W), 15))
    w1 = torch.from_numpy(cases.conv_w(Co, Ci, 16))
    w2 = torch.from_numpy(cases.conv_w(Co, Co, 17))
    wsc = torch.from_numpy(cases.conv_w(Co, Ci, 18, k=1))
    r = torch.from_numpy(det_normal((B, Co, H // 2, W // 2), 19))
    ref = [t.clone().requires_grad_(True) for t in (x, w1, w2, wsc)]
    yr = _pre_basic_block(ref[0], ref[1], ref[2], ref[3], 2)
    (yr * r).sum().backward()
    got = [t.cuda().requires_grad_(True) for t in (x, w1, w2, wsc)]
    y = ops.resblock_down(*got, engine=engine)
    (y * r.cuda()).sum().backward()
    torch.cuda.synchronize()
    assert _rel(y, yr) <= 1e-5
    for g, rf, name in zip(got, ref, ("gx", "gw1", "gw2", "gwsc")):
        assert _rel(g.grad, rf.grad) <= 2e-5, name
    # no-grad forward takes the tape-free path and must agree bit for bit
    with torch.no_grad():
        y0 = ops.resblock_down(*[t.detach() for t in got], engine=engine)
    assert torch.equal(y0, y)
    # input-gradient-only mode
    got2 = [t.cuda().requires_grad_(True) for t in (x, w1, w2, wsc)]
    with msb.input_grad_only():
        gx, = torch.autograd.grad((ops.resblock_down(*got2, engine=engine) * r.cuda()).sum(), [got2[0]])
    assert torch.equal(gx, got[0].grad)


def test_whole_model_runs_without_library_convolutions_and_matches_reference_gradients():
    """premetanode10 forward + loss + backward: EVERY compute kernel is one of this library's (stem, residual blocks, ODE
    blocks, pool + FC head, cross-entropy): no cuDNN / cuBLAS / ATen GEMM, convolution, reduction or loss kernel;
    gradients of ALL layers against the real reference."""
    import metasolver_b200
    from torch.profiler import profile, ProfilerActivity
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import premetanode10
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    from oracle import det_uniform
    from oracle.models import det_premetanode10_params, CIFAR_MEAN, CIFAR_STD
    g, g2 = golden("premetanode10.npz"), golden("premetanode10_resgrads.npz")
    model = premetanode10((Identity,) * 3, (lambda x: x,) * 3, (F.gelu,) * 3, in_planes=64, is_odenet=True)
    model.load_state_dict(det_premetanode10_params())
    model = model.cuda().eval()
    img = torch.from_numpy(det_uniform((4, 3, 32, 32), 900, 0.0, 1.0))
    x = ((img - torch.tensor(CIFAR_MEAN).view(1, 3, 1, 1)) / torch.tensor(CIFAR_STD).view(1, 3, 1, 1)).cuda()
    x.requires_grad_(True)
    solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda")
    solver.freeze_params()
    taps = {}
    model.layer2.blocks_res.register_forward_hook(lambda m, i, o: taps.__setitem__("l2", o.detach().cpu().numpy()))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        logits = model(x, [solver], Namespace(solver_mode="standalone"))
        metasolver_b200.cross_entropy(logits, torch.tensor([3, 1, 4, 1]).cuda()).backward()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
    foreign = [n for n in names if "msb::" not in n and
               any(t in n.lower() for t in ("cudnn", "convolve", "wgrad_alg", "dgrad_engine", "fft", "cgemm", "implicit_gemm",
                                            "conv2d", "gemm", "gemv", "cublas", "cutlass", "softmax", "nll_loss",
                                            "reduce_kernel", "mean"))]
    assert not foreign, foreign
    # what is left of ATen are layout copies / fills of the autograd glue (no arithmetic on activations)
    ours = [n for n in names if "msb::" in n]
    assert len(ours) >= 10, names
    print("non-msb kernels in a training step:", sorted(n for n in names if "msb::" not in n))
    assert max_rel(logits.detach().cpu().numpy(), g["logits"]) <= TOL
    assert max_rel(taps["l2"], g2["layer2_res_out"]) <= TOL
    assert max_rel(x.grad.cpu().numpy(), g["gx"]) <= TOL
    params = dict(model.named_parameters())
    for gold in (g, g2):
        for k in gold.files:
            if k.startswith("g_"):
                got = params[k[2:]].grad.cpu().numpy().reshape(-1)[::cases.WG_STRIDE]
                assert max_rel(got, gold[k]) <= TOL, (k, max_rel(got, gold[k]))


def test_pool_fc_head_and_cross_entropy_match_torch():
    """msb_pool_fc_forward / backward and msb_cross_entropy_* against AdaptiveAvgPool2d + Linear + F.cross_entropy
    (cifar10/layers.py:390-392, 425; train_and_attack.py:303-311): values and every gradient."""
    import metasolver_b200
    torch.manual_seed(0)
    for (B, C, HW, K) in ((37, 128, 16, 10), (4, 64, 32, 10), (512, 128, 16, 10)):
        x = torch.randn(B, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last)
        w = (torch.randn(K, C, device="cuda") / C ** 0.5)
        b = torch.randn(K, device="cuda") * 0.1
        y = torch.randint(0, K, (B,), device="cuda")
        xa, wa, ba = (t.clone().requires_grad_(True) for t in (x, w, b))
        la = metasolver_b200.cross_entropy(metasolver_b200.pool_fc(xa, wa, ba), y)
        la.backward()
        xb, wb, bb = (t.clone().double().requires_grad_(True) for t in (x, w, b))
        lb = F.cross_entropy(F.linear(F.adaptive_avg_pool2d(xb, (1, 1)).flatten(1), wb, bb), y)
        lb.backward()
        assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(lb)) + 1e-7
        for got, ref in ((xa.grad, xb.grad), (wa.grad, wb.grad), (ba.grad, bb.grad)):
            assert max_rel(got.cpu().numpy(), ref.cpu().numpy()) <= 1e-5
        # deterministic
        xa2, wa2, ba2 = (t.clone().requires_grad_(True) for t in (x, w, b))
        metasolver_b200.cross_entropy(metasolver_b200.pool_fc(xa2, wa2, ba2), y).backward()
        assert torch.equal(wa.grad, wa2.grad) and torch.equal(xa.grad, xa2.grad) and torch.equal(ba.grad, ba2.grad)


def test_layers_outside_the_fused_configurations_raise_unless_opted_in():
    """No silent cuDNN fallback: a BatchNorm residual block / stem, the post-activation BasicBlock and CPU tensors raise;
    metasolver_b200.set_library_fallback(True) is the explicit opt-in to PyTorch's kernels for those layers."""
    import torch.nn as nn
    import metasolver_b200
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import PreBasicBlock, BasicBlock, premetanode10
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    x = torch.randn(2, 64, 32, 32, device="cuda")
    bn_block = PreBasicBlock(64, 64, norm_layer=nn.BatchNorm2d, act_layer=F.gelu).cuda()
    post = BasicBlock(64, 64, norm_layer=Identity, act_layer=F.gelu).cuda()
    model = premetanode10((nn.BatchNorm2d,) * 3, (lambda m: m,) * 3, (F.gelu,) * 3, in_planes=64, is_odenet=True).cuda()
    for fn in (lambda: bn_block(x), lambda: post(x), lambda: model.conv1 and model(torch.randn(2, 3, 32, 32, device="cuda"), [], None)):
        with pytest.raises(NotImplementedError, match="set_library_fallback"):
            fn()
    metasolver_b200.set_library_fallback(True)
    try:
        assert bn_block(x).shape == x.shape and post(x).shape == x.shape
    finally:
        metasolver_b200.set_library_fallback(False)
    assert not metasolver_b200.library_fallback_allowed()
