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
    """premetanode10 forward+backward: every convolution is one of this library's kernels (torch only sees the
    pooling / FC head); gradients of ALL layers against the real reference."""
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
        F.cross_entropy(logits, torch.tensor([3, 1, 4, 1]).cuda()).backward()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
    foreign = [n for n in names if any(t in n.lower() for t in ("cudnn", "convolve", "wgrad_alg", "dgrad_engine", "fft", "cgemm",
                                                                 "implicit_gemm", "conv2d"))]
    assert not foreign, foreign
    assert max_rel(logits.detach().cpu().numpy(), g["logits"]) <= TOL
    assert max_rel(taps["l2"], g2["layer2_res_out"]) <= TOL
    assert max_rel(x.grad.cpu().numpy(), g["gx"]) <= TOL
    params = dict(model.named_parameters())
    for gold in (g, g2):
        for k in gold.files:
            if k.startswith("g_"):
                got = params[k[2:]].grad.cpu().numpy().reshape(-1)[::cases.WG_STRIDE]
                assert max_rel(got, gold[k]) <= TOL, (k, max_rel(got, gold[k]))
