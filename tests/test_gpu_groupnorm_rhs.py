"""GPU: the CIFAR pre- and post-activation right-hand sides with a per-sample normalisation inside the ODE block (SURVEY 8(f-3):
'GN', 'LN', 'IN' of sopa/src/models/odenet_cifar10/utils.py:26-36 -- all group norms) through the sopa API, against golden
vectors from the REAL reference: outputs, input gradient, conv and norm parameter gradients, nfe.  Tolerance 1e-4."""
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


def _engines(C, H, W):
    from metasolver_b200 import _cabi
    return ["simt"] + (["tcgen05"] if _cabi.lib().msb_shape_supports_tcgen05(C, H, W) else [])


ALL_GN = [(c, "pre") for c in cases.GN_CASES] + [(c, "post") for c in cases.GN_POST_CASES]


@pytest.mark.parametrize("case,order", ALL_GN, ids=[c[0] for c, _ in ALL_GN])
def test_group_norm_ode_block_vs_reference_golden(case, order):
    """order 'pre': PreBasicBlock2 (cifar10/layers.py:148-161); 'post': BasicBlock2 (cifar10/layers.py:108-121)."""
    import metasolver_b200 as msb
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2, BasicBlock2
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import get_normalization
    name, C, H, W, B, norm_key, groups, sv = case
    g = golden("gn_blocks.npz" if order == "pre" else "gn_post_blocks.npz")
    block_cls = PreBasicBlock2 if order == "pre" else BasicBlock2
    for engine in _engines(C, H, W):
        x, w1, w2, r = [torch.from_numpy(a).cuda() for a in cases.ode_case_inputs(C, H, W, B)]
        blk = MetaODEBlock(block_cls(C, norm_layer=get_normalization(norm_key, groups), act_layer=F.gelu)).cuda()
        rf = blk.rhs_func
        with torch.no_grad():
            rf.conv1.weight.copy_(w1)
            rf.conv2.weight.copy_(w2)
            if norm_key != "IN":
                for k, bn in enumerate((rf.bn1, rf.bn2)):
                    gw, gb = cases.gn_affine(C, k)
                    bn.weight.copy_(torch.from_numpy(gw))
                    bn.bias.copy_(torch.from_numpy(gb))
        solver = create_solver(*sv, torch.float32, "cuda")
        solver.freeze_params()
        msb.set_default_engine(engine)
        try:
            with torch.no_grad():
                y0 = blk(x, [solver], Namespace(solver_mode="standalone"))
            x.requires_grad_(True)
            y = blk(x, [solver], Namespace(solver_mode="standalone"))
            (y * r).sum().backward()
        finally:
            msb.set_default_engine("auto")
        assert torch.equal(y0, y.detach())                       # tape-recording forward == inference forward
        assert rf.nfe == 2 * int(g[name + "_nfe"])
        rf.nfe = 0
        assert max_rel(y.detach().cpu().numpy(), g[name + "_y"]) <= TOL, (engine, max_rel(y.detach().cpu().numpy(), g[name + "_y"]))
        assert max_rel(x.grad.cpu().numpy(), g[name + "_gx"]) <= TOL, (engine, max_rel(x.grad.cpu().numpy(), g[name + "_gx"]))
        assert max_rel(rf.conv1.weight.grad.cpu().numpy().reshape(-1)[::cases.WG_STRIDE], g[name + "_gw1"]) <= TOL, engine
        assert max_rel(rf.conv2.weight.grad.cpu().numpy().reshape(-1)[::cases.WG_STRIDE], g[name + "_gw2"]) <= TOL, engine
        if norm_key != "IN":
            for k, bn in enumerate((rf.bn1, rf.bn2)):
                assert max_rel(bn.weight.grad.cpu().numpy(), g["%s_gnorm%d_w" % (name, k + 1)]) <= TOL, (engine, k)
                assert max_rel(bn.bias.grad.cpu().numpy(), g["%s_gnorm%d_b" % (name, k + 1)]) <= TOL, (engine, k)


def test_group_norm_rhs_limits():
    import metasolver_b200  # noqa: F401
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2, BasicBlock2
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import get_normalization
    x = torch.randn(2, 64, 8, 32, device="cuda", requires_grad=True)
    s = create_solver("rk2", "u", 2, -1, 0.5, -1, torch.float32, "cuda")
    s.freeze_params()
    opts = Namespace(solver_mode="standalone")
    with pytest.raises(NotImplementedError):        # batch statistics couple the samples: not a per-sample right-hand side
        MetaODEBlock(PreBasicBlock2(64, norm_layer=get_normalization("BN"), act_layer=F.gelu)).cuda()(x, [s], opts)
    with pytest.raises(NotImplementedError):        # batch statistics in the post-activation block as well
        MetaODEBlock(BasicBlock2(64, norm_layer=get_normalization("BN"), act_layer=F.gelu)).cuda()(x, [s], opts)
    blk = MetaODEBlock(PreBasicBlock2(64, norm_layer=get_normalization("GN"), act_layer=F.gelu)).cuda()
    s.unfreeze_params()
    with pytest.raises(NotImplementedError):        # d/du through the GroupNorm right-hand side
        blk(x, [s], opts)
    s.freeze_params()
    # sample independence (what makes batch sharding exact) holds with per-sample normalisation
    with torch.no_grad():
        y = blk(x, [s], opts)
        y1 = blk(x[1:].contiguous(), [s], opts)
    assert torch.equal(y[1:], y1)
