#!/usr/bin/env python
"""Golden vectors for the CIFAR pre-activation right-hand side WITH a per-sample normalisation inside the ODE block
('GN', 'LN', 'IN' of sopa/src/models/odenet_cifar10/utils.py:26-36; PreBasicBlock2, layers.py:148-161) from the REAL
reference on the CPU.   -> tests/golden/gn_blocks.npz"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

from oracle.detrand import det_uniform  # noqa: E402
from sopa.src.solvers.utils import create_solver  # noqa: E402
from sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2, BasicBlock2  # noqa: E402
from sopa.src.models.odenet_cifar10.utils import get_normalization  # noqa: E402
from make_golden_cases import GN_CASES, GN_POST_CASES, WG_STRIDE, ode_case_inputs, gn_affine  # noqa: E402

torch.set_num_threads(8)
POST = "--post" in sys.argv          # python make_golden_gn.py --post  ->  gn_post_blocks.npz (BasicBlock2)
res = {}
for name, C, H, W, B, norm_key, groups, sv in (GN_POST_CASES if POST else GN_CASES):
    x, w1, w2, r = [torch.from_numpy(a) for a in ode_case_inputs(C, H, W, B)]
    blk = MetaODEBlock((BasicBlock2 if POST else PreBasicBlock2)(C, norm_layer=get_normalization(norm_key, groups), act_layer=F.gelu))
    rf = blk.rhs_func
    with torch.no_grad():
        rf.conv1.weight.copy_(w1)
        rf.conv2.weight.copy_(w2)
        if norm_key != "IN":
            for k, bn in enumerate((rf.bn1, rf.bn2)):
                gw, gb = gn_affine(C, k)
                bn.weight.copy_(torch.from_numpy(gw))
                bn.bias.copy_(torch.from_numpy(gb))
    solver = create_solver(*sv, torch.float32, "cpu")
    solver.freeze_params()
    x.requires_grad_(True)
    y = blk(x, [solver], Namespace(solver_mode="standalone"))
    (y * r).sum().backward()
    res[name + "_y"] = y.detach().numpy()
    res[name + "_gx"] = x.grad.numpy()
    res[name + "_gw1"] = rf.conv1.weight.grad.numpy().reshape(-1)[::WG_STRIDE].copy()
    res[name + "_gw2"] = rf.conv2.weight.grad.numpy().reshape(-1)[::WG_STRIDE].copy()
    if norm_key != "IN":
        for k, bn in enumerate((rf.bn1, rf.bn2)):
            res["%s_gnorm%d_w" % (name, k + 1)] = bn.weight.grad.numpy().copy()
            res["%s_gnorm%d_b" % (name, k + 1)] = bn.bias.grad.numpy().copy()
    res[name + "_nfe"] = np.int64(rf.nfe)
    print(name, float(np.abs(res[name + "_y"]).max()), rf.nfe, type(rf.bn1).__name__)
np.savez_compressed(os.path.join(HERE, "gn_post_blocks.npz" if POST else "gn_blocks.npz"), **res)
