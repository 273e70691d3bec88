#!/usr/bin/env python
"""Golden vectors for the MNIST steady-state regulariser (`loss_options.ss_loss`, odenet_mnist/layers.py:53-93, 117-122;
examples/mnist/train_and_attack.py:215-222) from the REAL reference on the CPU: the trained ODE block, loss =
sum(y * r) + 0.1 * ss_loss.   -> tests/golden/ss_loss_mnist.npz"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

from oracle.detrand import det_normal  # noqa: E402
from sopa.src.solvers.utils import create_solver  # noqa: E402
import sopa.src.models.odenet_mnist.layers as mnist_layers  # noqa: E402
from make_golden_cases import WG_STRIDE  # noqa: E402

torch.set_num_threads(8)
w = np.load(os.path.join(HERE, "mnist_odeblock_weights.npz"))
feat = np.load(os.path.join(HERE, "mnist_odeblock.npz"))["feat"]
res = {}
for tag, sv, opts in (("standalone_rk2", [("rk2", "u", 4, -1, 0.5, -1)], Namespace(solver_mode="standalone")),
                      ("ensemble_rk2x2", [("rk2", "u", 2, -1, 0.3, -1), ("rk2", "u", 2, -1, 1.0, -1)],
                       Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=[0.25, 0.75]))):
    blk = mnist_layers.MetaODEBlock()
    rf = blk.rhs_func
    with torch.no_grad():
        for i in (1, 2, 3):
            getattr(rf, "norm%d" % i).weight.copy_(torch.from_numpy(w["norm%d_w" % i]))
            getattr(rf, "norm%d" % i).bias.copy_(torch.from_numpy(w["norm%d_b" % i]))
        for i in (1, 2):
            getattr(rf, "conv%d" % i)._layer.weight.copy_(torch.from_numpy(w["conv%d_w" % i]))
            getattr(rf, "conv%d" % i)._layer.bias.copy_(torch.from_numpy(w["conv%d_b" % i]))
    solvers = [create_solver(*s, torch.float32, "cpu") for s in sv]
    for s in solvers:
        s.freeze_params()
    x = torch.from_numpy(feat).requires_grad_(True)
    torch.manual_seed(0)
    y = blk(x, solvers, opts)
    ss = blk.ss_loss(y, solvers, opts)
    r = torch.from_numpy(det_normal(tuple(y.shape), 77))
    ((y * r).sum() + 0.1 * ss).backward()
    res[tag + "_ss"] = np.float32(ss.item())
    res[tag + "_y"] = y.detach().numpy()
    res[tag + "_gx"] = x.grad.numpy()
    res[tag + "_gconv1_w"] = rf.conv1._layer.weight.grad.numpy().reshape(-1)[::WG_STRIDE].copy()
    res[tag + "_gnorm3_b"] = rf.norm3.bias.grad.numpy().copy()
    res[tag + "_nfe"] = np.int64(rf.nfe)
    print(tag, float(ss), rf.nfe)
np.savez_compressed(os.path.join(HERE, "ss_loss_mnist.npz"), **res)
