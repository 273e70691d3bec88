#!/usr/bin/env python
"""Golden gradients w.r.t. the solver parameters u, v for the TIME-DEPENDENT MNIST right-hand side (trained ODE-block
weights shipped with the reference): here t_i = t_n + c_i dt enters the convolutions, so dL/du also flows through c_i
(rk_parametric_order2stage2.py:81-86).  REAL reference on the CPU, fp32 and fp64.
-> tests/golden/solver_grads_mnist.npz    (inputs: tests/golden/mnist_odeblock.npz 'feat', mnist_odeblock_weights.npz)"""
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
from make_golden_cases import MNIST_SOLVER_GRAD_CASES  # noqa: E402

torch.set_num_threads(8)
w = np.load(os.path.join(HERE, "mnist_odeblock_weights.npz"))
feat = np.load(os.path.join(HERE, "mnist_odeblock.npz"))["feat"]
res = {}
for tag, sv in MNIST_SOLVER_GRAD_CASES:
    for dtype, dn in ((torch.float32, "f32"), (torch.float64, "f64")):
        blk = mnist_layers.MetaODEBlock().to(dtype)
        blk.integration_time = blk.integration_time.to(dtype)
        rf = blk.rhs_func
        with torch.no_grad():
            for i in (1, 2, 3):
                getattr(rf, "norm%d" % i).weight.copy_(torch.from_numpy(w["norm%d_w" % i]))
                getattr(rf, "norm%d" % i).bias.copy_(torch.from_numpy(w["norm%d_b" % i]))
            for i in (1, 2):
                getattr(rf, "conv%d" % i)._layer.weight.copy_(torch.from_numpy(w["conv%d_w" % i]))
                getattr(rf, "conv%d" % i)._layer.bias.copy_(torch.from_numpy(w["conv%d_b" % i]))
        solver = create_solver(*sv, dtype, "cpu")
        solver.unfreeze_params()
        x = torch.from_numpy(feat).to(dtype).requires_grad_(True)
        y = blk(x, [solver], Namespace(solver_mode="standalone"))
        r = torch.from_numpy(det_normal(tuple(y.shape), 77)).to(dtype)
        (y * r).sum().backward()
        res["%s_%s_du" % (tag, dn)] = solver.u.grad.numpy().copy()
        if solver.v is not None:
            res["%s_%s_dv" % (tag, dn)] = solver.v.grad.numpy().copy()
        if dn == "f32":
            res["%s_f32_gx" % tag] = x.grad.numpy().copy()
        print(tag, dn, {k: v for k, v in res.items() if k.startswith(tag + "_" + dn + "_d")})
np.savez_compressed(os.path.join(HERE, "solver_grads_mnist.npz"), **res)
