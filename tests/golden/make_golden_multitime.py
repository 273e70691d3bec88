#!/usr/bin/env python
"""Golden vectors for integrate() with INTERIOR output times (linear interpolation between grid points,
rk_parametric.py:104-123) from the REAL reference on the CPU.   -> tests/golden/multitime.npz"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

from sopa.src.solvers.utils import create_solver  # noqa: E402
from sopa.src.models.odenet_cifar10.layers import PreBasicBlock2  # noqa: E402
from sopa.src.models.odenet_cifar10.utils import Identity  # noqa: E402
from make_golden_cases import MULTITIME_CASES, WG_STRIDE, ode_case_inputs  # noqa: E402

torch.set_num_threads(8)
res = {}
for name, C, H, W, B, sv, times in MULTITIME_CASES:
    x, w1, w2, r = [torch.from_numpy(a) for a in ode_case_inputs(C, H, W, B)]
    rhs = PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)
    with torch.no_grad():
        rhs.conv1.weight.copy_(w1)
        rhs.conv2.weight.copy_(w2)
    solver = create_solver(*sv, torch.float32, "cpu")
    solver.freeze_params()
    x.requires_grad_(True)
    ys = solver.integrate(rhs, x, torch.tensor(times))
    loss = sum(((k + 1.0) * ys[k] * r).sum() for k in range(1, len(times)))
    loss.backward()
    res[name + "_y"] = ys.detach().numpy()[1:, :, ::3].copy()
    res[name + "_gx"] = x.grad.numpy()
    res[name + "_gw1"] = rhs.conv1.weight.grad.numpy().reshape(-1)[::WG_STRIDE].copy()
    res[name + "_nfe"] = np.int64(rhs.nfe)
    print(name, ys.shape, rhs.nfe)
np.savez_compressed(os.path.join(HERE, "multitime.npz"), **res)
