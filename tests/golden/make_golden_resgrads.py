#!/usr/bin/env python
"""Golden gradients of the NON-ODE layers of premetanode10 (stem is in premetanode10.npz; here: both residual
blocks incl. the strided one with its 1x1 shortcut, and the FC bias), from the REAL reference on the CPU.
Same model / input / loss as section F of make_golden.py.   -> tests/golden/premetanode10_resgrads.npz"""
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
from sopa.src.models.odenet_cifar10.layers import premetanode10  # noqa: E402
from sopa.src.models.odenet_cifar10.utils import Identity  # noqa: E402
from make_golden_cases import WG_STRIDE, conv_w  # noqa: E402

torch.set_num_threads(8)
model = premetanode10((Identity,) * 3, (lambda x: x,) * 3, (F.gelu,) * 3, in_planes=64, is_odenet=True)
new = {}
for i, (k, v) in enumerate(model.state_dict().items()):
    if v.dim() == 4:
        new[k] = torch.from_numpy(conv_w(v.shape[0], v.shape[1], 500 + i, v.shape[2]))
    elif v.dim() == 2:
        bound = 1.0 / np.sqrt(v.shape[1])
        new[k] = torch.from_numpy(det_uniform(tuple(v.shape), 500 + i, -bound, bound))
    else:
        new[k] = torch.from_numpy(det_uniform(tuple(v.shape), 500 + i, -0.1, 0.1))
model.load_state_dict(new)
model.eval()
img = torch.from_numpy(det_uniform((4, 3, 32, 32), 900, 0.0, 1.0))
mean = torch.tensor((0.4914, 0.4822, 0.4465)).view(1, 3, 1, 1)
std = torch.tensor((0.2023, 0.1994, 0.2010)).view(1, 3, 1, 1)
xin = ((img - mean) / std).requires_grad_(True)
solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cpu")
solver.freeze_params()
acts = {}
hooks = [model.conv1.register_forward_hook(lambda m, i, o: acts.__setitem__("stem_preact", o.detach().numpy())),
         model.layer2.blocks_res.register_forward_hook(lambda m, i, o: acts.__setitem__("layer2_res_out", o.detach().numpy()))]
logits = model(xin, [solver], Namespace(solver_mode="standalone"))
F.cross_entropy(logits, torch.tensor([3, 1, 4, 1])).backward()
out = dict(acts)
for k, p in model.named_parameters():
    if "blocks_res" in k or k == "fc_layers.2.bias":
        out["g_" + k] = p.grad.numpy().reshape(-1)[::WG_STRIDE].copy()
np.savez(os.path.join(HERE, "premetanode10_resgrads.npz"), **out)
print({k: v.shape for k, v in out.items()})
