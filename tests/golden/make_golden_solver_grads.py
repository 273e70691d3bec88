#!/usr/bin/env python
"""Golden gradients w.r.t. the SOLVER parameters u, v (unfreeze_params(), rk_parametric_order2stage2.py:104-109 and
the order-3 / order-4 analogues) from the REAL reference on the CPU, fp32 and fp64.
Same blocks / inputs / loss as the ODE cases of make_golden.py.   -> tests/golden/solver_grads.npz"""
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

from sopa.src.solvers.utils import create_solver  # noqa: E402
from sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2, BasicBlock2  # noqa: E402
from sopa.src.models.odenet_cifar10.utils import Identity  # noqa: E402
from make_golden_cases import SOLVER_GRAD_CASES, ode_case_inputs  # noqa: E402


def run(C, H, W, B, kind, sv, dtype):
    x, w1, w2, r = [torch.from_numpy(a).to(dtype) for a in ode_case_inputs(C, H, W, B)]
    cls = PreBasicBlock2 if kind == "preact" else BasicBlock2
    blk = MetaODEBlock(cls(C, norm_layer=Identity, act_layer=F.gelu)).to(dtype)
    blk.integration_time = blk.integration_time.to(dtype)
    with torch.no_grad():
        blk.rhs_func.conv1.weight.copy_(w1)
        blk.rhs_func.conv2.weight.copy_(w2)
    solver = create_solver(*sv, dtype, "cpu")
    solver.unfreeze_params()
    x.requires_grad_(True)
    y = blk(x, [solver], Namespace(solver_mode="standalone"))
    (y * r).sum().backward()
    out = dict(y=y.detach().numpy(), gx=x.grad.numpy(), du=solver.u.grad.numpy().copy())
    if solver.v is not None:
        out["dv"] = solver.v.grad.numpy().copy()
    return out


def main():
    torch.set_num_threads(8)
    res = {}
    for name, C, H, W, B, kind, sv in SOLVER_GRAD_CASES:
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            for k, v in run(C, H, W, B, kind, sv, dtype).items():
                if k in ("y", "gx") and tag == "f64":
                    continue
                res["%s_%s_%s" % (name, tag, k)] = v if k not in ("y", "gx") else v.reshape(-1)[::7].copy()
            print(name, tag, {k: v for k, v in res.items() if k.startswith(name + "_" + tag + "_d")})
    np.savez_compressed(os.path.join(HERE, "solver_grads.npz"), **res)


if __name__ == "__main__":
    main()
