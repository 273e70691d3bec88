#!/usr/bin/env python
"""Golden vectors for BASELINE config 3 (solver ensembling + model ensembling over RK2 u values and RK4),
produced by running the REAL reference (/root/reference) on the CPU.

    python tests/golden/make_golden_c3.py        -> tests/golden/ensemble_c3.npz

Authoring container only (the reference does not travel to the GPU box).  Inputs/weights are regenerated
from (shape, seed) by oracle.detrand; the fixture holds outputs only.
"""
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
from sopa.src.models.odenet_cifar10.layers import premetanode10, MetaODEBlock, PreBasicBlock2  # noqa: E402
from sopa.src.models.odenet_cifar10.utils import Identity  # noqa: E402

from make_golden_cases import (WG_STRIDE, C3_RK2_SOLVERS, C3_RK4_SOLVERS, C3_WEIGHTS, conv_w,  # noqa: E402
                                ode_case_inputs)


def block_case(C, H, W, B, svs, weights):
    x, w1, w2, r = [torch.from_numpy(a) for a in ode_case_inputs(C, H, W, B)]
    blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu))
    with torch.no_grad():
        blk.rhs_func.conv1.weight.copy_(w1)
        blk.rhs_func.conv2.weight.copy_(w2)
    solvers = [create_solver(*sv, torch.float32, "cpu") for sv in svs]
    for s in solvers:
        s.freeze_params()
    x.requires_grad_(True)
    y = blk(x, solvers, Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=weights))
    (y * r).sum().backward()
    return dict(y=y.detach().numpy(), gx=x.grad.numpy(),
                gw1=blk.rhs_func.conv1.weight.grad.numpy().reshape(-1)[::WG_STRIDE].copy(),
                gw2=blk.rhs_func.conv2.weight.grad.numpy().reshape(-1)[::WG_STRIDE].copy(),
                nfe=np.int64(blk.rhs_func.nfe))


def main():
    torch.set_num_threads(8)
    out = {}
    for tag, args in (("c64_rk2x4_uniform", (64, 8, 32, 2, C3_RK2_SOLVERS, None)),
                      ("c64_rk2x4_weighted", (64, 8, 32, 2, C3_RK2_SOLVERS, C3_WEIGHTS)),
                      ("c64_rk4x2_uniform", (64, 8, 32, 2, C3_RK4_SOLVERS, None)),
                      ("c128_rk2x4_uniform", (128, 8, 16, 1, C3_RK2_SOLVERS, None))):
        for k, v in block_case(*args).items():
            out["%s_%s" % (tag, k)] = v
        print(tag, "done")

    # whole premetanode10, solver ensembling in every ODE block; then model ensembling (FGSM2Ensemble)
    model = premetanode10((Identity,) * 3, (lambda x: x,) * 3, (F.gelu,) * 3, in_planes=64, is_odenet=True)
    sd = model.state_dict()
    new = {}
    for i, (k, v) in enumerate(sd.items()):
        if v.dim() == 4:
            new[k] = torch.from_numpy(conv_w(v.shape[0], v.shape[1], 500 + i, v.shape[2]))
        elif v.dim() == 2:
            bound = 1.0 / np.sqrt(v.shape[1])
            new[k] = torch.from_numpy(det_uniform(tuple(v.shape), 500 + i, -bound, bound))
        else:
            new[k] = torch.from_numpy(det_uniform(tuple(v.shape), 500 + i, -0.1, 0.1))
    model.load_state_dict(new)
    model.eval()
    mean_t, std_t = (0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)
    mean = torch.tensor(mean_t).view(1, 3, 1, 1)
    std = torch.tensor(std_t).view(1, 3, 1, 1)
    img = torch.from_numpy(det_uniform((4, 3, 32, 32), 920, 0.0, 1.0))
    labels = torch.tensor([3, 1, 4, 1])
    solvers = [create_solver(*sv, torch.float32, "cpu") for sv in C3_RK2_SOLVERS]
    for s in solvers:
        s.freeze_params()
    xin = ((img - mean) / std).requires_grad_(True)
    logits = model(xin, solvers, Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=None))
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    out["model_solver_ens_logits"] = logits.detach().numpy()
    out["model_solver_ens_gx"] = xin.grad.numpy()
    for k, p in model.named_parameters():
        if "rhs_func" in k:
            out["model_solver_ens_g_" + k] = p.grad.numpy().reshape(-1)[::WG_STRIDE].copy()
    model.zero_grad()

    from MegaAdversarial.src.attacks import FGSM2Ensemble
    import MegaAdversarial.src.attacks.attack as _att
    _att.device = torch.device("cpu")
    xin = ((img - mean) / std)
    kwargs_arr = [{"solvers": [s], "solver_options": Namespace(solver_mode="standalone")} for s in solvers]
    with torch.no_grad():
        probs = 0
        for kw in kwargs_arr:
            probs = probs + torch.softmax(model(xin, **kw), dim=1)
        out["model_ens_probs"] = (probs / len(kwargs_arr)).numpy()
    x_adv, _ = FGSM2Ensemble([model] * len(solvers), eps=8 / 255., mean=mean_t, std=std_t)(xin, labels, kwargs_arr)
    out["model_ens_fgsm_x"] = x_adv.numpy()
    np.savez(os.path.join(HERE, "ensemble_c3.npz"), **out)
    print("bytes", os.path.getsize(os.path.join(HERE, "ensemble_c3.npz")))


if __name__ == "__main__":
    main()
