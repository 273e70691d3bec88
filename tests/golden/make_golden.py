#!/usr/bin/env python
"""Generate golden vectors by running the REAL reference (/root/reference) on the CPU.

Run in the authoring container only (the reference is Python and does not travel to the GPU
box):   python tests/golden/make_golden.py
Writes small .npz fixtures next to this file.  Inputs/weights are NOT stored: they are
regenerated bit-identically from (shape, seed) by oracle.detrand, so fixtures hold outputs only
(weight gradients are stored strided, see WG_STRIDE).

Nothing in tests/, bench.py or smoke() reads /root/reference at run time; they read these files.
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

from oracle.detrand import det_uniform, det_normal  # noqa: E402
from sopa.src.solvers.utils import create_solver  # noqa: E402
from sopa.src.models.odenet_cifar10.layers import (premetanode10, MetaODEBlock, PreBasicBlock2,  # noqa: E402
                                                   BasicBlock2)
from sopa.src.models.odenet_cifar10.utils import Identity  # noqa: E402
import sopa.src.models.odenet_mnist.layers as mnist_layers  # noqa: E402

from make_golden_cases import (WG_STRIDE, TABLEAU_CASES, ODE_CASES, REGIME_SOLVERS, conv_w,  # noqa: E402
                                ode_case_inputs)


def run_ode_case(C, H, W, B, kind, sv):
    x, w1, w2, r = [torch.from_numpy(a) for a in ode_case_inputs(C, H, W, B)]
    cls = PreBasicBlock2 if kind == "preact" else BasicBlock2
    blk = MetaODEBlock(cls(C, norm_layer=Identity, act_layer=F.gelu))
    with torch.no_grad():
        blk.rhs_func.conv1.weight.copy_(w1)
        blk.rhs_func.conv2.weight.copy_(w2)
    solver = create_solver(*sv, torch.float32, "cpu")
    solver.freeze_params()
    x.requires_grad_(True)
    y = blk(x, [solver], Namespace(solver_mode="standalone"))
    (y * r).sum().backward()
    return dict(y=y.detach().numpy(), gx=x.grad.numpy(),
                gw1=blk.rhs_func.conv1.weight.grad.numpy().reshape(-1)[::WG_STRIDE].copy(),
                gw2=blk.rhs_func.conv2.weight.grad.numpy().reshape(-1)[::WG_STRIDE].copy(),
                nfe=np.int64(blk.rhs_func.nfe))


def main():
    torch.set_num_threads(8)
    # ---- A. tableaus (fp32 and fp64)
    tab = {}
    for i, (m, p, u0, v0) in enumerate(TABLEAU_CASES):
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            s = create_solver(m, p, 4, -1, u0, v0, dt, "cpu")
            c, w, b = s.build_ButcherTableau(return_tableau=True)
            n = len(c)
            wm = np.zeros((n, n))
            for a in range(n):
                for bb in range(len(w[a])):
                    wm[a, bb] = float(w[a][bb])
            tab["%d_%s_c" % (i, tag)] = c.double().numpy()
            tab["%d_%s_b" % (i, tag)] = b.double().numpy()
            tab["%d_%s_w" % (i, tag)] = wm
    np.savez(os.path.join(HERE, "tableaus.npz"), **tab)

    # ---- A2. time grids (linspace quirks, step_size grid)
    grids = {}
    t01 = torch.tensor([0, 1]).float()
    for n in (1, 2, 3, 5, 7, 8, 10, 16):
        s = create_solver("rk2", "u", n, -1, 0.5, -1, torch.float32, "cpu")
        grids["n%d" % n] = s.grid_constructor(t01).numpy()
    for ss in (0.3, 0.125, 0.4):
        s = create_solver("rk2", "u", -1, ss, 0.5, -1, torch.float32, "cpu")
        grids["ss%g" % ss] = s.grid_constructor(t01).numpy()
    np.savez(os.path.join(HERE, "grids.npz"), **grids)

    # ---- B/C. ODE-block forward + gradients
    for name, C, H, W, B, kind, sv in ODE_CASES:
        out = run_ode_case(C, H, W, B, kind, sv)
        np.savez(os.path.join(HERE, "ode_%s.npz" % name), **out)
        print(name, float(np.abs(out["y"]).max()), int(out["nfe"]))

    # ---- D. MNIST: trained ODE-block weights shipped with the reference
    ckpt = "/root/reference/examples/mnist/checkpoints/checkpoint_15444.pth"
    sys.modules.setdefault("sopa.src.models.odenet_mnist.layers", mnist_layers)
    model = torch.load(ckpt, map_location="cpu", weights_only=False).float().eval()
    blk = model.blocks[0]
    rf = blk.rhs_func
    params = dict(norm1_w=rf.norm1.weight, norm1_b=rf.norm1.bias, norm2_w=rf.norm2.weight, norm2_b=rf.norm2.bias,
                  norm3_w=rf.norm3.weight, norm3_b=rf.norm3.bias, conv1_w=rf.conv1._layer.weight,
                  conv1_b=rf.conv1._layer.bias, conv2_w=rf.conv2._layer.weight, conv2_b=rf.conv2._layer.bias)
    np.savez(os.path.join(HERE, "mnist_odeblock_weights.npz"),
             **{k: v.detach().numpy() for k, v in params.items()})
    torch.manual_seed(0)
    img = torch.rand(128, 1, 28, 28)
    with torch.no_grad():
        feat = model.downsampling_layers(img)[:8].contiguous()     # (8,64,6,6) trained-stem features
    mn = dict(feat=feat.numpy())
    for tag, sv in (("rk2_u05_n8", ("rk2", "u", 8, -1, 0.5, -1)), ("rk4_u2_n2", ("rk4", "u2", 2, -1, 1 / 3., -1)),
                    ("euler_n4", ("euler", None, 4, -1, -1, -1))):
        solver = create_solver(*sv, torch.float32, "cpu")
        solver.freeze_params()
        xf = feat.clone().requires_grad_(True)
        model.zero_grad()
        y = blk(xf, [solver], Namespace(solver_mode="standalone"))
        r = torch.from_numpy(det_normal(tuple(y.shape), 77))
        (y * r).sum().backward()
        mn[tag + "_y"] = y.detach().numpy()
        mn[tag + "_gx"] = xf.grad.numpy()
        mn[tag + "_gconv1_w"] = rf.conv1._layer.weight.grad.numpy().reshape(-1)[::WG_STRIDE].copy()
        mn[tag + "_gconv2_b"] = rf.conv2._layer.bias.grad.numpy().copy()
        mn[tag + "_gnorm1_w"] = rf.norm1.weight.grad.numpy().copy()
        mn[tag + "_gnorm3_b"] = rf.norm3.bias.grad.numpy().copy()
    solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cpu")
    solver.freeze_params()
    with torch.no_grad():
        logits = model(img, [solver], Namespace(solver_mode="standalone"))
    mn["logits_first4"] = logits[0, :4].numpy()                     # SURVEY 8(c) pin (3)
    np.savez(os.path.join(HERE, "mnist_odeblock.npz"), **mn)
    print("mnist logits[0,:4]", logits[0, :4])

    # ---- E. regimes (switch / solver-ensemble), same host RNGs as the reference
    x, w1, w2, r = [torch.from_numpy(a) for a in ode_case_inputs(64, 8, 32, 2)]
    blk = MetaODEBlock(PreBasicBlock2(64, norm_layer=Identity, act_layer=F.gelu))
    with torch.no_grad():
        blk.rhs_func.conv1.weight.copy_(w1)
        blk.rhs_func.conv2.weight.copy_(w2)
    svs = REGIME_SOLVERS
    solvers = [create_solver(*sv, torch.float32, "cpu") for sv in svs]
    for s in solvers:
        s.freeze_params()
    reg = {}
    with torch.no_grad():
        np.random.seed(123)
        ids = []
        for rep in range(3):
            opts = Namespace(solver_mode="switch", switch_probs=[0.1, 0.2, 0.3, 0.4])
            reg["switch_y%d" % rep] = blk(x, solvers, opts).numpy()
            ids.append(opts.switch_solver_id)
        reg["switch_ids"] = np.array(ids)
        torch.manual_seed(5)
        opts = Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=[0.4, 0.3, 0.2, 0.1])
        reg["ens_weighted_y"] = blk(x, solvers, opts).numpy()
        opts = Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=None)
        reg["ens_uniform_y"] = blk(x, solvers, opts).numpy()
        opts = Namespace(solver_mode="ensemble", ensemble_prob=0.0, ensemble_weights=None)
        reg["ens_tails_y"] = blk(x, solvers, opts).numpy()
    np.savez(os.path.join(HERE, "regimes.npz"), **reg)

    # ---- F. whole premetanode10 (published config: NF + GeLU, in_planes 64), det weights
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
    img = torch.from_numpy(det_uniform((4, 3, 32, 32), 900, 0.0, 1.0))
    mean = torch.tensor((0.4914, 0.4822, 0.4465)).view(1, 3, 1, 1)
    std = torch.tensor((0.2023, 0.1994, 0.2010)).view(1, 3, 1, 1)
    xin = ((img - mean) / std).requires_grad_(True)
    labels = torch.tensor([3, 1, 4, 1])
    solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cpu")
    solver.freeze_params()
    outs = {}
    hooks = []
    for n, m in model.named_modules():
        if isinstance(m, MetaODEBlock):
            hooks.append(m.register_forward_hook(lambda mod, i, o, n=n: outs.__setitem__(n, o.detach().numpy())))
    logits = model(xin, [solver], Namespace(solver_mode="standalone"))
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    full = dict(logits=logits.detach().numpy(), loss=np.float64(loss.item()), gx=xin.grad.numpy(),
                keys=np.array(list(sd.keys())))
    for n, o in outs.items():
        full["odeblock_" + n] = o
    for k, p in model.named_parameters():
        if "rhs_func" in k or k in ("conv1.weight", "fc_layers.2.weight"):
            full["g_" + k] = p.grad.numpy().reshape(-1)[::WG_STRIDE].copy()
    np.savez(os.path.join(HERE, "premetanode10.npz"), **full)
    print("premetanode10 logits", logits[0])
    # ---- G. attacks on the whole model (callers of the input-gradient backward) + FGSM-random train step
    from MegaAdversarial.src.attacks import FGSM, FGSMRandom, PGD
    import MegaAdversarial.src.attacks.attack as _att
    _att.device = torch.device("cpu")
    nb = 8
    img = torch.from_numpy(det_uniform((nb, 3, 32, 32), 910, 0.0, 1.0))
    xin = (img - mean) / std
    labels = torch.tensor([3, 1, 4, 1, 5, 9, 2, 6])
    kw = {"solvers": [solver], "solver_options": Namespace(solver_mode="standalone")}
    mean_t, std_t = (0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)
    att = {}
    model.eval()
    with torch.no_grad():
        att["clean_logits"] = model(xin, **kw).numpy()
    x_f, _ = FGSM(model, eps=8 / 255., mean=mean_t, std=std_t)(xin, labels, kw)
    att["fgsm_x"] = x_f.numpy()
    torch.manual_seed(77)
    start = torch.zeros(nb, 3, 32, 32).uniform_(-8 / 255., 8 / 255.)
    torch.manual_seed(77)
    x_p, _ = PGD(model, eps=8 / 255., lr=2 / 255., n_iter=7, mean=mean_t, std=std_t)(xin, labels, kw)
    att["pgd_start"] = start.numpy()
    att["pgd_x"] = x_p.numpy()
    with torch.no_grad():
        att["fgsm_logits"] = model(x_f, **kw).numpy()
        att["pgd_logits"] = model(x_p, **kw).numpy()
    # FGSM-random training step (train_and_attack.py:246-327): zero_grad -> attack (backward #1) -> train pass
    model.train()
    model.zero_grad()
    torch.manual_seed(78)
    u01 = torch.rand(nb, 3, 32, 32)
    torch.manual_seed(78)
    x_r, _ = FGSMRandom(model, alpha=10 / 255., epsilon=8 / 255., mu=mean_t, std=std_t)(xin, labels, kw)
    loss = F.cross_entropy(model(x_r, **kw), labels)
    loss.backward()
    att["fgsmr_u01"] = u01.numpy()
    att["fgsmr_x"] = x_r.numpy()
    att["train_loss"] = np.float64(loss.item())
    for k, prm in model.named_parameters():
        if "rhs_func" in k or k in ("conv1.weight", "fc_layers.2.weight"):
            att["train_g_" + k] = prm.grad.numpy().reshape(-1)[::WG_STRIDE].copy()
    np.savez(os.path.join(HERE, "attacks.npz"), **att)
    print("attacks: clean pred", att["clean_logits"].argmax(1), "pgd pred", att["pgd_logits"].argmax(1))
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print("fixture bytes:", tot)


if __name__ == "__main__":
    main()
