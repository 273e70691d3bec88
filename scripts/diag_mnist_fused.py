"""Fused single-launch MNIST solve vs the multi-launch SIMT path: error statistics per solver (outputs, per-sample input
gradients, parameter gradients).   python scripts/diag_mnist_fused.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from argparse import Namespace
import metasolver_b200 as msb
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_mnist.layers import MetaODEBlock
torch.manual_seed(4)
blk = MetaODEBlock().cuda()
with torch.no_grad():
    for prm in blk.parameters():
        if prm.dim() == 1:
            prm.add_(0.1 * torch.randn_like(prm))
for sv, B in ((("rk2", "u", 8, -1, 0.5, -1), 128), (("rk4", "uv", 3, -1, 1 / 3., 2 / 3.), 5), (("rk3", "uv", 2, -1, 0.4, 0.7), 300),
              (("euler", None, 5, -1, -1, -1), 1)):
    solver = create_solver(*sv, torch.float32, "cuda"); solver.freeze_params()
    x = torch.randn(B, 64, 6, 6, device="cuda"); r = torch.randn(B, 64, 6, 6, device="cuda")
    res = {}
    for fused in (1, 0):
        msb.set_option("mnist_fused", fused)
        xg = x.clone().requires_grad_(True); blk.zero_grad()
        y = blk(xg, [solver], Namespace(solver_mode="standalone")); (y * r).sum().backward()
        res[fused] = [y.detach().double().cpu().numpy(), xg.grad.double().cpu().numpy()] + [p.grad.double().cpu().numpy() for p in blk.parameters()]
    a, b = res[1], res[0]
    per = np.abs(a[1] - b[1]).reshape(B, -1).max(1) / np.abs(b[1]).max()
    print(sv[0], "B=%d" % B, "y %.2e" % (np.abs(a[0] - b[0]).max() / np.abs(b[0]).max()),
          "| gx per-sample: median %.2e, frac<=5e-5 %.2f, max %.2e" % (np.median(per), (per <= 5e-5).mean(), per.max()),
          "| params", " ".join("%.1e" % (np.abs(p - q).max() / max(np.abs(q).max(), 1e-30)) for p, q in zip(a[2:], b[2:])), flush=True)
msb.set_option("mnist_fused", 1)
