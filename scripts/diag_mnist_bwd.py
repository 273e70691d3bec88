import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from argparse import Namespace
import numpy as np, torch
import metasolver_b200 as msb
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_mnist.layers import MetaODEBlock
import oracle
from oracle import det_normal
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
w = np.load(G + "/mnist_odeblock_weights.npz"); g = np.load(G + "/mnist_odeblock.npz")
blk = MetaODEBlock(); rf = blk.rhs_func
with torch.no_grad():
    for i in (1, 2, 3):
        getattr(rf, "norm%d" % i).weight.copy_(torch.from_numpy(w["norm%d_w" % i])); getattr(rf, "norm%d" % i).bias.copy_(torch.from_numpy(w["norm%d_b" % i]))
    for i in (1, 2):
        getattr(rf, "conv%d" % i)._layer.weight.copy_(torch.from_numpy(w["conv%d_w" % i])); getattr(rf, "conv%d" % i)._layer.bias.copy_(torch.from_numpy(w["conv%d_b" % i]))
blk = blk.cuda()
x = torch.from_numpy(g["feat"]).cuda()
for tag, sv, uv in (("rk2_u05_n8", ("rk2", "u", 8, -1, 0.5, -1), ("rk2", "u", np.float32(0.5), None)),
                    ("rk4_u2_n2", ("rk4", "u2", 2, -1, 1 / 3., -1), ("rk4", "u2", np.float32(1 / 3.), None)),
                    ("euler_n4", ("euler", None, 4, -1, -1, -1), ("euler", None, None, None))):
    solver = create_solver(*sv, torch.float32, "cuda"); solver.freeze_params()
    blk.zero_grad()
    xg = x.clone().requires_grad_(True)
    yg = blk(xg, [solver], Namespace(solver_mode="standalone"))
    r = torch.from_numpy(det_normal(tuple(yg.shape), 77)).cuda()
    (yg * r).sum().backward()
    a = xg.grad.cpu().numpy().astype(np.float64); b = g[tag + "_gx"].astype(np.float64)
    err = np.abs(a - b) / np.abs(b).max()
    print(tag, "y err", np.abs(yg.detach().cpu().numpy() - g[tag + "_y"]).max() / np.abs(g[tag + "_y"]).max(),
          "gx max", err.max(), "n>1e-5:", (err > 1e-5).sum(), "n>1e-6:", (err > 1e-6).sum(), "of", err.size,
          "samples with err>1e-5:", np.unique(np.nonzero(err > 1e-5)[0]))
    # fp64 oracle for reference-side noise
    for dt in (torch.float32, torch.float64):
        po = {k: torch.from_numpy(w[k]).to(dt).requires_grad_(True) for k in w.files}
        xo = torch.from_numpy(g["feat"]).to(dt).requires_grad_(True)
        tab = oracle.butcher_tableau(*uv) if dt == torch.float32 else oracle.butcher_tableau(uv[0], uv[1], None if uv[2] is None else np.float64(uv[2]), None, torch.float64)
        yo = oracle.integrate(tab, oracle.rhs_mnist(po), xo, torch.tensor([0, 1]).to(dt), n_steps=sv[2])[-1]
        (yo * r.cpu().to(dt)).sum().backward()
        e2 = np.abs(xo.grad.double().numpy() - b) / np.abs(b).max()
        e3 = np.abs(xo.grad.double().numpy() - a) / np.abs(b).max()
        print("   oracle", dt, "vs golden", e2.max(), " ours vs this oracle", e3.max(), "n>1e-5", (e3 > 1e-5).sum())
