"""Interleaved same-process A/B of the C = 128 pair convolution: full-size epilogue stage + 2 activation stages vs
half-size stage + 3 activation stages (option tcp2_half_stage).  B = 512, ODE block RK2 4 steps, fwd + bwd."""
import os
import sys
from argparse import Namespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import metasolver_b200 as msb
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity

C, HW, B = 128, 16, 512
torch.manual_seed(0)
blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
s = create_solver("rk2", "u", 4, -1, 0.5, -1, torch.float32, "cuda")
s.freeze_params()
x = torch.randn(B, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)


def step():
    blk.zero_grad()
    x.grad = None
    blk(x, [s], Namespace(solver_mode="standalone")).sum().backward()


import itertools
for rep in range(3):
    for halo, hs in itertools.product((0, 1), (0, 1)):
        msb.set_option("tcp2_halo", halo)
        msb.set_option("tcp2_half_stage", hs)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        msb.profile_enable(True)
        for _ in range(6):
            step()
        ms, fl, n = msb.profile_read(0)
        wms, wfl, wn = msb.profile_read(1)
        msb.profile_enable(False)
        print("halo=%d half_stage=%d  conv: %d launches avg %.1f us   wgrad: %d avg %.1f us" % (halo, hs, n, ms / max(n, 1) * 1e3, wn, wms / max(wn, 1) * 1e3), flush=True)
