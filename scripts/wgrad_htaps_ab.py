"""Interleaved same-process A/B of the weight-gradient geometry (option wgrad_htaps): a CTA owns a vertical tap and reads the
three horizontal taps from ONE staged copy (N atoms 128 B apart) vs a CTA owns a horizontal tap with its own shifted box.
B = 512, ODE block RK2 4 steps, fwd + bwd, both channel counts."""
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

B = 512
import os
SHAPES = (((64, 32, 32),) if os.environ.get("HT_AB_C64") else ((64, 32, 32), (128, 16, 16))) if not os.environ.get("HT_AB_ALL") else ((64, 32, 32), (64, 64, 16), (128, 16, 16), (128, 8, 32))
for C, H, W in SHAPES:
    torch.manual_seed(0)
    blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
    s = create_solver("rk2", "u", 4, -1, 0.5, -1, torch.float32, "cuda")
    s.freeze_params()
    x = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)

    def step():
        blk.zero_grad()
        x.grad = None
        blk(x, [s], Namespace(solver_mode="standalone")).sum().backward()

    for rep in range(2 if os.environ.get("HT_AB_ALL") else int(os.environ.get("HT_AB_REPS", "3"))):
        for ht in (0, 1):
            msb.set_option("wgrad_htaps", ht)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            msb.profile_enable(True)
            for _ in range(6):
                step()
            ms, fl, n = msb.profile_read(0)
            wms, wfl, wn = msb.profile_read(1)
            msb.profile_enable(False)
            print("C=%d %dx%d wgrad_htaps=%d  conv: %d launches avg %.1f us   wgrad: %d avg %.1f us" % (C, H, W, ht, n, ms / max(n, 1) * 1e3, wn, wms / max(wn, 1) * 1e3), flush=True)
