"""Interleaved A/B of the C = 64 tcgen05 conv forms on one B200, same process: average conv launch time of one RK2
8-step ODE block (forward only / forward + backward, B = 512) per setting.
    python scripts/conv_form_ab.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from argparse import Namespace
import metasolver_b200
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
REPS, ROUNDS = 3, 4
torch.manual_seed(0)
def S(form=2, pf=0, band=0, pdl=1, prod=3):
    return dict(tc_form_c64=form, epi_l2_prefetch=pf, tct_band=band, pdl=pdl, tct_products=prod)
sets = [S(1), S(2), S(2, prod=3), S(2, band=8), S(2, pf=1)]
C, HW = 64, 32
if os.environ.get("CONV_AB_C") == "128":          # CTA-pair kernel: role placement A/B
    C, HW = 128, 16
    sets = [dict(uniform_issue=0), dict(uniform_issue=1), dict(uniform_issue=0), dict(uniform_issue=1)]
elif os.environ.get("CONV_AB_C") == "64wg":
    sets = [dict(uniform_issue=0), dict(uniform_issue=1), dict(uniform_issue=0), dict(uniform_issue=1)]
elif os.environ.get("CONV_AB_C") == "64pm":
    sets = [dict(tc_form_c64=1, mma_warp_high=0), dict(tc_form_c64=1, mma_warp_high=1), dict(tc_form_c64=2)]
blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda"); solver.freeze_params()
x = torch.randn(B, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last)
for grad in (False, True):
    def run():
        if grad:
            xx = x.clone().requires_grad_(True)
            y = blk(xx, [solver], Namespace(solver_mode="standalone")); y.backward(y)
        else:
            with torch.no_grad():
                blk(x, [solver], Namespace(solver_mode="standalone"))
    for _ in range(3): run()
    acc = [[0.0, 0, 0.0, 0, 0.0] for _ in sets]
    for rnd in range(ROUNDS):
        for i, st in enumerate(sets):
            for k, v in st.items():
                metasolver_b200.set_option(k, v)
            run(); torch.cuda.synchronize()
            metasolver_b200.profile_enable(True)
            for _ in range(REPS): run()
            ms, fl, n = metasolver_b200.profile_read(0)
            msw, flw, nw = metasolver_b200.profile_read(1)
            metasolver_b200.profile_enable(False)
            a = acc[i]; a[0] += ms; a[1] += n; a[2] += msw; a[3] += nw; a[4] += fl
    for st, a in zip(sets, acc):
        print("C=%3d %s %-60s conv: %5d launches avg %6.1f us (%5.1f TF/s alg)   wgrad: %5d avg %6.1f us"
              % (C, "fwd+bwd" if grad else "fwd    ", str(st), a[1], 1e3 * a[0] / max(a[1], 1),
                 a[4] / max(a[0], 1e-9) / 1e9, a[3], 1e3 * a[2] / max(a[3], 1)), flush=True)
