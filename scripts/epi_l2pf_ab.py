"""A/B of the convolution tuning options on one B200, same process: average conv / wgrad launch time of one RK2
8-step ODE block (forward + backward, B = 512) per option setting.  Results never depend on the options
(tests/test_gpu_odeblock.py::test_options_do_not_change_results)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from argparse import Namespace
import metasolver_b200
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
REPS = 3
torch.manual_seed(0)
OPTS = ("epi_l2_prefetch", "tc_resident", "tcp_epi_warps", "tc_form_c64", "tc_pair", "wait_backoff_ns", "wgrad_multicast")
def S(pf=0, res=0, ew=16, pm64=1, pair=2, bo=0, mc=0):
    return dict(epi_l2_prefetch=pf, tc_resident=res, tcp_epi_warps=ew, tc_form_c64=pm64, tc_pair=pair, wait_backoff_ns=bo,
                wgrad_multicast=mc)
settings = {64: [S(), S(mc=1)],
            128: [S(), S(mc=1)]}
ROUNDS = 5
for C, HW in ((64, 32), (128, 16)):
    blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
    solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda"); solver.freeze_params()
    x = torch.randn(B, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last)
    sets = settings[C]
    for grad in (False, True):
        def run():
            if grad:
                xx = x.clone().requires_grad_(True)
                y = blk(xx, [solver], Namespace(solver_mode="standalone")); y.backward(y)
            else:
                with torch.no_grad():
                    blk(x, [solver], Namespace(solver_mode="standalone"))
        for _ in range(3): run()                       # warm up (clocks settle under the power cap)
        acc = [[0.0, 0, 0.0, 0, 0.0] for _ in sets]
        for rnd in range(ROUNDS):                      # interleaved: clock drift hits every setting alike
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
            print("C=%3d %s l2pf=%d resident=%d epi_warps=%2d pm64=%d pair=%d backoff=%3d wgrad_mc=%d  conv: %5d launches avg %6.1f us (%5.1f TF/s alg)   wgrad: %5d avg %6.1f us"
                  % (C, "fwd+bwd" if grad else "fwd    ", st["epi_l2_prefetch"], st["tc_resident"], st["tcp_epi_warps"], st["tc_form_c64"], st["tc_pair"], st["wait_backoff_ns"], st["wgrad_multicast"], a[1], 1e3 * a[0] / max(a[1], 1),
                     a[4] / max(a[0], 1e-9) / 1e9, a[3], 1e3 * a[2] / max(a[3], 1)), flush=True)
