"""Bottleneck decomposition of conv3x3_tc: time the ODE-block forward (inference, B=512) with parts of the kernel
switched off (library built with MSB_NVCC_EXTRA=-DMSB_CONV_DEBUG).  Results are only timings -- outputs are garbage."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from argparse import Namespace
import metasolver_b200
from metasolver_b200 import _cabi
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity

lib = ctypes.CDLL(_cabi.LIB_PATH)
torch.manual_seed(0)
for C, HW in ((64, 32),):
    blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
    solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda"); solver.freeze_params()
    x = torch.randn(512, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last)
    for grad in (False, True):
        for flags, name in ((0, "full"), (1, "no-epilogue"), (2, "no-mma"), (3, "no-epi,no-mma (TMA only)"), (4, "no-weight-tma"),
                            (8, "no-act-tma"), (12, "no-tma"), (13, "mma only"), (14, "epilogue only"),
                            (16, "no epilogue stores"), (32, "no epilogue loads"), (48, "no epilogue loads/stores"),
                            (14 + 16, "epilogue only, no stores"), (14 + 32, "epilogue only, no loads")):
            lib.msb_debug_conv_flags(flags)
            def run():
                if grad:
                    xx = x.clone().requires_grad_(True)
                    y = blk(xx, [solver], Namespace(solver_mode="standalone")); y.backward(y)
                else:
                    with torch.no_grad():
                        blk(x, [solver], Namespace(solver_mode="standalone"))
            run(); torch.cuda.synchronize()
            metasolver_b200.profile_enable(True)
            for _ in range(3): run()
            ms, fl, n = metasolver_b200.profile_read(0)
            metasolver_b200.profile_enable(False)
            print("C=%d %s %-28s conv launches %4d avg %.1f us" % (C, "fwd+bwd" if grad else "fwd    ", name, n, 1e3 * ms / n), flush=True)
        lib.msb_debug_conv_flags(0)
