"""Same-box A/B of the mbarrier suspend-time hint and of the two tcgen05 conv forms (debug build)."""
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
for C, HW in ((64, 32), (128, 16)):
    blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
    solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda"); solver.freeze_params()
    x = torch.randn(512, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last)
    for rep in range(2):
      for hint in (0, 200, 2000, 20000):
        lib.msb_debug_suspend_hint(hint)
        for flags, name in ((0, "full"), (14, "epilogue only"), (13, "mma only"), (1, "no-epilogue")):
            lib.msb_debug_conv_flags(flags)
            def run():
                xx = x.clone().requires_grad_(True)
                y = blk(xx, [solver], Namespace(solver_mode="standalone")); y.backward(y)
            run(); torch.cuda.synchronize()
            metasolver_b200.profile_enable(True)
            for _ in range(3): run()
            ms, fl, n = metasolver_b200.profile_read(0)
            metasolver_b200.profile_enable(False)
            print("%s C=%d hint %5d %-16s avg %.1f us" % (os.environ.get("MSB_TC_CONV", "pm"), C, hint, name, 1e3 * ms / n), flush=True)
    lib.msb_debug_conv_flags(0)
