"""ncu driver for ONE conv form: forward ODE block (RK2, 1 step = 4 conv launches: two conv1-type and two conv2-type
epilogues) at B images, after one warm-up pass.   python scripts/prof_conv.py <form 0|1|2> <debug flags> [B] [C]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from argparse import Namespace
import metasolver_b200
from metasolver_b200 import _cabi
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity

form, flags = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 512
C = int(sys.argv[4]) if len(sys.argv) > 4 else 64
HW = 32 if C == 64 else 16
metasolver_b200.set_option("tc_form_c64", form)
lib = _cabi.lib()
if flags:
    assert lib.msb_debug_conv_flags(flags) == 0
torch.manual_seed(0)
blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
solver = create_solver("rk2", "u", 1, -1, 0.5, -1, torch.float32, "cuda"); solver.freeze_params()
x = torch.randn(B, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last)
grad = os.environ.get("PROF_GRAD") == "1"
for it in range(2):
    if grad:
        xx = x.clone().requires_grad_(True)
        y = blk(xx, [solver], Namespace(solver_mode="standalone")); y.backward(y)
    else:
        with torch.no_grad():
            y = blk(x, [solver], Namespace(solver_mode="standalone"))
torch.cuda.synchronize()
print("ok")
