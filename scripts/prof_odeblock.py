"""Small driver for ncu: one ODE block (RK2, 2 steps) forward+backward at bench shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from argparse import Namespace
import metasolver_b200
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity

C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
HW = 32 if C == 64 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
torch.manual_seed(0)
blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
solver = create_solver("rk2", "u", 2, -1, 0.5, -1, torch.float32, "cuda"); solver.freeze_params()
x = torch.randn(B, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
for it in range(2):
    y = blk(x, [solver], Namespace(solver_mode="standalone"))
    y.square().mean().backward()
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
