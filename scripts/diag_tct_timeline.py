"""Kernel timeline of the C = 64 TMEM-resident-weight convolution over a real ODE-block forward + backward (instrumented
library: python neural-ode-metasolver_b200/build.py --debug)."""
import ctypes
import os
import sys
from argparse import Namespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MSB_LIB_PATH", os.path.join(ROOT, "neural-ode-metasolver_b200", "libmetasolver_b200_dbg.so"))
import torch
import torch.nn.functional as F
import metasolver_b200  # noqa: F401
from metasolver_b200 import _cabi
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity

lib = _cabi.lib()
lib.msb_debug_tct_time.argtypes = [ctypes.c_void_p, ctypes.c_int]
B, C, HW = int(os.environ.get("DIAG_B", "512")), 64, 32
torch.manual_seed(0)
blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
s = create_solver("rk2", "u", 4, -1, 0.5, -1, torch.float32, "cuda")
s.freeze_params()
x = torch.randn(B, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        lib.msb_debug_tct_time(None, 1)
    blk.zero_grad()
    blk(x, [s], Namespace(solver_mode="standalone")).sum().backward()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 12)()
lib.msb_debug_tct_time(buf, 0)
n = max(int(buf[5]), 1)
t = [buf[i] / n for i in range(5)]
rows = B * HW // 148 + 1
print("conv3x3_tct, %d CTA-launches (%d launches): clocks from kernel entry, mean over CTAs" % (n, n // 148))
print("  past griddepcontrol.wait %.0f | weights in TMEM %.0f | MMA issue loops done %.0f | last epilogue warp done %.0f | exit %.0f" % tuple(t))
print("  prologue + wait %.1f%%  weights %.1f%%  MMA loops %.1f%%  tail epilogue %.1f%%   (MMA floor: ~%d rows x 36 x 32 clk = %d clk = %.1f%% of the CTA)"
      % (100 * t[0] / t[4], 100 * (t[1] - t[0]) / t[4], 100 * (t[2] - t[1]) / t[4], 100 * (t[3] - t[2]) / t[4], rows, rows * 1152, 100.0 * rows * 1152 / t[4]))
w = [buf[6 + i] / n for i in range(4)]
loop = t[2] - t[1]
print("  the two issuing warps together, clocks per CTA: wait accumulator %.0f | wait input rows %.0f | wait turn %.0f | issuing %.0f   (loop %.0f per warp)"
      % (w[0], w[1], w[2], w[3], loop))
