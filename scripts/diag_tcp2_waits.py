"""Where the MMA-issuing warp of the C = 128 CTA-pair convolution waits (instrumented library:
python neural-ode-metasolver_b200/build.py --debug; MSB_LIB_PATH=.../libmetasolver_b200_dbg.so)."""
import ctypes
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MSB_LIB_PATH", os.path.join(ROOT, "neural-ode-metasolver_b200", "libmetasolver_b200_dbg.so"))
import torch
import metasolver_b200  # noqa: F401
from metasolver_b200 import ops, _cabi

lib = _cabi.lib()
lib.msb_debug_tcp2_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
B, C, H, W = int(os.environ.get("DIAG_B", "512")), 128, 16, 16
torch.manual_seed(0)
x = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
w = torch.randn(C, C, 3, 3, device="cuda") / 34.0
split, _ = ops.act_split(x)
for _ in range(3):
    out = ops.conv3x3(split, w, False, "tcgen05")
torch.cuda.synchronize()
lib.msb_debug_tcp2_read(None, 1)
N = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    out = ops.conv3x3(split, w, False, "tcgen05")
e1.record()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 8)()
lib.msb_debug_tcp2_read(buf, 0)
n = max(int(buf[4]), 1)
tot = buf[3] / n
print("conv128 pair (plain epilogue, B=%d): %.1f us per launch (instrumented build)" % (B, e0.elapsed_time(e1) / N * 1e3))
print("MMA warp, clocks per leader per launch: loop %.0f | waits: weights %.0f (%.1f%%)  activations %.0f (%.1f%%)  accumulator %.0f (%.1f%%)"
      % (tot, buf[0] / n, 100 * buf[0] / buf[3], buf[1] / n, 100 * buf[1] / buf[3], buf[2] / n, 100 * buf[2] / buf[3]))
print("timeline per leader: entry -> MMA loop start %.0f clk | MMA loop %.0f | MMA loop end -> last epilogue warp done %.0f | entry -> exit %.0f"
      % (buf[5] / n, tot, buf[6] / n, buf[7] / n))
tiles = B * H * W // 256
per_leader = -(-tiles // 74)
print("MMA floor: %d tile pairs per leader x 72 k-steps x 192 clk = %d clk (%.1f%% of the loop)" % (per_leader, per_leader * 72 * 192, 100.0 * per_leader * 72 * 192 / tot))

# the same counters over a real ODE-block forward + backward (fused RK epilogues, dgrad launches)
import torch.nn.functional as F
from argparse import Namespace
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
s = create_solver("rk2", "u", 4, -1, 0.5, -1, torch.float32, "cuda")
s.freeze_params()
xx = x.clone().requires_grad_(True)
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        lib.msb_debug_tcp2_read(None, 1)
    blk.zero_grad()
    blk(xx, [s], Namespace(solver_mode="standalone")).sum().backward()
torch.cuda.synchronize()
lib.msb_debug_tcp2_read(buf, 0)
launches = int(buf[4]) // 74
print("timeline per leader: entry -> MMA loop start %.0f clk | MMA loop end -> last epilogue warp done %.0f | entry -> exit %.0f"
      % (buf[5] / max(int(buf[4]), 1), buf[6] / max(int(buf[4]), 1), buf[7] / max(int(buf[4]), 1)))
print("ODE block fwd+bwd (rk2, 4 steps): %d conv launches; MMA-warp loop %.0f clk per launch | waits: weights %.1f%%  activations %.1f%%  accumulator %.1f%% | MMA floor %.1f%%"
      % (launches, buf[3] / max(int(buf[4]), 1), 100 * buf[0] / buf[3], 100 * buf[1] / buf[3], 100 * buf[2] / buf[3],
         100.0 * per_leader * 72 * 192 / (buf[3] / max(int(buf[4]), 1))))
