"""Where the weight-gradient kernel's time goes (instrumented library), both geometries (option wgrad_htaps), C = 64."""
import ctypes, os, sys
from argparse import Namespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MSB_LIB_PATH", os.path.join(ROOT, "neural-ode-metasolver_b200", "libmetasolver_b200_dbg.so"))
import torch, torch.nn.functional as F
import metasolver_b200 as msb
from metasolver_b200 import _cabi
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
lib = _cabi.lib()
lib.msb_debug_wgrad_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
C = int(os.environ.get("DIAG_C", "64")); HW = 32 if C == 64 else 16
torch.manual_seed(0)
blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
s = create_solver("rk2", "u", 4, -1, 0.5, -1, torch.float32, "cuda"); s.freeze_params()
x = torch.randn(512, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
for ht in (0, 1, 0, 1):
    msb.set_option("wgrad_htaps", ht)
    for it in range(2):
        if it == 1:
            torch.cuda.synchronize(); lib.msb_debug_wgrad_read(None, 1)
        blk.zero_grad(); x.grad = None
        blk(x, [s], Namespace(solver_mode="standalone")).sum().backward()
    torch.cuda.synchronize()
    b = (ctypes.c_ulonglong * 8)(); lib.msb_debug_wgrad_read(b, 0)
    n = max(int(b[5]), 1)
    print("C=%d wgrad_htaps=%d: per CTA-launch: entry -> loop %.0f clk | MMA loop %.0f (waiting for a full stage %.0f = %.0f%%) | loop end -> exit %.0f | producer waiting for an empty stage %.0f | last MMA issued -> all complete %.0f | final TMEM -> global pass %.0f"
          % (C, ht, b[3] / n, b[1] / n, b[0] / n, 100.0 * b[0] / max(b[1], 1), b[4] / n, b[2] / n, b[6] / n, b[7] / n))
