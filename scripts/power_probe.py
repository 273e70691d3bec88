"""Is the convolution power-capped?  For each bottleneck-decomposition variant (library built with
MSB_NVCC_EXTRA=-DMSB_CONV_DEBUG) run the ODE-block forward+backward for ~1.5 s while sampling SM clock and board power
(NVML, 10 ms), and print the average conv launch time next to them."""
import ctypes, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from argparse import Namespace
import pynvml
import metasolver_b200
from metasolver_b200 import _cabi
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
lib = ctypes.CDLL(_cabi.LIB_PATH)
have_dbg = hasattr(lib, "msb_debug_conv_flags")
samples = []
stop = False
def sampler():
    while not stop:
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.01)
th = threading.Thread(target=sampler, daemon=True); th.start()

DUR = float(os.environ.get("PROBE_SECONDS", "1.5"))
torch.manual_seed(0)
variants = [(0, "full")]
if have_dbg:
    variants += [(13, "mma only"), (14, "epilogue only"), (3, "TMA only"), (1, "no epilogue"), (48, "no epi loads/stores"), (16, "no epi stores"), (32, "no epi loads"),
                 (64, "epi loads L2-resident"), (128, "epi stores L2-resident"), (192, "epi ld+st L2-resident"), (0, "full (again)")]
SHAPES = [(64, 32), (128, 16)] if not os.environ.get("PROBE_C") else [(int(os.environ["PROBE_C"]), 32 if os.environ["PROBE_C"] == "64" else 16)]
for C, HW in SHAPES:
    blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
    solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda"); solver.freeze_params()
    x = torch.randn(512, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last)
    for grad in ((True,) if os.environ.get("PROBE_BWD_ONLY") else (False, True)):
        for flags, name in variants:
            if have_dbg: lib.msb_debug_conv_flags(flags)
            def run():
                if grad:
                    xx = x.clone().requires_grad_(True)
                    y = blk(xx, [solver], Namespace(solver_mode="standalone")); y.backward(y)
                else:
                    with torch.no_grad():
                        blk(x, [solver], Namespace(solver_mode="standalone"))
            run(); torch.cuda.synchronize()
            t0 = time.time()
            metasolver_b200.profile_enable(True)
            while time.time() - t0 < DUR:
                run(); torch.cuda.synchronize()
            t1 = time.time()
            ms, fl, n = metasolver_b200.profile_read(0)
            msw, flw, nw = metasolver_b200.profile_read(1)
            metasolver_b200.profile_enable(False)
            ss = [s for s in samples if t0 + 0.3 <= s[0] <= t1]
            clk = sum(s[1] for s in ss) / max(len(ss), 1)
            pw = sum(s[2] for s in ss) / max(len(ss), 1)
            rs = 0
            for s in ss: rs |= s[3]
            busy = (ms + msw) / 1e3 / (t1 - t0)
            print("C=%3d %s %-22s conv avg %6.1f us  wgrad avg %6.1f us | SM %4.0f MHz  %4.0f W  throttle 0x%x  (kernels busy %.0f%%, %d samples)"
                  % (C, "fwd+bwd" if grad else "fwd    ", name, 1e3 * ms / max(n, 1), 1e3 * msw / max(nw, 1), clk, pw, rs, 100 * busy, len(ss)), flush=True)
            time.sleep(0.5)
        if have_dbg: lib.msb_debug_conv_flags(0)
stop = True
