"""Take the TMEM-resident-weight conv (conv_tct.cu) apart with its `tct_debug` switches (release build; results are
garbage, only timings matter): average conv launch time of one RK2 8-step ODE block at B = 512 per variant, with SM clock
and board power sampled meanwhile.   python scripts/tct_decompose.py [fwd|bwd]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from argparse import Namespace
import pynvml
import metasolver_b200
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity

pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = False
def sampler():
    while not stop:
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.01)
threading.Thread(target=sampler, daemon=True).start()
grad = len(sys.argv) > 1 and sys.argv[1] == "bwd"
torch.manual_seed(0)
C, HW, B = 64, 32, 512
blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()
solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda"); solver.freeze_params()
x = torch.randn(B, C, HW, HW, device="cuda").contiguous(memory_format=torch.channels_last)
def run():
    if grad:
        xx = x.clone().requires_grad_(True)
        y = blk(xx, [solver], Namespace(solver_mode="standalone")); y.backward(y)
    else:
        with torch.no_grad():
            blk(x, [solver], Namespace(solver_mode="standalone"))
metasolver_b200.set_option("tc_form_c64", 2)
for flags, name in ((0, "full"), (1, "no epilogue"), (2, "no MMA"), (8, "no TMA"), (9, "MMA only"), (10, "epilogue only"), (3, "TMA only"),
                    (11, "hand-shakes only"), (0, "full (again)")):
    metasolver_b200.set_option("tct_debug", flags)
    run(); torch.cuda.synchronize()
    t0 = time.time()
    metasolver_b200.profile_enable(True)
    while time.time() - t0 < 1.0:
        run(); torch.cuda.synchronize()
    t1 = time.time()
    ms, fl, n = metasolver_b200.profile_read(0)
    metasolver_b200.profile_enable(False)
    ss = [s for s in samples if t0 + 0.3 <= s[0] <= t1]
    print("tct %s %-18s conv avg %6.1f us | SM %4.0f MHz %4.0f W" % ("fwd+bwd" if grad else "fwd    ", name, 1e3 * ms / max(n, 1),
          sum(s[1] for s in ss) / max(len(ss), 1), sum(s[2] for s in ss) / max(len(ss), 1)), flush=True)
metasolver_b200.set_option("tct_debug", 0)
stop = True
