"""Times the stem forward / backward kernels (B = 512, 32x32, 3 -> 64) through the layer API."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import metasolver_b200 as msb
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import premetanode10
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
model = premetanode10((Identity,) * 3, (lambda t: t,) * 3, (F.gelu,) * 3, in_planes=64, is_odenet=True).cuda().to(memory_format=torch.channels_last)
x = torch.randn(512, 3, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last)
from torch.profiler import profile, ProfilerActivity
from argparse import Namespace
from metasolver_b200.sopa.src.solvers.utils import create_solver
s = create_solver("rk2", "u", 2, -1, 0.5, -1, torch.float32, "cuda"); s.freeze_params()
opts = Namespace(solver_mode="standalone")
for _ in range(2):
    model.zero_grad(); model(x, [s], opts).sum().backward()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        model.zero_grad(); model(x, [s], opts).sum().backward()
    torch.cuda.synchronize()
for e in prof.key_averages():
    if "stem" in e.key or "act_split" in e.key or "s2d" in e.key or "d2s" in e.key or "pool_fc" in e.key:
        print("%-60s n=%3d avg %.1f us" % (e.key[:60], e.count, e.device_time_total / max(e.count, 1)))
