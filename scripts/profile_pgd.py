"""Kernel-time breakdown of one PGD-7 evaluation batch (B=512) with torch.profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from argparse import Namespace
from torch.profiler import profile, ProfilerActivity
import metasolver_b200
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import premetanode10
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
from metasolver_b200.MegaAdversarial.src.attacks import PGD
MEAN, STD = (0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)
torch.manual_seed(0)
model = premetanode10((Identity,) * 3, (lambda x: x,) * 3, (F.gelu,) * 3, in_planes=64).cuda().to(memory_format=torch.channels_last).eval()
solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda"); solver.freeze_params()
kw = {"solvers": [solver], "solver_options": Namespace(solver_mode="standalone")}
x = torch.randn(512, 3, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 10, (512,), device="cuda")
pgd = PGD(model, eps=8 / 255., lr=2 / 255., n_iter=7, mean=MEAN, std=STD)
def run():
    xa, _ = pgd(x, y, kw)
    with torch.no_grad():
        return (model(xa, **kw).argmax(1) == y).sum()
run(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows if e.device_type == torch.autograd.DeviceType.CUDA)
print("total device ms", tot / 1e3)
for e in rows[:22]:
    if e.device_type == torch.autograd.DeviceType.CUDA:
        print("%-80s n=%5d  %9.1f us  %5.1f%%" % (e.key[:80], e.count, e.device_time_total, 100 * e.device_time_total / tot))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("wall ms", e0.elapsed_time(e1))
