"""Kernel-time breakdown of one bench step with torch.profiler (no replay; cheap)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from argparse import Namespace
from torch.profiler import profile, ProfilerActivity
import metasolver_b200
from metasolver_b200.sopa.src.solvers.utils import create_solver
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import premetanode10
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(602)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
model = premetanode10((Identity,) * 3, (lambda x: x,) * 3, (F.gelu,) * 3, in_planes=64).cuda().to(memory_format=torch.channels_last)
solver = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda"); solver.freeze_params()
opts = Namespace(solver_mode="standalone")
x = torch.randn(B, 3, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 10, (B,), device="cuda")
def step():
    model.zero_grad(set_to_none=True)
    F.cross_entropy(model(x, [solver], opts), y).backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows if e.device_type == torch.autograd.DeviceType.CUDA)
print("total device us per step", tot / 3)
for e in rows[:28]:
    if e.device_type == torch.autograd.DeviceType.CUDA:
        print("%-90s n=%5d  %9.1f us/step  %5.1f%%" % (e.key[:90], e.count // 3, e.device_time_total / 3, 100 * e.device_time_total / tot))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
print("wall ms/step", e0.elapsed_time(e1) / 5)
# which framework-side copies / elementwise ops are left (CPU-side op names with input shapes and their device time)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof2:
    step()
    torch.cuda.synchronize()
print("-- aten ops with device time (one step), by input shape --")
for e in sorted(prof2.key_averages(group_by_input_shape=True), key=lambda e: -e.self_device_time_total)[:40]:
    if e.self_device_time_total > 5 and e.key.startswith("aten::"):
        print("%-28s n=%3d %8.1f us  %s" % (e.key, e.count, e.self_device_time_total, str(e.input_shapes)[:110]))
