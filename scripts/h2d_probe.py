"""Times a pinned host -> device copy of the bench's batch (6.3 MB) and a CUDA-graph launch gap, to explain e2e vs value."""
import torch
x = torch.rand(512, 3, 32, 32).contiguous(memory_format=torch.channels_last).pin_memory()
d = torch.empty_strided(x.shape, x.stride(), device="cuda")
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    d.copy_(x, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("H2D %.1f MB pinned: %.3f ms  (%.1f GB/s)" % (x.numel() * 4 / 1e6, ms, x.numel() * 4 / ms / 1e6))
