#!/usr/bin/env python
"""Times the CIFAR ODE block with a per-sample normalisation inside the right-hand side (GN / LN / IN; premetanode10 shapes,
RK2 8 steps, forward + backward) next to the normalisation-free block.  One JSON line."""
import json
import os
import sys
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.nn.functional as F
    import metasolver_b200  # noqa: F401
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import get_normalization
    B = int(os.environ.get("GN_BENCH_BATCH", "256"))
    out = {}
    for C, H in ((64, 32), (128, 16)):
        x = torch.randn(B, C, H, H, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
        s = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cuda")
        s.freeze_params()
        for key in ("NF", "GN", "LN", "IN"):
            blk = MetaODEBlock(PreBasicBlock2(C, norm_layer=get_normalization(key, 32), act_layer=F.gelu)).cuda()

            def step():
                blk.zero_grad()
                x.grad = None
                blk(x, [s], Namespace(solver_mode="standalone")).sum().backward()
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                step()
            e1.record()
            torch.cuda.synchronize()
            out["C%d_%s_ms" % (C, key)] = e0.elapsed_time(e1) / 3
    out["batch"] = B
    out["gn_block"] = metasolver_b200.get_option("gn_block") if hasattr(metasolver_b200, "get_option") else None
    print(json.dumps(out))


if __name__ == "__main__":
    main()
