"""Diagnostic for the TMEM-resident-weight conv form (conv_tct.cu): one convolution against the fp32 SIMT engine on the
same split operand, with an error map by (image row, pixel, channel) to localise layout mistakes.
    python scripts/diag_tct.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import metasolver_b200
from metasolver_b200 import ops

torch.manual_seed(0)
dev = "cuda"
for (B, H) in ((1, 4), (2, 8), (3, 32), (40, 32)):
    C, W = 64, 32
    x = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.randn(C, C, 3, 3, device=dev) / 24.0
    split, _ = ops.act_split(x)
    for transpose in (False, True):
        ref = ops.conv3x3(split, w, transpose, "simt")
        metasolver_b200.set_option("tc_form_c64", 2)
        out = ops.conv3x3(split, w, transpose, "tcgen05")
        torch.cuda.synchronize()
        err = (out - ref).abs()
        rel = float(err.max() / ref.abs().max())
        print("B=%d H=%d transpose=%d  max rel err %.3e" % (B, H, transpose, rel), flush=True)
        if rel > 1e-5:
            e = err.permute(0, 2, 3, 1)      # (B, H, W, C)
            print("  err by image row:", [round(float(v), 4) for v in e.amax(dim=(0, 2, 3))])
            print("  err by pixel    :", [round(float(v), 4) for v in e.amax(dim=(0, 1, 3))])
            print("  err by channel  :", [round(float(v), 4) for v in e.amax(dim=(0, 1, 2))])
            print("  err by image    :", [round(float(v), 4) for v in e.amax(dim=(1, 2, 3))][:8])
            # single-tap probes: which (tap, c_in parity) is mis-wired
            for tap in range(9):
                w1 = torch.zeros_like(w); w1[:, :, tap // 3, tap % 3] = w[:, :, tap // 3, tap % 3]
                r1 = ops.conv3x3(split, w1, transpose, "simt"); o1 = ops.conv3x3(split, w1, transpose, "tcgen05")
                print("   tap %d: rel err %.3e" % (tap, float((o1 - r1).abs().max() / r1.abs().max())))
            for ci in (0, 1, 2, 17, 63):
                w1 = torch.zeros_like(w); w1[:, ci] = w[:, ci]
                r1 = ops.conv3x3(split, w1, transpose, "simt"); o1 = ops.conv3x3(split, w1, transpose, "tcgen05")
                print("   c_in %d: rel err %.3e" % (ci, float((o1 - r1).abs().max() / r1.abs().max())))
            break
print("done")
