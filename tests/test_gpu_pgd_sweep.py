"""GPU parity of BASELINE config 5 (PGD-7 robust-accuracy sweep over the RK2 u grid) against goldens from the REAL
reference (tests/golden/make_golden_pgd_sweep.py, 512 images x 4 u values x 2 attack strengths): the sharded evaluation
driver scripts/eval_pgd_sweep.py must reproduce the reference's labels, per-image adversarial predictions and the integer
`total_correct` of every u exactly.

Two sweeps: the published attack (eps 8/255, lr 2/255), which flips every image of the random-init network (count 0 for
every u), and a weak attack (eps 1/255, lr 0.25/255) whose robust accuracy is mid-range, so that images near the decision
boundary decide the count."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import golden, ROOT

sys.path.insert(0, os.path.join(ROOT, "scripts"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prefix,eps255,lr255", [("", 8.0, 2.0), ("w_", 1.0, 0.25)])
def test_pgd_sweep_counts_identical_to_reference(prefix, eps255, lr255):
    import eval_pgd_sweep as ev
    g = golden("pgd_sweep.npz")
    n = int(g["n_images"])
    us = [float(u) for u in g["u_grid"]]
    labels, preds, counts, _ = ev.sweep(n, 256, us, eps=eps255 / 255., lr=lr255 / 255.)
    labels = labels.cpu().numpy()
    assert np.array_equal(labels, g["labels"])            # clean predictions of all 512 images (min clean margin 9.6e-5)
    for u in us:
        tag = prefix + ("%.2f" % u).replace(".", "p")
        ref = g["pred_u" + tag]
        got = preds[u].cpu().numpy()
        bad = np.nonzero(ref != got)[0]
        assert bad.size == 0, (u, bad[:8], g["margin_u" + tag][bad[:8]])
        assert counts[u] == int(g["correct_u" + tag]) == int((got == labels).sum()), u
    if prefix == "w_":      # the weak sweep must be a non-trivial count
        assert all(0 < int(g["correct_u" + prefix + ("%.2f" % u).replace(".", "p")]) < n for u in us)
