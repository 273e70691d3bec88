#!/usr/bin/env python
"""BASELINE config 5: PGD-7 robust-accuracy sweep over the RK2 u grid, batch-sharded over the ranks.

    python scripts/eval_pgd_sweep.py [--n-images 8192] [--batch 512] [--u-grid 0.05:1.0:0.05 | --u 0.1,0.35,0.5,1.0]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/eval_pgd_sweep.py ...

Protocol (reference: examples/cifar10/train_and_attack.py:212-243, MegaAdversarial/src/attacks/pgd.py:23-57,
sopa/src/models/odenet_mnist/metrics.py:27-41): fixed synthetic set img ~ U[0,1) from a counter hash (the same on every
rank / device; each rank generates only its shard), labels := the clean argmax under the nominal solver (RK2 u = 0.5,
8 steps) so clean accuracy is 100 %, PGD eps = 8/255, lr = 2/255, 7 iterations from a HOST-generated random start; for
each u the attack and the evaluation use the same solver.  The images are split contiguously over the ranks
(parallel.shard_range), there is no data-path communication: the per-u counts stay on the device and ONE integer
all-reduce at the end of the sweep sums the whole `total_correct` vector (parallel.allreduce_sum_counts; round 2 start: one
host-synchronising all-reduce per u, and the NCCL communicator set-up inside the timed region).  Prints one JSON line on rank 0: {"total_correct": {u: count}, "images_per_s": ...}.
With --golden FILE (tests/golden/pgd_sweep.npz) the labels, per-image predictions and counts of the first images are
compared with the reference's.
"""
import argparse
import json
import os
import sys
import time
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SEED_IMG, SEED_NOISE = 9100, 9101
EPS, LR, N_ITER = 8 / 255., 2 / 255., 7


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-images", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--u-grid", default="0.05:1.0:0.05", help="lo:hi:step (inclusive)")
    ap.add_argument("--u", default=None, help="comma-separated u values (overrides --u-grid)")
    ap.add_argument("--checkpoint", default=None, help="state dict to load (default: deterministic random-init weights)")
    ap.add_argument("--eps255", type=float, default=8.0, help="attack radius in 1/255 units (8 = the published setting)")
    ap.add_argument("--lr255", type=float, default=2.0, help="attack step in 1/255 units")
    ap.add_argument("--golden", default=None)
    ap.add_argument("--golden-prefix", default="", help="key prefix inside the golden file ('w_' = the weak-attack sweep)")
    ap.add_argument("--out", default=None)
    return ap.parse_args()


def u_values(a):
    if a.u:
        return [float(v) for v in a.u.split(",")]
    lo, hi, st = (float(v) for v in a.u_grid.split(":"))
    n = int(round((hi - lo) / st)) + 1
    return [round(lo + i * st, 10) for i in range(n)]


def sweep(n_images, batch, us, checkpoint=None, device=None, rank=0, world=1, log=None, eps=EPS, lr=LR):
    """Returns (labels of this rank's shard, {u: predictions of the shard}, {u: global total_correct}, seconds)."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    import metasolver_b200  # noqa: F401
    from metasolver_b200 import parallel, detrand
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import premetanode10
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    from metasolver_b200.sopa.src.models.odenet_cifar10.data import CIFAR_MEAN, CIFAR_STD, normalize
    from metasolver_b200.MegaAdversarial.src.attacks import PGD

    dev = device or torch.device("cuda", torch.cuda.current_device())
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = premetanode10((Identity,) * 3, (lambda x: x,) * 3, (F.gelu,) * 3, in_planes=64, is_odenet=True)
    if checkpoint:
        model.load_state_dict(torch.load(checkpoint, map_location="cpu"))
    else:
        model.load_state_dict(detrand.premetanode10_state_dict(model))
    model = model.to(dev).to(memory_format=torch.channels_last).eval()

    lo, hi = parallel.shard_range(n_images, rank, world)
    per_img = 3 * 32 * 32
    img = torch.from_numpy(detrand.uniform((hi - lo, 3, 32, 32), SEED_IMG, 0.0, 1.0, offset=lo * per_img))
    noise = torch.from_numpy(detrand.uniform((hi - lo, 3, 32, 32), SEED_NOISE, -eps, eps, offset=lo * per_img))
    x = normalize(img).to(dev).contiguous(memory_format=torch.channels_last)
    noise = noise.to(dev).contiguous(memory_format=torch.channels_last)

    def kw(u):
        s = create_solver("rk2", "u", 8, -1, u, -1, torch.float32, dev)
        s.freeze_params()
        return {"solvers": [s], "solver_options": Namespace(solver_mode="standalone")}

    with torch.no_grad():
        k = kw(0.5)
        labels = torch.cat([model(x[i:i + batch], **k).argmax(1) for i in range(0, hi - lo, batch)]) if hi > lo \
            else torch.zeros(0, dtype=torch.long, device=dev)
    preds = {}
    counts_dev = torch.zeros(len(us), dtype=torch.int64, device=dev)
    if hi > lo:                                        # untimed: first-use costs of the attack path (module load, workspaces)
        m = min(hi - lo, 32)
        PGD(model, eps=eps, lr=lr, n_iter=1, mean=CIFAR_MEAN, std=CIFAR_STD)(x[:m], labels[:m], kw(0.5), noise=noise[:m])
    parallel.allreduce_sum_counts(torch.zeros(1, dtype=torch.int64, device=dev))      # untimed: communicator set-up
    torch.cuda.synchronize(dev)
    t0 = time.time()
    for j, u in enumerate(us):
        k = kw(u)
        attack = PGD(model, eps=eps, lr=lr, n_iter=N_ITER, mean=CIFAR_MEAN, std=CIFAR_STD)
        p = []
        for i in range(0, hi - lo, batch):
            xa, _ = attack(x[i:i + batch], labels[i:i + batch], k, noise=noise[i:i + batch])
            with torch.no_grad():
                p.append(model(xa, **k).argmax(1))
        p = torch.cat(p) if p else labels
        preds[u] = p
        counts_dev[j] = (p == labels).sum()            # stays on the device: no host synchronisation inside the sweep
    totals = parallel.allreduce_sum_counts(counts_dev)                                   # ONE integer all-reduce for the sweep
    torch.cuda.synchronize(dev)
    secs = time.time() - t0
    counts = {u: totals[j] for j, u in enumerate(us)}
    if log:
        for u in us:
            log("u=%.2f total_correct %d / %d" % (u, counts[u], n_images))
    return labels, preds, counts, secs


def main():
    a = parse()
    import numpy as np
    import torch
    from metasolver_b200 import parallel
    rank, world, dev = parallel.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("eval_pgd_sweep.py: no CUDA device; the product has no CPU path")
    us = u_values(a)
    labels, preds, counts, secs = sweep(a.n_images, a.batch, us, a.checkpoint, dev, rank, world,
                                        log=(lambda s: print(s, file=sys.stderr, flush=True)) if rank == 0 else None,
                                        eps=a.eps255 / 255., lr=a.lr255 / 255.)
    line = dict(config="C5 PGD-7 (eps %g/255, lr %g/255) robust accuracy over the RK2 u grid, premetanode10, synthetic set" % (a.eps255, a.lr255),
                n_images=a.n_images, n_gpus=world, batch=a.batch, u=us,
                total_correct={("%.2f" % u): counts[u] for u in us},
                images_per_s=a.n_images * len(us) / secs, seconds=secs)
    if a.golden and rank == 0:
        g = np.load(a.golden)
        ng = int(g["n_images"])
        lo, hi = parallel.shard_range(a.n_images, rank, world)
        m = min(ng, hi - lo)
        cmp = dict(images_compared=m, labels_equal=bool((labels[:m].cpu().numpy() == g["labels"][:m]).all()))
        for u in us:
            tag = a.golden_prefix + ("%.2f" % u).replace(".", "p")
            if "pred_u" + tag in g.files:
                ref = g["pred_u" + tag][:m]
                got = preds[u][:m].cpu().numpy()
                cmp["u%.2f" % u] = dict(pred_mismatches=int((ref != got).sum()),
                                        correct_ref=int((ref == g["labels"][:m]).sum()),
                                        correct_ours=int((got == labels[:m].cpu().numpy()).sum()))
        line["vs_reference_golden"] = cmp
    if rank == 0:
        s = json.dumps(line)
        print(s)
        if a.out:
            open(a.out, "w").write(s + "\n")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
