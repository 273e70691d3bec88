#!/usr/bin/env python
"""BASELINE config 4: premetanode10 FGSM-random adversarial training with solver smoothing, data-parallel.

    python scripts/bench_c4.py [--steps 8] [--batch 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_c4.py

One step per rank (examples/cifar10/train_and_attack.py:246-327 with --opt-level O0): draw u ~ N(0.5, 0.0125) on rank 0
and broadcast it (parallel.sync_solver_params), zero_grad, FGSM-random attack pass (forward + backward; its parameter
gradients accumulate, fgsm.py:98), training pass (forward + backward) on the adversarial batch, ONE all-reduce of the flat
fp32 gradient (FusedSGD.all_reduce), fused SGD-momentum update, CyclicLR schedule of :500-505 (metasolver_b200.CyclicLR).
The batch comes from a uint8 dataset resident in HBM through the one-kernel crop / flip / normalise transform
(metasolver_b200.augment_normalize, data.py:40-46).  Per-GPU batch fixed (weak scaling).  Timing: CUDA events, max over ranks.  One JSON line on rank 0."""
import argparse
import json
import os
import sys
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--nccl-allreduce", action="store_true", help="NCCL all-reduce + update kernel instead of the one peer-memory kernel")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    import metasolver_b200 as msb
    from metasolver_b200 import parallel, detrand
    from metasolver_b200.sopa.src.solvers.utils import create_solver, sample_solver_by_noising_params
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import premetanode10
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    from metasolver_b200.sopa.src.models.odenet_cifar10.data import CIFAR_MEAN, CIFAR_STD
    from metasolver_b200.MegaAdversarial.src.attacks import FGSMRandom
    rank, world, dev = parallel.init_distributed()
    torch.manual_seed(602)
    model = premetanode10((Identity,) * 3, (lambda t: t,) * 3, (F.gelu,) * 3, in_planes=64, is_odenet=True)
    model = model.to(dev).to(memory_format=torch.channels_last).train()
    # world > 1: the gradient lives in a PeerExchange buffer, all-reduce + SGD update = ONE kernel per rank (csrc/peer.cu)
    opt = msb.FusedSGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-4, peer=(world > 1 and not a.nccl_allreduce))
    sched = msb.CyclicLR(opt, base_lr=1e-4, max_lr=0.2, step_size_up=2000, mode="triangular2", cycle_momentum=True)
    atk = FGSMRandom(model, alpha=10 / 255., epsilon=8 / 255., mu=CIFAR_MEAN, std=CIFAR_STD)
    base = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, dev)
    base.freeze_params()
    opts = Namespace(solver_mode="standalone")
    B = a.batch
    n_data = 4 * B                                             # this rank's shard of a synthetic uint8 dataset, resident in HBM
    data_u8 = torch.from_numpy((detrand.uniform((n_data, 32, 32, 3), 700 + rank, 0.0, 256.0)).astype("uint8")).to(dev)
    labels = torch.from_numpy((detrand.uniform((n_data,), 800 + rank, 0.0, 10.0)).astype("int64") % 10).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    it = [0]

    def step():
        index = torch.randint(0, n_data, (B,), device=dev, generator=gen)
        x = msb.augment_normalize(data_u8, index, generator=gen)     # crop + flip + ToTensor + Normalize: one kernel
        y = labels[index]
        s = sample_solver_by_noising_params(base, std=0.0125, bernoulli_p=1.0, noise_type="normal")
        parallel.sync_solver_params([s])                       # every rank integrates with rank 0's u
        kws = {"solvers": [s], "solver_options": opts}
        opt.zero_grad()
        xa, _ = atk(x, y, kws)
        loss = msb.cross_entropy(model(xa, **kws), y)
        loss.backward()
        opt.reduce_and_step()                                  # ONE exchange of the flat gradient, 1/world and the update folded in
        sched.step()
        it[0] += 1
        return loss

    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):            # sample_solver_by_noising_params prints the draw (as the reference does)
        for _ in range(a.warmup):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loss = step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item()) / a.steps
    exchange = "single rank: update kernel only"
    identical = None
    if world > 1:
        if opt.peer is not None:
            opt.peer.check()                                   # raises if a handshake timed out
            exchange = "one kernel per rank over peer memory (msb_peer_allreduce_sgd): rank-order sum + SGD update"
        else:
            exchange = "NCCL all-reduce + update kernel" + ("".join("; " + n for n in opt.peer_note))
        chk = torch.stack([opt.flat_param.double().sum(), opt.flat_param.double().abs().sum()])
        allchk = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allchk, chk)
        identical = all(bool(torch.equal(c, allchk[0])) for c in allchk)      # replicas after warmup + steps updates
    if rank == 0:
        print(json.dumps(dict(config="C4 FGSM-random adversarial training step, solver smoothing u~N(0.5,0.0125) per batch, FusedSGD + CyclicLR, "
                                     "on-GPU crop/flip augmentation", n_gpus=world, batch_per_gpu=B, ms_per_step=ms,
                              images_per_s=world * B / ms * 1e3, scaling="weak", final_loss=float(loss.item()),
                              collectives_per_step="1 all-reduce of %d B + 1 broadcast of 16 B" % (opt.flat_grad.numel() * 4),
                              exchange=exchange, replicas_identical=identical)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
