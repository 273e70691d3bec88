#!/usr/bin/env python
"""The data-parallel exchange step in isolation: premetanode10's 2.70 MB flat gradient.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_exchange.py

Times, with CUDA events on the launching stream (max over ranks), back-to-back
  (a) NCCL all-reduce (SUM) + msb_sgd_step (1/world folded in)           -- two kernels, the round-1 form
  (b) msb_peer_allreduce_sgd                                             -- ONE kernel over peer memory (csrc/peer.cu)
and checks that both leave the same parameters (to rounding: NCCL's summation order differs) and that (b)'s replicas are
bitwise identical across the ranks.  One JSON line on rank 0; peer-read GB/s = (world - 1) x bytes / time per rank."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=674762)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--forms", default="1,2", help="peer_form values to time (1 = one-shot, 2 = two-shot)")
    a = ap.parse_args()
    import ctypes
    import torch
    import torch.distributed as dist
    import metasolver_b200 as msb  # noqa: F401
    from metasolver_b200 import parallel, _cabi
    rank, world, dev = parallel.init_distributed()
    if world < 2 or not torch.cuda.is_available():
        raise SystemExit("bench_exchange.py: run under torchrun with >= 2 CUDA ranks")
    n = a.n
    ex = parallel.PeerExchange(n)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    grad = torch.randn(n, device=dev, generator=g)
    ex.grad.copy_(grad)
    torch.manual_seed(5)
    w0 = torch.randn(n, device=dev)
    lib = _cabi.lib()
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())

    p_a, m_a, red = w0.clone(), torch.zeros(n, device=dev), torch.empty(n, device=dev)
    p_b, m_b = w0.clone(), torch.zeros(n, device=dev)

    def two_kernels(first):
        red.copy_(grad)                                   # the all-reduce is in place: restore the local gradient (not timed apart)
        dist.all_reduce(red, op=dist.ReduceOp.SUM)
        _cabi.check(lib.msb_sgd_step(vp(p_a), vp(red), vp(m_a), n, 0.01, 0.9, 5e-4, 1.0 / world, 1 if first else 0, st()), "sgd_step")

    def one_kernel(first):
        ex.allreduce_sgd(params=p_b, momentum_buf=m_b, lr=0.01, momentum=0.9, weight_decay=5e-4, first_step=first)

    def timed(fn):
        for i in range(5):
            fn(False)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.iters):
            fn(False)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / a.iters * 1e3

    # parity on three updates from the same start
    for i in range(3):
        two_kernels(i == 0)
        one_kernel(i == 0)
    rel = float((p_a - p_b).abs().max() / p_b.abs().max())
    chk = torch.stack([p_b.double().sum(), p_b.double().abs().sum()])
    allchk = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    identical = all(bool(torch.equal(c, allchk[0])) for c in allchk)

    def copy_only(first):
        red.copy_(grad)
    us_copy = timed(copy_only)
    us_a = timed(two_kernels) - us_copy
    us_form = {}
    for f in [int(v) for v in a.forms.split(",")]:
        msb.set_option("peer_form", f)
        us_form[{1: "one_shot", 2: "two_shot"}.get(f, str(f))] = timed(one_kernel)
    msb.set_option("peer_form", 0)
    us_b = timed(one_kernel)
    ex.check()
    if rank == 0:
        print(json.dumps(dict(config="exchange step of data-parallel training: flat fp32 gradient of %d floats (%d bytes)" % (n, 4 * n),
                              n_gpus=world, iters=a.iters,
                              nccl_allreduce_plus_sgd_us=us_a, peer_kernel_us=us_b, speedup=us_a / us_b, peer_kernel_us_by_form=us_form,
                              default_form="two-shot" if world >= 3 else "one-shot",
                              max_rel_diff_params_vs_nccl=rel, replicas_bitwise_identical=identical,
                              note="back-to-back exchanges (every one fully synchronises the ranks: two system-scope handshakes per "
                                   "launch); NCCL leg = ncclAllReduce + msb_sgd_step minus the restore copy")))
    ex.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
