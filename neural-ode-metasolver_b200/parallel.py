"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box,
gloo in the CPU tests).  The ODE-block path shards over the batch with no data-path collective;
training adds exactly one all-reduce of the flat fp32 gradient per step; evaluation
adds one integer all-reduce of the correct-prediction count."""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Read RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment (torchrun contract).
    Returns (rank, world, device).  With WORLD_SIZE unset or 1 nothing is initialised."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world)
    return rank, world, device


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of n_items for this rank (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradAllReducer:
    """One all-reduce per step over a single flat fp32 buffer holding every parameter gradient.

    The gradients are gathered into the flat buffer by ONE concatenation kernel, the buffer is reduced (NCCL: averaging
    inside the collective), and every `p.grad` is then re-pointed at its slice of the buffer (host-side view
    assignments: no unpack copies).  Round 1 issued ~36 pack copies, the all-reduce, a scale and ~36 unpack copies."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            k = p.numel()
            self.views.append(self.flat[off:off + k].view(p.shape))
            off += k

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def __call__(self, average=True):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        pieces = []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
                pieces.append(v.reshape(-1))
            else:
                pieces.append(p.grad.reshape(-1))
        if any(g.data_ptr() != v.data_ptr() for g, v in zip(pieces, self.views)):
            with torch.no_grad():
                torch.cat(pieces, out=self.flat)                  # one gather kernel (no-op when grads already live in `flat`)
        if average and dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.mul_(1.0 / dist.get_world_size())
        for p, v in zip(self.params, self.views):
            p.grad = v


def sync_solver_params(solvers, src=0):
    """Solver smoothing draws u (and v) on the host; all ranks must integrate with the same tableau
    to equal the single-process semantics.  Broadcast rank `src`'s values and rebuild the tableaus."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    vals = []
    for s in solvers:
        for p in (s.u, s.v):
            vals.append(float(p.detach().reshape(-1)[0]) if p is not None else 0.0)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    dist.broadcast(t, src)
    vals = t.cpu().tolist()
    i = 0
    for s in solvers:
        if s.u is not None:
            s.u = torch.tensor((vals[i],), dtype=s.u.dtype)      # keep the dtype the draw had (float32 after noise_params)
        if s.v is not None:
            s.v = torch.tensor((vals[i + 1],), dtype=s.v.dtype)
        i += 2
        s.build_ButcherTableau()


def allreduce_sum_int(value, device):
    """Integer count reduction for accuracy / robust-accuracy evaluation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
