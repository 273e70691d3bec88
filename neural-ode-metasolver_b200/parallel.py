"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box,
gloo in the CPU tests).  The ODE-block path shards over the batch with no data-path collective;
training adds exactly one all-reduce of the flat fp32 gradient per step; evaluation
adds one integer all-reduce of the correct-prediction count.

On a CUDA node that all-reduce is not an NCCL call: `PeerExchange` maps the flat gradient buffers of the ranks into each
other (CUDA IPC over NVLink / NVSwitch peer access) and ONE kernel per rank reduces them in rank order and applies the
optimizer update in the same pass (csrc/peer.cu, msb_peer_allreduce_sgd)."""
import ctypes
import os
import socket

import torch
import torch.distributed as dist

from . import _cabi


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


class _DeviceArray:
    """`__cuda_array_interface__` carrier: lets torch view memory the C library allocated (no copy, no ownership)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerExchange:
    """The flat fp32 gradient buffers of all ranks of ONE node, mapped into every rank (C ABI: msb_peer_*).

        ex = PeerExchange(n_floats)          # collective: every rank of the group calls it (handles travel by all_gather_object)
        ex.grad                              # this rank's flat gradient: a torch view of the exchange buffer's payload
        ex.result                            # the averaged gradient after an exchange (second array of the payload)
        ex.allreduce_sgd(params=..., momentum_buf=..., lr=..., ...)      # one kernel: average + SGD update (msb_peer_allreduce_sgd)
        ex.allreduce_sgd()                                               # one kernel: average into ex.result
        ex.check()                           # synchronises and raises if a handshake of this rank timed out

    Works between devices with peer access and between processes sharing one device (the GPU tests run two ranks on
    cuda:0).  Raises if the ranks are not on one host.  `timeout_ms` bounds every in-kernel wait."""

    def __init__(self, n_floats, group=None, timeout_ms=30000):
        if not torch.cuda.is_available():
            raise RuntimeError("metasolver_b200.PeerExchange: needs CUDA devices (the CPU tests use the gloo all-reduce)")
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        if self.world > _cabi.PEER_MAX_RANKS:
            raise RuntimeError("metasolver_b200.PeerExchange: at most %d ranks (one node)" % _cabi.PEER_MAX_RANKS)
        self.n = int(n_floats)
        self._result_off = (self.n + 3) // 4 * 4           # payload = [gradient | result], both 16-byte aligned
        self.timeout_ms = int(timeout_ms)
        self.device = torch.device("cuda", torch.cuda.current_device())
        lib = _cabi.lib()
        self._own, self._opened = None, []
        # every failure is agreed on collectively, so that all ranks raise (or none does) and no rank is left in a collective
        err = None
        handle = ctypes.create_string_buffer(_cabi.PEER_HANDLE_BYTES)
        try:
            base = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                _cabi.check(lib.msb_peer_alloc((self._result_off + self.n) * 4, ctypes.byref(base), handle), "peer_alloc")
            self._own = base.value
        except Exception as exc:
            err = "rank %d: %s" % (self.rank, str(exc).splitlines()[0][:160])
        bases = [None] * self.world
        bases[self.rank] = self._own
        if self.world > 1:
            mine = (socket.gethostname(), bytes(handle.raw), self.n, err)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            errs = [h[3] for h in everyone if h[3]]
            if any(h[0] != mine[0] for h in everyone):
                errs.append("ranks on different hosts %s (peer memory is per node; use the NCCL all-reduce across nodes)"
                            % sorted({h[0] for h in everyone}))
            if any(h[2] != self.n for h in everyone):
                errs.append("ranks disagree on the gradient size %s" % [h[2] for h in everyone])
            if not errs:
                try:
                    with torch.cuda.device(self.device):
                        for r, h in enumerate(everyone):
                            if r == self.rank:
                                continue
                            p = ctypes.c_void_p()
                            _cabi.check(lib.msb_peer_open(h[1], ctypes.byref(p)), "peer_open (rank %d)" % r)
                            self._opened.append(p.value)
                            bases[r] = p.value
                except Exception as exc:
                    err = "rank %d: %s" % (self.rank, str(exc).splitlines()[0][:160])
                second = [None] * self.world
                dist.all_gather_object(second, err, group=group)          # doubles as the barrier after the mappings
                errs = [e for e in second if e]
            if errs:
                self.close(collective=False)
                raise RuntimeError("metasolver_b200.PeerExchange: " + "; ".join(errs))
        elif err:
            raise RuntimeError("metasolver_b200.PeerExchange: " + err)
        self._bases = (ctypes.c_void_p * self.world)(*bases)
        self._carrier = _DeviceArray(self._own + _cabi.PEER_HEADER_BYTES, self._result_off + self.n)
        payload = torch.as_tensor(self._carrier, device=self.device)
        self.grad = payload[:self.n]
        self.result = payload[self._result_off:]

    def allreduce_sgd(self, params=None, momentum_buf=None, lr=0.0, momentum=0.0, weight_decay=0.0, grad_scale=None,
                      first_step=False, offset=0, n=None):
        """Reduce floats [offset, offset+n) of every rank's gradient (rank order), scale (default 1/world) and apply the SGD
        update to `params` (`momentum_buf`): flat fp32 CUDA tensors of n elements; without `params` the average is left in
        `self.result[offset:offset+n]` (valid until this rank's next exchange).  Enqueued on the current stream; all ranks
        must make the same calls."""
        n = self.n - offset if n is None else int(n)
        for t in (params, momentum_buf):
            if t is not None and (t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or t.numel() != n):
                raise ValueError("metasolver_b200.PeerExchange.allreduce_sgd: operands must be contiguous CUDA float32 of %d elements" % n)
        scale = 1.0 / self.world if grad_scale is None else float(grad_scale)
        vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().msb_peer_allreduce_sgd(self._bases, self.world, self.rank, int(offset), n, self._result_off,
                                                           1 if params is None else 0, vp(params), vp(momentum_buf), float(lr),
                                                           float(momentum), float(weight_decay), scale, 1 if first_step else 0,
                                                           self.timeout_ms,
                                                           ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                        "peer_allreduce_sgd")

    def status(self):
        """(error word, last finished epoch) of this rank's header; synchronises the device."""
        err, ep = ctypes.c_uint(), ctypes.c_uint()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            _cabi.check(_cabi.lib().msb_peer_status(ctypes.c_void_p(self._own), ctypes.byref(err), ctypes.byref(ep)), "peer_status")
        return err.value, ep.value

    def check(self):
        err, _ = self.status()
        if err:
            raise RuntimeError("metasolver_b200.PeerExchange: a peer handshake timed out on rank %d (error %d: %s)"
                               % (self.rank, err, {1: "a rank never announced its gradient", 2: "a rank never finished reading / its slice of the average never arrived"}.get(err, "?")))

    def close(self, collective=True):
        """Unmap the peers' buffers, then (after a barrier when `collective`) free the own one."""
        lib = _cabi.lib()
        self.grad = self.result = None
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                lib.msb_peer_close(ctypes.c_void_p(p))
            self._opened = []
            if collective and self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)
            if self._own is not None:
                lib.msb_peer_free(ctypes.c_void_p(self._own))
        self._own = None


def peer_exchange_or_none(n_floats, note=None):
    """PeerExchange when this is a multi-rank CUDA job on one node and the mapping succeeds ON EVERY RANK, else None
    (callers then use the NCCL all-reduce).  `note`, a list, receives the reason."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or not torch.cuda.is_available():
        return None
    try:
        return PeerExchange(n_floats)          # raises on every rank or on none (failures are agreed on collectively)
    except RuntimeError as exc:                # e.g. IPC not permitted in this container, no peer access between the devices
        if note is not None:
            note.append("peer exchange unavailable (%s): NCCL all-reduce" % str(exc).splitlines()[0][:200])
        return None


class GradAllReducer:
    """One all-reduce per step over a single flat fp32 buffer holding every parameter gradient.

    The gradients are gathered into the flat buffer by ONE concatenation kernel, the buffer is reduced, and every `p.grad`
    is then re-pointed at its slice of the result (host-side view assignments: no unpack copies).  Round 1 issued ~36 pack
    copies, the all-reduce, a scale and ~36 unpack copies.

    peer=True (multi-rank CUDA job on one node): the flat buffer is the payload of a `PeerExchange` and the reduction is
    ONE kernel per rank over peer memory (rank-order sum, bitwise identical on all ranks, 1/world folded in) leaving the
    average in the exchange buffer's result array; when the mapping is refused (`note` says why) or peer is False: NCCL
    (`ReduceOp.AVG` in place) / gloo."""

    def __init__(self, params, peer=False):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.note = []
        self.peer = peer_exchange_or_none(n, self.note) if peer and dev.type == "cuda" else None
        self.flat = self.peer.grad if self.peer is not None else torch.zeros(n, dtype=torch.float32, device=dev)
        self.result = self.peer.result if self.peer is not None else self.flat
        self.views = []
        self._in_views = []
        off = 0
        for p in self.params:
            k = p.numel()
            self.views.append(self.result[off:off + k].view(p.shape))
            self._in_views.append(self.flat[off:off + k].view(p.shape))
            off += k

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def __call__(self, average=True):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        pieces = []
        for p, v in zip(self.params, self._in_views):
            if p.grad is None:
                pieces.append(torch.zeros_like(v).reshape(-1))    # not a view of `flat`: torch.cat refuses overlapping in / out
            else:
                pieces.append(p.grad.reshape(-1))
        if any(g.data_ptr() != v.data_ptr() for g, v in zip(pieces, self._in_views)):
            with torch.no_grad():
                torch.cat(pieces, out=self.flat)                  # one gather kernel (no-op when grads already live in `flat`)
        if self.peer is not None:
            self.peer.allreduce_sgd(grad_scale=None if average else 1.0)          # average -> peer.result = self.result
        elif average and dist.get_backend() == "nccl":
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


def allreduce_sum_counts(counts):
    """Sum an int64 vector of counts (one entry per evaluated solver / u value) over the ranks with ONE all-reduce and return
    it as a list of Python ints.  The vector stays on the device until this call: a sweep accumulates into it without host
    synchronisation."""
    t = counts.to(torch.int64).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() != "nccl":
            t = t.cpu()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.cpu().tolist()]
