"""Fused callers either side of the ODE-block path:

  * `attack_step`: the elementwise steps of FGSM / FGSM-random / PGD (MegaAdversarial/src/attacks/fgsm.py:27-40,
    93-105, pgd.py:28-53) as ONE CUDA kernel each, bit-identical to the reference's chain of torch calls;
  * `FusedSGD`: torch.optim.SGD (momentum, weight decay; examples/cifar10/train_and_attack.py:98-99) over ONE flat
    parameter / gradient buffer: one all-reduce over the flat gradient (no pack / unpack copies) and one update
    kernel with the 1/world average folded in;
  * `CyclicLR`: the learning-rate / momentum schedule of examples/cifar10/train_and_attack.py:500-505
    (torch.optim.lr_scheduler.CyclicLR, which only accepts torch.optim.Optimizer instances) for `FusedSGD`: host
    arithmetic, value for value what torch computes;
  * `augment_normalize`: the training-input transform (crop / flip / ToTensor / Normalize) as ONE kernel.

They call the C ABI (include/metasolver_b200.h: msb_attack_step, msb_sgd_step, msb_augment_batch); CUDA fp32 tensors only.
"""
import ctypes
import math

import torch
import torch.distributed as dist

from . import _cabi


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _layout(x):
    """(tensor with dense memory, channels_last flag) for a (B,C,H,W) image tensor."""
    if x.dim() != 4:
        raise ValueError("metasolver_b200.attack_step: expected a (B,C,H,W) tensor, got shape %s" % (tuple(x.shape),))
    if x.is_contiguous():
        return x, 0
    if x.is_contiguous(memory_format=torch.channels_last):
        return x, 1
    return x.contiguous(), 0


def attack_step(kind, a, grad=None, ref=None, eps=0.0, step=0.0, normalize_out=False, chan_consts=None):
    """One fused attack step (see msb_attack_step).  `a`, `grad`, `ref`: CUDA fp32 (B,C,H,W) tensors of one memory
    format; chan_consts: up to four per-channel constant lists [c0, c1, c2, c3] (missing ones default to 0 / 1)."""
    if not a.is_cuda or a.dtype != torch.float32:
        raise RuntimeError("metasolver_b200.attack_step: CUDA float32 tensors only (got %s %s)" % (a.device, a.dtype))
    a, cl = _layout(a)
    B, C, H, W = a.shape
    if C > _cabi.ATTACK_MAX_CHANNELS:
        raise ValueError("metasolver_b200.attack_step: at most %d channels" % _cabi.ATTACK_MAX_CHANNELS)

    def same(t):
        if t is None:
            return None
        if t.shape != a.shape or t.dtype != a.dtype or t.device != a.device:
            raise ValueError("metasolver_b200.attack_step: operand shape / dtype / device mismatch")
        return t.contiguous(memory_format=torch.channels_last) if cl else t.contiguous()

    grad, ref = same(grad), same(ref)
    out = torch.empty_like(a)
    consts = None
    if chan_consts is not None:
        rows = []
        for j in range(4):
            row = chan_consts[j] if j < len(chan_consts) and chan_consts[j] is not None else ([1.0] * C if j == 1 else [0.0] * C)
            row = [float(v) for v in (row.reshape(-1).tolist() if isinstance(row, torch.Tensor) else
                                      (row if isinstance(row, (list, tuple)) else [row] * C))]
            if len(row) == 1:
                row = row * C
            if len(row) != C:
                raise ValueError("metasolver_b200.attack_step: per-channel constant of length %d for %d channels" % (len(row), C))
            rows += row
        consts = (ctypes.c_float * (4 * C))(*rows)
    with torch.cuda.device(a.device):
        _cabi.check(_cabi.lib().msb_attack_step(int(kind), _ptr(a), _ptr(grad), _ptr(ref), _ptr(out), a.numel(), C, H * W, cl,
                                                float(eps), float(step), 1 if normalize_out else 0, consts,
                                                ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)),
                    "attack_step")
    return out


class FusedSGD:
    """SGD with momentum and weight decay on one flat buffer.

    The parameters of `params` are re-pointed at views of ONE flat fp32 buffer and their `.grad`s at views of one flat
    gradient buffer, so that `step()` is a single kernel and a data-parallel run all-reduces the flat gradient in
    place (one all-reduce of 2.70 MB per step for premetanode10)."""

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, peer=None):
        """peer: None / False = gradients in ordinary device memory, `all_reduce()` is an NCCL / gloo call; True = allocate a
        `parallel.PeerExchange` for the flat gradient (collective; multi-rank CUDA job on one node; falls back to None
        when the mapping is refused) or pass one: `reduce_and_step()` is then ONE kernel per rank (msb_peer_allreduce_sgd)."""
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedSGD: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.dtype != torch.float32 or p.device != dev for p in self.params):
            raise RuntimeError("metasolver_b200.FusedSGD: CUDA float32 parameters on one device only")
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.param_groups = [{"lr": self.lr, "momentum": self.momentum, "weight_decay": self.weight_decay}]   # schedulers poke "lr"
        n = sum(p.numel() for p in self.params)
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.peer_note = []
        if peer is True:
            from . import parallel
            peer = parallel.peer_exchange_or_none(n, self.peer_note)
        elif peer is False:
            peer = None
        if peer is not None and (peer.n != n or peer.device != dev):
            raise ValueError("metasolver_b200.FusedSGD: the PeerExchange holds %d floats on %s, the parameters %d on %s"
                             % (peer.n, peer.device, n, dev))
        self.peer = peer
        self.flat_grad = peer.grad if peer is not None else torch.zeros(n, dtype=torch.float32, device=dev)
        self.momentum_buf = torch.zeros(n, dtype=torch.float32, device=dev) if self.momentum != 0.0 else None
        self._seen = [False] * len(self.params)      # torch.optim.SGD creates a momentum buffer at a parameter's first update
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_param[off:off + k].view(p.shape)
                p.grad = self.flat_grad[off:off + k].view(p.shape)
                off += k

    def zero_grad(self, set_to_none=False):
        """Gradients stay views of the flat buffer (set_to_none would detach them from it)."""
        self.flat_grad.zero_()
        off = 0
        for p in self.params:                       # re-attach if something replaced p.grad
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + k].view(p.shape)
            off += k

    def _gather_grads(self):
        """Make the flat gradient buffer hold what the parameters' `.grad`s hold.

        `model.zero_grad()` / `zero_grad(set_to_none=True)` (the torch >= 2.0 default) detach `.grad` from the flat buffer:
        the next backward then writes fresh tensors the flat buffer never sees.  Stray gradients are copied in and
        re-attached; a parameter whose grad is None gets a zero slice (for the all-reduce sum) and is reported so that
        step() can skip it like torch.optim.SGD does.  Returns the list of booleans "has a gradient" per parameter."""
        has = []
        off = 0
        base = self.flat_grad.data_ptr()
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                view = self.flat_grad[off:off + k].view(p.shape)
                if p.grad is None:
                    view.zero_()
                    has.append(False)
                else:
                    if p.grad.data_ptr() != base + 4 * off:
                        if p.grad.shape != p.shape or p.grad.dtype != torch.float32 or p.grad.device != self.flat_grad.device:
                            raise RuntimeError("metasolver_b200.FusedSGD: gradient of shape %s / %s on %s does not match its "
                                               "parameter" % (tuple(p.grad.shape), p.grad.dtype, p.grad.device))
                        view.copy_(p.grad)
                        p.grad = view
                    has.append(True)
                off += k
        return has

    def all_reduce(self):
        """Sum the flat gradient over the ranks; returns the scale (1/world) that step() must apply."""
        self._gather_grads()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            return 1.0 / dist.get_world_size()
        return 1.0

    def _begin_step(self):
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("metasolver_b200.FusedSGD.step() passes lr / grad_scale / the first-step flag by value: a "
                               "captured step would replay a frozen learning rate.  Call it outside the CUDA graph.")
        has = self._gather_grads()
        self.lr = float(self.param_groups[0]["lr"])
        self.momentum = float(self.param_groups[0].get("momentum", self.momentum))      # CyclicLR(cycle_momentum=True) pokes it
        if self.momentum != 0.0 and self.momentum_buf is None:
            self.momentum_buf = torch.zeros_like(self.flat_param)
        return self._runs(has)

    def _runs(self, has):
        # contiguous runs of parameters that have a gradient and share the "first update" state: ONE run (one kernel)
        # in the normal case; parameters without a gradient are skipped entirely (no weight decay, no momentum
        # update), exactly like torch.optim.SGD
        runs = []
        off = 0
        for i, p in enumerate(self.params):
            k = p.numel()
            if has[i]:
                first = not self._seen[i]
                if runs and runs[-1][0] + runs[-1][1] == off and runs[-1][2] == first:
                    runs[-1][1] += k
                else:
                    runs.append([off, k, first])
                self._seen[i] = True
            off += k
        return runs

    def reduce_and_step(self):
        """The exchange step of data-parallel training: average the flat gradient over the ranks and apply the update.
        With a PeerExchange this is ONE kernel per rank (peer-memory reads in rank order + update, csrc/peer.cu) -- every
        rank must hold gradients for the same parameters; without one it is `step(all_reduce())`."""
        if self.peer is None:
            return self.step(self.all_reduce())
        for off, k, first in self._begin_step():
            mom = self.momentum_buf[off:off + k] if self.momentum_buf is not None else None
            self.peer.allreduce_sgd(params=self.flat_param[off:off + k], momentum_buf=mom, lr=self.lr, momentum=self.momentum,
                                    weight_decay=self.weight_decay, first_step=first, offset=off, n=k)

    def step(self, grad_scale=1.0):
        runs = self._begin_step()
        dev = self.flat_param.device
        with torch.cuda.device(dev):
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for off, k, first in runs:
                mom = self.momentum_buf[off:off + k] if self.momentum_buf is not None else None
                _cabi.check(_cabi.lib().msb_sgd_step(_ptr(self.flat_param[off:off + k]), _ptr(self.flat_grad[off:off + k]),
                                                     _ptr(mom), k, self.lr, self.momentum, self.weight_decay,
                                                     float(grad_scale), 1 if first else 0, st), "sgd_step")


class CyclicLR:
    """torch.optim.lr_scheduler.CyclicLR (examples/cifar10/train_and_attack.py:500-505) for any optimizer object with
    `param_groups` (FusedSGD): same constructor arguments, same values -- the arithmetic below is torch's, in Python
    doubles.  Construction sets lr = base_lr (and momentum = max_momentum when cycle_momentum); call `step()` after every
    optimizer step."""

    def __init__(self, optimizer, base_lr, max_lr, step_size_up=2000, step_size_down=None, mode="triangular", gamma=1.0,
                 scale_fn=None, scale_mode="cycle", cycle_momentum=True, base_momentum=0.8, max_momentum=0.9):
        if mode not in ("triangular", "triangular2", "exp_range") and scale_fn is None:
            raise ValueError("mode is invalid and scale_fn is None")
        self.optimizer = optimizer
        self.base_lr, self.max_lr = float(base_lr), float(max_lr)
        up = float(step_size_up)
        down = float(step_size_down) if step_size_down is not None else up
        self.total_size = up + down
        self.step_ratio = up / self.total_size
        self.gamma = gamma
        if scale_fn is not None:
            self.scale_fn, self.scale_mode = scale_fn, scale_mode
        elif mode == "triangular":
            self.scale_fn, self.scale_mode = (lambda x: 1.0), "cycle"
        elif mode == "triangular2":
            self.scale_fn, self.scale_mode = (lambda x: 1 / (2.0 ** (x - 1))), "cycle"
        else:
            self.scale_fn, self.scale_mode = (lambda x: self.gamma ** x), "iterations"
        self.cycle_momentum = bool(cycle_momentum)
        self.base_momentum, self.max_momentum = float(base_momentum), float(max_momentum)
        if self.cycle_momentum and "momentum" not in optimizer.param_groups[0]:
            raise ValueError("optimizer must support momentum with `cycle_momentum` option enabled")
        self.last_epoch = -1
        self._last_lr = [self.base_lr]
        self.step()

    def _values(self):
        cycle = math.floor(1 + self.last_epoch / self.total_size)
        x = 1.0 + self.last_epoch / self.total_size - cycle
        scale_factor = x / self.step_ratio if x <= self.step_ratio else (x - 1) / (self.step_ratio - 1)
        arg = cycle if self.scale_mode == "cycle" else self.last_epoch
        lr = self.base_lr + (self.max_lr - self.base_lr) * scale_factor * self.scale_fn(arg)
        momentum = self.max_momentum - (self.max_momentum - self.base_momentum) * scale_factor * self.scale_fn(arg)
        return lr, momentum

    def step(self):
        self.last_epoch += 1
        lr, momentum = self._values()
        for g in self.optimizer.param_groups:
            g["lr"] = lr
            if self.cycle_momentum:
                g["momentum"] = momentum
        self._last_lr = [lr for _ in self.optimizer.param_groups]

    def get_last_lr(self):
        return self._last_lr


def augment_normalize(images_u8, index=None, generator=None, padding=4, train=True, mean=None, std=None, draws=None):
    """The input transform of sopa/src/models/odenet_cifar10/data.py:40-57 for a batch taken from a uint8 dataset resident
    on the GPU, as ONE kernel (msb_augment_batch): RandomCrop(H, padding) + RandomHorizontalFlip + ToTensor + Normalize when
    `train`, ToTensor + Normalize otherwise.  images_u8: uint8 (N, H, W, C) CUDA tensor (CIFAR's native layout); index:
    int64 CUDA tensor of sample ids (None = all N in order); crop offsets / flip bits are drawn on the device from
    `generator` (or given as draws = (dx, dy, flip)).  Returns the normalised float32 batch (B, C, H, W) in channels_last
    memory -- the stem kernel's input layout -- bit-identical to torchvision's result for the same draws."""
    from .sopa.src.models.odenet_cifar10.data import CIFAR_MEAN, CIFAR_STD
    if not images_u8.is_cuda or images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or not images_u8.is_contiguous():
        raise RuntimeError("metasolver_b200.augment_normalize: contiguous uint8 CUDA tensor (N, H, W, C) required")
    N, H, W, C = images_u8.shape
    dev = images_u8.device
    mean = tuple(CIFAR_MEAN if mean is None else mean)
    std = tuple(CIFAR_STD if std is None else std)
    if len(mean) != C or len(std) != C or C > _cabi.ATTACK_MAX_CHANNELS:
        raise ValueError("augment_normalize: %d channels need %d means / stds (at most %d channels)" % (C, C, _cabi.ATTACK_MAX_CHANNELS))
    if index is not None:
        index = index.to(device=dev, dtype=torch.int64).contiguous()
    B = N if index is None else index.numel()
    dx = dy = flip = None
    if draws is not None:
        dx, dy, flip = draws
    elif train:
        dx = torch.randint(0, 2 * padding + 1, (B,), device=dev, generator=generator)
        dy = torch.randint(0, 2 * padding + 1, (B,), device=dev, generator=generator)
        flip = torch.rand(B, device=dev, generator=generator) < 0.5
    if dx is not None:
        dx = dx.to(device=dev, dtype=torch.int32).contiguous()
        dy = dy.to(device=dev, dtype=torch.int32).contiguous()
        flip = flip.to(device=dev, dtype=torch.uint8).contiguous()
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
    f3 = ctypes.c_float * C
    with torch.cuda.device(dev):
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(_cabi.lib().msb_augment_batch(_ptr(images_u8), _ptr(index), _ptr(dx), _ptr(dy), _ptr(flip), B, H, W, C,
                                                  int(padding), f3(*mean), f3(*std), _ptr(out), st), "augment_batch")
    return out
