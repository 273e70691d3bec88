"""Fused callers either side of the ODE-block path:

  * `attack_step`: the elementwise steps of FGSM / FGSM-random / PGD (MegaAdversarial/src/attacks/fgsm.py:27-40,
    93-105, pgd.py:28-53) as ONE CUDA kernel each, bit-identical to the reference's chain of torch calls;
  * `FusedSGD`: torch.optim.SGD (momentum, weight decay; examples/cifar10/train_and_attack.py:98-99) over ONE flat
    parameter / gradient buffer: one all-reduce over the flat gradient (no pack / unpack copies) and one update
    kernel with the 1/world average folded in.

Both call the C ABI (include/metasolver_b200.h: msb_attack_step, msb_sgd_step); CUDA fp32 tensors only.
"""
import ctypes

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

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0):
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
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
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

    def step(self, grad_scale=1.0):
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("metasolver_b200.FusedSGD.step() passes lr / grad_scale / the first-step flag by value: a "
                               "captured step would replay a frozen learning rate.  Call it outside the CUDA graph.")
        has = self._gather_grads()
        self.lr = float(self.param_groups[0]["lr"])
        dev = self.flat_param.device
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
        with torch.cuda.device(dev):
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for off, k, first in runs:
                mom = self.momentum_buf[off:off + k] if self.momentum_buf is not None else None
                _cabi.check(_cabi.lib().msb_sgd_step(_ptr(self.flat_param[off:off + k]), _ptr(self.flat_grad[off:off + k]),
                                                     _ptr(mom), k, self.lr, self.momentum, self.weight_decay,
                                                     float(grad_scale), 1 if first else 0, st), "sgd_step")
