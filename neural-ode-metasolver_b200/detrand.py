"""Deterministic, platform-independent synthetic tensors (splitmix64 counter hash; element i = hash(seed, i)).

Synthetic evaluation sets (scripts/eval_pgd_sweep.py, bench.py) must be the same on every rank, device and numpy
version, and shardable: any slice [lo, hi) of a tensor can be generated on its own (`offset`), so a rank builds only
its shard.  Pure uint64 arithmetic in numpy."""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniform(shape, seed, lo=-1.0, hi=1.0, dtype=np.float32, offset=0):
    """U[lo, hi) with 24 random mantissa bits; `offset` = linear index of the first element in the full tensor."""
    n = int(np.prod(shape)) if len(shape) else 1
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _mix(idx ^ _mix(np.uint64(seed) * np.uint64(0x2545F4914F6CDD1D) & _M64))
    u = (h >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    return (lo + (hi - lo) * u).astype(dtype).reshape(shape)


def premetanode10_state_dict(model, seed0=500):
    """Deterministic weights for a premetanode10-shaped model: parameter i of the state dict (in order) is
    U(+-1/sqrt(fan_in)) from seed0 + i (conv and linear weights), U(+-0.1) for biases."""
    import torch
    new = {}
    for i, (k, v) in enumerate(model.state_dict().items()):
        if v.dim() == 4:
            bound = 1.0 / np.sqrt(v.shape[1] * v.shape[2] * v.shape[3])
        elif v.dim() == 2:
            bound = 1.0 / np.sqrt(v.shape[1])
        else:
            bound = 0.1
        new[k] = torch.from_numpy(uniform(tuple(v.shape), seed0 + i, -bound, bound))
    return new
