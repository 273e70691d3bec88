"""Deterministic, platform-independent pseudo-random tensors (splitmix64 counter hash).

Used so that golden fixtures only need to store OUTPUTS: inputs/weights are regenerated
bit-identically here, on the GPU box, and inside tests from (shape, seed).
Pure uint64 integer arithmetic in numpy -> identical on every machine / numpy version.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def det_uniform(shape, seed, lo=-1.0, hi=1.0, dtype=np.float32):
    """U[lo,hi) with 24 random mantissa bits, element i = hash(seed, i)."""
    n = int(np.prod(shape)) if len(shape) else 1
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _splitmix64(idx ^ _splitmix64(np.uint64(seed) * np.uint64(0x2545F4914F6CDD1D) & _M64))
    u = (h >> np.uint64(40)).astype(np.float64) / float(1 << 24)       # exact in fp64 and fp32
    out = (lo + (hi - lo) * u).astype(dtype)
    return out.reshape(shape)


def det_normal(shape, seed, std=1.0, dtype=np.float32):
    """Approximately normal (sum of 4 uniforms, variance-matched); cheap and portable."""
    acc = np.zeros(shape, dtype=np.float64)
    for k in range(4):
        acc += det_uniform(shape, seed * 4 + k + 1000003, -1.0, 1.0, np.float64)
    return (acc * (std * (3.0 / 4.0) ** 0.5)).astype(dtype)
