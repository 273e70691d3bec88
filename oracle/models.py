"""Oracle: whole-network forward of the published CIFAR-10 config (premetanode10, NF + GeLU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functional restatement of `sopa/src/models/odenet_cifar10/layers.py`:
  MetaNODE.forward :408-426 (note :339-342 -- `is_preactivation` is always False because the
  isinstance test is applied to a class, so the stem runs `act(bn1(conv1(x)))` and there is no
  activation before pooling), MetaLayer.forward :289-301, PreBasicBlock.forward :77-81,
  premetanode10 factory :520-530 (`[(1,1),(1,1)]`: one residual + one ODE block per layer).
`params` uses the reference's state-dict key names.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .detrand import det_uniform
from .rk import integrate, rhs_preact, ode_block_forward

PREMETANODE10_KEYS = [
    ("conv1.weight", (64, 3, 3, 3)),
    ("layer1.blocks_res.0.conv1.weight", (64, 64, 3, 3)),
    ("layer1.blocks_res.0.conv2.weight", (64, 64, 3, 3)),
    ("layer1.blocks_ode.0.rhs_func.conv1.weight", (64, 64, 3, 3)),
    ("layer1.blocks_ode.0.rhs_func.conv2.weight", (64, 64, 3, 3)),
    ("layer2.blocks_res.0.conv1.weight", (128, 64, 3, 3)),
    ("layer2.blocks_res.0.conv2.weight", (128, 128, 3, 3)),
    ("layer2.blocks_res.0.shortcut.0.weight", (128, 64, 1, 1)),
    ("layer2.blocks_ode.0.rhs_func.conv1.weight", (128, 128, 3, 3)),
    ("layer2.blocks_ode.0.rhs_func.conv2.weight", (128, 128, 3, 3)),
    ("fc_layers.2.weight", (10, 128)),
    ("fc_layers.2.bias", (10,)),
]

CIFAR_MEAN = (0.4914, 0.4822, 0.4465)      # sopa/src/models/odenet_cifar10/data.py:45
CIFAR_STD = (0.2023, 0.1994, 0.2010)


def det_premetanode10_params(seed0=500):
    """Deterministic weights, same recipe as tests/golden/make_golden.py section F."""
    p = {}
    for i, (k, shape) in enumerate(PREMETANODE10_KEYS):
        if len(shape) == 4:
            bound = 1.0 / np.sqrt(shape[1] * shape[2] * shape[3])
            a = det_uniform(shape, seed0 + i, -bound, bound)
        elif len(shape) == 2:
            bound = 1.0 / np.sqrt(shape[1])
            a = det_uniform(shape, seed0 + i, -bound, bound)
        else:
            a = det_uniform(shape, seed0 + i, -0.1, 0.1)
        p[k] = torch.from_numpy(a)
    return p


def _pre_basic_block(x, w1, w2, w_sc, stride):
    # cifar10/layers.py:77-81 with bn = Identity, act = gelu
    out = F.conv2d(F.gelu(x), w1, None, stride, 1)
    out = F.conv2d(F.gelu(out), w2, None, 1, 1)
    sc = x if w_sc is None else F.conv2d(x, w_sc, None, stride, 0)
    out = out + sc
    return out


def premetanode10_forward(p, x, tableau, grid, counters=None, taps=None, ensemble_weights=None):
    """x: normalised images (B,3,32,32) -> logits (B,10).  `taps` (dict) collects ODE-block outputs.
    `tableau` / `grid` may be lists: every ODE block then runs the solver-ensemble regime
    (ensemble_prob = 1, cifar10/layers.py:190-203) with `ensemble_weights` (None = uniform)."""
    t = torch.tensor([0, 1]).float()
    out = F.gelu(F.conv2d(x, p["conv1.weight"], None, 1, 1))                       # :411-413
    for li, stride in ((1, 1), (2, 2)):
        pre = "layer%d." % li
        out = _pre_basic_block(out, p[pre + "blocks_res.0.conv1.weight"], p[pre + "blocks_res.0.conv2.weight"],
                               p.get(pre + "blocks_res.0.shortcut.0.weight"), stride)
        rhs = rhs_preact(p[pre + "blocks_ode.0.rhs_func.conv1.weight"], p[pre + "blocks_ode.0.rhs_func.conv2.weight"],
                         "gelu", None if counters is None else counters[li - 1])
        if isinstance(tableau, (list, tuple)):
            out = ode_block_forward(out, rhs, list(tableau), list(grid), "ensemble", ensemble_prob=1.0,
                                    ensemble_weights=ensemble_weights, t=t)
        else:
            out = integrate(tableau, rhs, out, t, **grid)[-1]                       # MetaODEBlock :176-177,207
        if taps is not None:
            taps[pre + "blocks_ode.0"] = out
    out = F.adaptive_avg_pool2d(out, (1, 1)).flatten(1)                             # :390-392, 425
    return F.linear(out, p["fc_layers.2.weight"], p["fc_layers.2.bias"])
