"""String -> layer factories of sopa/src/models/odenet_cifar10/utils.py:15-76 (same keys)."""
from functools import partial

import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm, weight_norm


class Identity(nn.Module):
    """'NF' normalisation: no-op that swallows the channel-count argument (utils.py:8-13)."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, x):
        return x


_NORMS = {'BN': lambda g: nn.BatchNorm2d, 'LN': lambda g: partial(nn.GroupNorm, 1),
          'GN': lambda g: partial(nn.GroupNorm, g), 'IN': lambda g: nn.InstanceNorm2d, 'NF': lambda g: Identity}
_PARAM_NORMS = {'SN': spectral_norm, 'WN': weight_norm, 'PNF': (lambda m: m)}
_ACTS = {'ReLU': F.relu, 'GeLU': F.gelu, 'Softsign': F.softsign, 'Tanh': F.tanh,
         'AF': partial(F.leaky_relu, negative_slope=1)}


def get_normalization(key, num_groups=32):
    if key not in _NORMS:
        raise NameError('Unknown layer normalization type')
    return _NORMS[key](num_groups)


def get_param_normalization(key):
    if key not in _PARAM_NORMS:
        raise NameError('Unknown param normalization type')
    return _PARAM_NORMS[key]


def get_activation(key):
    if key not in _ACTS:
        raise NameError('Unknown activation type')
    return _ACTS[key]
