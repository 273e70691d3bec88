"""CIFAR-10 model zoo of the `sopa` API (sopa/src/models/odenet_cifar10/layers.py) with the ODE
blocks routed to the fused B200 kernels.

Hot path (ours): MetaODEBlock + the RHS modules PreBasicBlock2 / BasicBlock2 -- their forward is
never run op-by-op; `solver.integrate` asks the RHS module for its weights (`fused_rhs_spec`) and
launches the fused steps x stages kernels.  The rest of the published pre-activation networks
(stem conv + activation, identity-shortcut residual block = one Euler step of the ODE
kernels, strided residual block via space-to-depth, average pool + FC head) runs on this library's
own kernels too; configurations those kernels do not cover (BatchNorm residual blocks, the
post-activation `BasicBlock`, weight-normed convolutions, CPU tensors) RAISE unless the caller opts
in to PyTorch's kernels with `metasolver_b200.set_library_fallback(True)`.
Module / parameter names equal the reference's so its checkpoints load unchanged.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .utils import Identity
from ..... import _cabi
from .....ops import ode_block_integrate, resblock_down, stem_conv_act, pool_fc, require_fallback
from ...solvers.rk_parametric import can_stack, integrate_stacked

_EULER_STEP = dict(stages=1, c=[0.0], b=[1.0], w=[[0.0]])

__all__ = ['MetaNODE', 'MetaODEBlock', 'MetaLayer', 'BasicBlock', 'PreBasicBlock', 'BasicBlock2', 'PreBasicBlock2',
           'metanode4', 'metanode6', 'metanode10', 'metanode18', 'metanode34',
           'premetanode4', 'premetanode6', 'premetanode10', 'premetanode18', 'premetanode34']


def _conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, bias=False)


def _act_code(fn):
    if fn is F.gelu:
        return _cabi.ACT_GELU_ERF
    if fn is F.relu:
        return _cabi.ACT_RELU
    raise NotImplementedError("metasolver_b200: ODE-block activation %r is not implemented on the fused path "
                              "(supported: F.gelu, F.relu)" % (fn,))


class Flatten(nn.Module):
    def forward(self, x):
        return x.reshape(x.shape[0], -1)


class BasicBlock(nn.Module):
    """Residual block, post-activation order (layers.py:22-51): act(x + conv2(act(conv1(x)))) is not a step of either fused
    right-hand side, so this block has no kernel of ours: it needs metasolver_b200.set_library_fallback(True)."""
    expansion = 1

    def __init__(self, in_planes, planes, stride=1, norm_layer=None, act_layer=None, param_norm=lambda x: x):
        super().__init__()
        self.conv1 = param_norm(_conv3x3(in_planes, planes, stride))
        self.bn1 = norm_layer(planes)
        self.conv2 = param_norm(_conv3x3(planes, planes))
        self.bn2 = norm_layer(planes)
        self.act = act_layer
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(
                param_norm(nn.Conv2d(in_planes, self.expansion * planes, kernel_size=1, stride=stride, bias=False)),
                norm_layer(self.expansion * planes))

    def forward(self, x):
        require_fallback("the post-activation residual block (BasicBlock)", "only the pre-activation blocks are fused")
        out = self.act(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.act(out + self.shortcut(x))


class PreBasicBlock(nn.Module):
    """Residual block, pre-activation order (layers.py:54-81): one Euler step of the fused ODE kernels (identity
    shortcut) or the space-to-depth strided block (stride 2, 1x1 shortcut); other configurations raise."""
    expansion = 1

    def __init__(self, in_planes, planes, stride=1, norm_layer=None, act_layer=None, param_norm=lambda x: x):
        super().__init__()
        self.bn1 = norm_layer(in_planes)
        self.conv1 = param_norm(_conv3x3(in_planes, planes, stride))
        self.bn2 = norm_layer(planes)
        self.conv2 = param_norm(_conv3x3(planes, planes))
        self.act = act_layer
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(
                param_norm(nn.Conv2d(in_planes, self.expansion * planes, kernel_size=1, stride=stride, bias=False)))

    def _fusable(self, x):
        """An identity-shortcut pre-activation block computes x + conv2(act(conv1(act(x)))): exactly ONE
        explicit-Euler step (dt = 1) of the ODE right-hand side PreBasicBlock2 -- so it can run through the
        same fused tcgen05 kernels (the non-ODE remainder becomes the Amdahl limit)."""
        if not (x.is_cuda and x.dtype == torch.float32 and len(self.shortcut) == 0):
            return False
        if not (isinstance(self.bn1, Identity) and isinstance(self.bn2, Identity)):
            return False
        if self.act not in (F.gelu, F.relu) or type(self.conv1) is not nn.Conv2d or type(self.conv2) is not nn.Conv2d:
            return False
        if hasattr(self.conv1, "weight_orig") or hasattr(self.conv1, "weight_g"):
            return False
        B, C, H, W = x.shape
        return self.conv1.stride == (1, 1) and bool(_cabi.lib().msb_shape_supports_tcgen05(C, H, W))

    def _fusable_down(self, x):
        """Stride-2 block with the 1x1 stride-2 shortcut, C -> 2C: runs on the library's own convolution
        engines after a space-to-depth re-indexing (ops.resblock_down) instead of cuDNN."""
        if not (x.is_cuda and x.dtype == torch.float32 and len(self.shortcut) == 1):
            return False
        if not (isinstance(self.bn1, Identity) and isinstance(self.bn2, Identity)):
            return False
        convs = (self.conv1, self.conv2, self.shortcut[0])
        if self.act not in (F.gelu, F.relu) or any(type(c) is not nn.Conv2d for c in convs):
            return False
        if any(hasattr(c, "weight_orig") or hasattr(c, "weight_g") for c in convs):
            return False
        B, C, H, W = x.shape
        return (self.conv1.stride == (2, 2) and self.conv1.out_channels == 2 * C and H % 2 == 0 and W % 2 == 0
                and C % 4 == 0 and self.shortcut[0].stride == (2, 2))

    def forward(self, x):
        if self._fusable(x):
            return ode_block_integrate(x, self.conv1.weight, self.conv2.weight, _EULER_STEP, (0.0, 1.0),
                                       rhs_kind=_cabi.RHS_PREACT_NF, act=_act_code(self.act))
        if self._fusable_down(x):
            return resblock_down(x, self.conv1.weight, self.conv2.weight, self.shortcut[0].weight,
                                 act=_act_code(self.act))
        require_fallback("this pre-activation residual block (%s)" % (tuple(x.shape),),
                         "fused: CUDA fp32, Identity norm, gelu / relu, plain nn.Conv2d, tcgen05-tiled shape or stride-2 C -> 2C")
        out = self.conv1(self.act(self.bn1(x)))
        out = self.conv2(self.act(self.bn2(out)))
        return out + self.shortcut(x)


class _FusedRhs(nn.Module):
    """Common part of the two ODE right-hand sides: parameters + the description the kernels need."""
    expansion = 1
    rhs_kind = None

    def __init__(self, dim, norm_layer=None, act_layer=None, param_norm=lambda x: x):
        super().__init__()
        self.nfe = 0
        self._build(dim, norm_layer, param_norm)
        self.act = act_layer
        self.shortcut = nn.Sequential()

    def _group_norm_params(self, bn, tag):
        """(weight, bias, groups, eps) of a per-sample normalisation: nn.GroupNorm ('GN', 'LN' = one group) or a plain
        nn.InstanceNorm2d ('IN' = one group per channel, no affine parameters); None for anything else."""
        if isinstance(bn, nn.GroupNorm) and bn.affine:
            return bn.weight, bn.bias, bn.num_groups, bn.eps
        if isinstance(bn, nn.InstanceNorm2d) and not bn.affine and not bn.track_running_stats:
            ones = getattr(self, "_in_ones_" + tag, None)
            if ones is None or ones.device != self.conv1.weight.device:
                ones = torch.ones(bn.num_features, device=self.conv1.weight.device)
                setattr(self, "_in_ones_" + tag, ones)
            return ones, torch.zeros_like(ones), bn.num_features, bn.eps
        return None

    def fused_rhs_spec(self):
        for conv in (self.conv1, self.conv2):
            if type(conv) is not nn.Conv2d or hasattr(conv, "weight_orig") or hasattr(conv, "weight_g"):
                raise NotImplementedError("metasolver_b200: weight-normalised ODE-block convolutions are not "
                                          "implemented on the fused path (published config is 'PNF')")
        if isinstance(self.bn1, Identity) and isinstance(self.bn2, Identity):
            return dict(rhs_kind=self.rhs_kind, act=_act_code(self.act), w1=self.conv1.weight, w2=self.conv2.weight)
        n1, n2 = self._group_norm_params(self.bn1, "1"), self._group_norm_params(self.bn2, "2")
        if n1 is None or n2 is None or n1[2:] != n2[2:]:
            raise NotImplementedError("metasolver_b200: ODE-block normalisation %s / %s is not implemented on the fused "
                                      "path (implemented: 'NF', and the per-sample 'GN' / 'LN' / 'IN'; batch statistics "
                                      "would couple the samples of a batch)" % (type(self.bn1).__name__, type(self.bn2).__name__))
        kind = _cabi.RHS_PREACT_GN if self.rhs_kind == _cabi.RHS_PREACT_NF else _cabi.RHS_POSTACT_GN
        return dict(rhs_kind=kind, act=_act_code(self.act), groups=n1[2], eps=n1[3],
                    params=dict(norm1_w=n1[0], norm1_b=n1[1], norm2_w=n2[0], norm2_b=n2[1],
                                conv1_w=self.conv1.weight, conv2_w=self.conv2.weight))

    def forward(self, t, x, ss_loss=False):
        raise RuntimeError("metasolver_b200: ODE right-hand sides are evaluated inside the fused CUDA kernels; "
                           "calling the module directly is not supported (no unfused path)")


class BasicBlock2(_FusedRhs):
    """RHS act(bn2(conv2(act(bn1(conv1(x)))))) (layers.py:84-121)."""
    rhs_kind = _cabi.RHS_POSTACT_NF

    def _build(self, dim, norm_layer, param_norm):
        self.conv1 = param_norm(_conv3x3(dim, dim))
        self.bn1 = norm_layer(dim)
        self.conv2 = param_norm(_conv3x3(dim, dim))
        self.bn2 = norm_layer(dim)


class PreBasicBlock2(_FusedRhs):
    """RHS conv2(act(bn2(conv1(act(bn1(x)))))) (layers.py:124-161) -- the premetanode10 right-hand side."""
    rhs_kind = _cabi.RHS_PREACT_NF

    def _build(self, dim, norm_layer, param_norm):
        self.bn1 = norm_layer(dim)
        self.conv1 = param_norm(_conv3x3(dim, dim))
        self.bn2 = norm_layer(dim)
        self.conv2 = param_norm(_conv3x3(dim, dim))


class MetaODEBlock(nn.Module):
    """Regime dispatch around solver.integrate (layers.py:164-207): standalone / switch / ensemble."""

    def __init__(self, odefunc=None):
        super().__init__()
        self.rhs_func = odefunc
        self.integration_time = torch.tensor([0, 1]).float()

    def forward(self, x, solvers, solver_options):
        t = self.integration_time
        mode = solver_options.solver_mode
        n = len(solvers)
        if mode == 'standalone':
            return solvers[0].integrate_end(self.rhs_func, x, t)
        elif mode == 'switch':
            probs = solver_options.switch_probs
            if probs is None:
                probs = [1. / n for _ in range(n)]
            solver_id = np.random.choice(range(n), p=probs)          # numpy global RNG, drawn per block call
            solver_options.switch_solver_id = solver_id
            return solvers[solver_id].integrate_end(self.rhs_func, x, t)
        elif mode == 'ensemble':
            coin_flip = torch.bernoulli(torch.tensor((1,)), solver_options.ensemble_prob)
            solver_options.ensemble_coin_flip = coin_flip
            if coin_flip:
                weights = solver_options.ensemble_weights
                if weights is None:
                    weights = [1. / n for _ in range(n)]
                y = None
                if can_stack(solvers, self.rhs_func, t):
                    # stacked solver axis: the N solvers' stage evaluations are ONE batched set of launches
                    ys = integrate_stacked(solvers, self.rhs_func, x, t)
                    for k, wi in enumerate(weights):
                        yi = wi * ys[k]
                        y = yi if y is None else y + yi
                    return y
                for wi, solver in zip(weights, solvers):
                    yi = wi * solver.integrate_end(self.rhs_func, x, t)
                    y = yi if y is None else y + yi
                return y
            return solvers[0].integrate_end(self.rhs_func, x, t)
        elif mode == 'stacked':
            # extension (not in the reference): the batch holds len(solvers) equal slices, slice k is
            # integrated by solvers[k] -- K model copies that differ only in their solver, evaluated as one
            # model on a K-fold batch (model ensembling, MegaAdversarial/src/attacks/fgsm.py:135-143)
            return integrate_stacked(solvers, self.rhs_func, x, t, replicate=False)
        else:
            raise ValueError("unknown solver_mode %r" % (mode,))

    def ss_loss(self, y, solvers, solver_options):
        raise NotImplementedError("metasolver_b200: the steady-state regulariser is not implemented "
                                  "(the reference's CIFAR version raises NameError, layers.py:211)")


class MetaLayer(nn.Module):
    """`num_blocks = (n_res, n_ode)` residual blocks followed by ODE blocks (layers.py:252-314)."""

    def __init__(self, planes, num_blocks, stride, norm_layers_, param_norm_layers_, act_layers_, in_planes,
                 resblock=None, odefunc=None):
        super().__init__()
        n_res, n_ode = num_blocks
        self.in_planes = in_planes
        res = []
        for k in range(n_res):
            res.append(resblock(self.in_planes, planes, stride if k == 0 else 1, norm_layer=norm_layers_[0],
                                param_norm=param_norm_layers_[0], act_layer=act_layers_[0]))
            self.in_planes = planes * resblock.expansion
        ode = [MetaODEBlock(odefunc(self.in_planes, norm_layer=norm_layers_[1], param_norm=param_norm_layers_[1],
                                    act_layer=act_layers_[1])) for _ in range(n_ode)]
        self.blocks_res = nn.Sequential(*res)
        self.blocks_ode = nn.ModuleList(ode)

    def forward(self, x, solvers=None, solver_options=None, loss_options=None):
        x = self.blocks_res(x)
        self.ss_loss = 0
        for block in self.blocks_ode:
            x = block(x, solvers, solver_options)
            if (loss_options is not None) and loss_options.ss_loss:
                self.ss_loss += block.ss_loss(x, solvers, solver_options)
        return x

    def get_ss_loss(self):
        return self.ss_loss

    @property
    def nfe(self):
        return sum(block.rhs_func.nfe for block in self.blocks_ode)

    @nfe.setter
    def nfe(self, value):
        for block in self.blocks_ode:
            block.rhs_func.nfe = value


class MetaNODE(nn.Module):
    """Stem conv + up to four MetaLayers + avg-pool/FC head (layers.py:317-426)."""

    def __init__(self, num_blocks, num_classes=10, norm_layers_=(None, None, None),
                 param_norm_layers_=(lambda x: x, lambda x: x, lambda x: x), act_layers_=(None, None, None),
                 in_planes_=64, resblock=None, odefunc=None):
        super().__init__()
        self.n_layers = len(num_blocks)
        # The reference tests isinstance(<class>, PreBasicBlock), which is never true (layers.py:339-342):
        # the stem always applies act(bn1(.)) and nothing is applied before pooling.  Kept on purpose.
        self.is_preactivation = False
        self.conv1 = param_norm_layers_[2](_conv3x3(3, in_planes_))
        self.bn1 = norm_layers_[2](in_planes_)
        self.act = act_layers_[2]
        width, prev = in_planes_, in_planes_
        for li in range(self.n_layers):
            layer = MetaLayer(width, num_blocks[li], stride=1 if li == 0 else 2, norm_layers_=norm_layers_[:2],
                              param_norm_layers_=param_norm_layers_[:2], act_layers_=act_layers_[:2],
                              in_planes=prev, resblock=resblock, odefunc=odefunc)
            setattr(self, 'layer%d' % (li + 1), layer)
            prev = layer.in_planes
            self.n_features_linear = width
            width *= 2
        self.fc_layers = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), Flatten(),
                                       nn.Linear(self.n_features_linear * resblock.expansion, num_classes))

    def _fusable_stem(self, x):
        c = self.conv1
        return (x.is_cuda and x.dtype == torch.float32 and not self.is_preactivation and isinstance(self.bn1, Identity)
                and self.act in (F.gelu, F.relu) and type(c) is nn.Conv2d and not hasattr(c, "weight_orig")
                and not hasattr(c, "weight_g") and c.in_channels == 3 and c.out_channels % 64 == 0
                and c.out_channels <= 256 and c.stride == (1, 1))

    def _layers(self):
        return [getattr(self, 'layer%d' % i) for i in range(1, self.n_layers + 1)]

    @property
    def nfe(self):
        return sum(layer.nfe for layer in self._layers())

    @nfe.setter
    def nfe(self, value):
        for layer in self._layers():
            layer.nfe = value

    def forward(self, x, solvers=None, solver_options=None, loss_options=None):
        self.ss_loss = 0
        if self._fusable_stem(x):
            out = stem_conv_act(x, self.conv1.weight, _act_code(self.act))
        else:
            require_fallback("this stem (conv1 + norm + activation)", "fused: CUDA fp32 3-channel input, Identity norm, gelu / relu, "
                             "plain nn.Conv2d with 64 / 128 / 192 / 256 output channels")
            out = self.conv1(x)
            if not self.is_preactivation:
                out = self.act(self.bn1(out))
        for layer in self._layers():
            out = layer(out, solvers=solvers, solver_options=solver_options, loss_options=loss_options)
            self.ss_loss += layer.ss_loss
        if self.is_preactivation:
            out = self.act(self.bn1(out))
        return self._head(out)

    def _head(self, out):
        """AdaptiveAvgPool2d((1,1)) + Flatten + Linear (layers.py:390-392, 425) as one kernel of ours each way."""
        fc = self.fc_layers
        if (out.is_cuda and out.dtype == torch.float32 and out.dim() == 4 and out.shape[1] % 4 == 0 and len(fc) == 3
                and isinstance(fc[0], nn.AdaptiveAvgPool2d) and fc[0].output_size in ((1, 1), 1) and isinstance(fc[1], Flatten)
                and type(fc[2]) is nn.Linear and fc[2].in_features == out.shape[1]):
            return pool_fc(out, fc[2].weight, fc[2].bias)
        require_fallback("this network head", "fused: CUDA fp32 map, AdaptiveAvgPool2d((1,1)) + Flatten + nn.Linear")
        return self.fc_layers(out)


_DEPTHS = {4: ([(0, 1)], [(1, 0)]), 6: ([(1, 1)], [(2, 0)]), 10: ([(1, 1)] * 2, [(2, 0)] * 2),
           18: ([(1, 1)] * 4, [(2, 0)] * 4), 34: ([(1, 2), (1, 3), (1, 5), (1, 2)], [(3, 0), (4, 0), (6, 0), (3, 0)])}


def _factory(depth, pre):
    res, ode = (PreBasicBlock, PreBasicBlock2) if pre else (BasicBlock, BasicBlock2)

    def make(norm_layers, param_norm_layers, act_layers, in_planes, is_odenet=True):
        blocks = _DEPTHS[depth][0 if is_odenet else 1]
        return MetaNODE(list(blocks), norm_layers_=norm_layers, param_norm_layers_=param_norm_layers,
                        act_layers_=act_layers, in_planes_=in_planes, resblock=res, odefunc=ode)
    make.__name__ = ('premetanode%d' if pre else 'metanode%d') % depth
    make.__doc__ = "layers.py:429-556 factory: %s" % make.__name__
    return make


metanode4, metanode6, metanode10, metanode18, metanode34 = (_factory(d, False) for d in (4, 6, 10, 18, 34))
premetanode4, premetanode6, premetanode10, premetanode18, premetanode34 = (_factory(d, True) for d in (4, 6, 10, 18, 34))
