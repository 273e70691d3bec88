"""MNIST model of the `sopa` API (sopa/src/models/odenet_mnist/layers.py).  The ODE block
(MetaODEBlock + ODEfunc + ConcatConv2d, :8-50, :134-171, :240-253) runs in the fused CUDA path;
the down-sampling stem, head and ResBlock (:173-237, tiny) stay PyTorch.  Names match the reference
so its checkpoints (examples/mnist/checkpoints/checkpoint_15444.pth state) load unchanged."""
import torch
import torch.nn as nn

from ..odenet_cifar10.layers import MetaODEBlock as _RegimeDispatch, Flatten
from ..... import _cabi


def norm(dim):
    return nn.GroupNorm(min(32, dim), dim)


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)


class ConcatConv2d(nn.Module):
    """Convolution over cat([t * ones, x]) (:240-253).  Only the parameter container is used here:
    the fused kernels fold the time plane into a per-pixel bias."""

    def __init__(self, dim_in, dim_out, ksize=3, stride=1, padding=0, dilation=1, groups=1, bias=True, transpose=False):
        super().__init__()
        if transpose or ksize != 3 or stride != 1 or padding != 1 or dilation != 1 or groups != 1 or not bias:
            raise NotImplementedError("metasolver_b200: only the 3x3 / stride 1 / pad 1 / biased ConcatConv2d of the "
                                      "reference's ODEfunc is implemented")
        self._layer = nn.Conv2d(dim_in + 1, dim_out, kernel_size=3, stride=1, padding=1, bias=True)


class ODEfunc(nn.Module):
    """GN -> ReLU -> ConcatConv -> GN -> ReLU -> ConcatConv -> GN (:134-171).  As in the reference the
    activation argument is validated but the block always uses ReLU (:139-151)."""

    def __init__(self, dim, activation_type='relu'):
        super().__init__()
        if activation_type not in ('tanh', 'softplus', 'softsign', 'relu'):
            raise NotImplementedError('{} activation is not implemented'.format(activation_type))
        self.norm1 = norm(dim)
        self.relu = nn.ReLU(inplace=True)
        self.conv1 = ConcatConv2d(dim, dim, 3, 1, 1)
        self.norm2 = norm(dim)
        self.conv2 = ConcatConv2d(dim, dim, 3, 1, 1)
        self.norm3 = norm(dim)
        self.nfe = 0

    def fused_rhs_spec(self):
        p = dict(norm1_w=self.norm1.weight, norm1_b=self.norm1.bias, norm2_w=self.norm2.weight,
                 norm2_b=self.norm2.bias, norm3_w=self.norm3.weight, norm3_b=self.norm3.bias,
                 conv1_w=self.conv1._layer.weight, conv1_b=self.conv1._layer.bias,
                 conv2_w=self.conv2._layer.weight, conv2_b=self.conv2._layer.bias)
        return dict(rhs_kind=_cabi.RHS_MNIST_GN_T, act=_cabi.ACT_RELU, params=p, groups=self.norm1.num_groups,
                    eps=self.norm1.eps)

    def forward(self, t, x, ss_loss=False):
        raise RuntimeError("metasolver_b200: ODE right-hand sides are evaluated inside the fused CUDA kernels; "
                           "calling the module directly is not supported (no unfused path)")


class MetaODEBlock(_RegimeDispatch):
    """Same regime dispatch as the CIFAR block (:16-50); builds its own ODEfunc(64) (:9-13)."""

    def __init__(self, activation_type='relu'):
        super().__init__(ODEfunc(64, activation_type))

    def ss_loss(self, y, solvers, solver_options):
        """Steady-state regulariser (:53-93): integrate one more unit of time from the block output and penalise the
        mean per-sample distance travelled.  The reference passes `partial(self.rhs_func, ss_loss=True).func` to the
        solver -- `.func` is the un-wrapped module, so the `abs` branch of ODEfunc.forward (:168-169) is never taken:
        the plain right-hand side is integrated over t in [1, 2].  Reproduced as is; runs on the fused path."""
        z0 = y
        t_ss = self.integration_time + 1
        mode = solver_options.solver_mode
        n = len(solvers)
        if mode == 'standalone':
            z = solvers[0].integrate(self.rhs_func, y, t_ss)
        elif mode == 'switch':
            z = solvers[solver_options.switch_solver_id].integrate(self.rhs_func, y, t_ss)
        elif mode == 'ensemble':
            if solver_options.ensemble_coin_flip:
                weights = solver_options.ensemble_weights
                if weights is None:
                    weights = [1. / n for _ in range(n)]
                z = None
                for wi, solver in zip(weights, solvers):
                    zi = wi * solver.integrate(self.rhs_func, y, t_ss)
                    z = zi if z is None else z + zi
            else:
                z = solvers[0].integrate(self.rhs_func, y, t_ss)
        else:
            raise ValueError("unknown solver_mode %r" % (mode,))
        z = z[-1] - z0
        z = torch.norm(z.reshape((z.shape[0], -1)), dim=1)
        return torch.mean(z)


class ResBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.norm1 = norm(inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.norm2 = norm(planes)
        self.conv2 = conv3x3(planes, planes)

    def forward(self, x):
        shortcut = x
        out = self.relu(self.norm1(x))
        if self.downsample is not None:
            shortcut = self.downsample(out)
        out = self.conv2(self.relu(self.norm2(self.conv1(out))))
        return out + shortcut


def build_downsampling_layers(downsampling_method='conv', in_channels=1):
    if downsampling_method == 'conv':
        return [nn.Conv2d(in_channels, 64, 3, 1), norm(64), nn.ReLU(inplace=True), nn.Conv2d(64, 64, 4, 2, 1),
                norm(64), nn.ReLU(inplace=True), nn.Conv2d(64, 64, 4, 2, 1)]
    if downsampling_method == 'res':
        return [nn.Conv2d(in_channels, 64, 3, 1), ResBlock(64, 64, stride=2, downsample=conv1x1(64, 64, 2)),
                ResBlock(64, 64, stride=2, downsample=conv1x1(64, 64, 2))]
    raise ValueError(downsampling_method)


def build_fc_layers():
    return [norm(64), nn.ReLU(inplace=True), nn.AdaptiveAvgPool2d((1, 1)), Flatten(), nn.Linear(64, 10)]


class MetaNODE(nn.Module):
    def __init__(self, downsampling_method='conv', is_odenet=True, activation_type='relu', in_channels=1):
        super().__init__()
        self.is_odenet = is_odenet
        self.downsampling_layers = nn.Sequential(*build_downsampling_layers(downsampling_method, in_channels))
        self.fc_layers = nn.Sequential(*build_fc_layers())
        if is_odenet:
            self.blocks = nn.ModuleList([MetaODEBlock(activation_type)])
        else:
            self.blocks = nn.ModuleList([ResBlock(64, 64) for _ in range(6)])

    def forward(self, x, solvers=None, solver_options=None, loss_options=None):
        self.ss_loss = 0
        x = self.downsampling_layers(x)
        for block in self.blocks:
            if self.is_odenet:
                x = block(x, solvers, solver_options)
                if (loss_options is not None) and loss_options.ss_loss:          # :117-122
                    self.ss_loss += block.ss_loss(x, solvers, solver_options)
            else:
                x = block(x)
        return self.fc_layers(x)

    def get_ss_loss(self):
        return self.ss_loss
