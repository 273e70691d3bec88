"""FGSM, FGSM-random, PGD and the 2-model-ensemble FGSM (MegaAdversarial/src/attacks/{attack,base,fgsm,pgd}.py).

Differences from the reference, none of which changes results:
  * no torchvision dependency -- (inverse) normalisation is the same sub/div arithmetic on fp32 tensors;
  * tensors follow `x.device` instead of a module-level global device;
  * every randomised attack takes an optional `noise=` tensor so a test (or a multi-GPU run that must
    reproduce single-process semantics) can supply host-generated randomness; without it the same
    torch RNG calls as the reference are made;
  * FGSM / PGD / FGSM2Ensemble differentiate w.r.t. the input only (pgd.py:44-46, fgsm.py:34-36), so
    the ODE-block backward is run in input-gradient-only mode (`metasolver_b200.input_grad_only`).
    FGSMRandom keeps `loss.backward()` (fgsm.py:98): parameter gradients DO accumulate there, which
    the published training recipe relies on (examples/cifar10/train_and_attack.py:258-259, 290).
"""
import torch
import torch.nn as nn

from .... import _cabi
from ....ops import input_grad_only
from ....train_ops import attack_step


def _fusable(x):
    """The elementwise attack steps run as one CUDA kernel each on CUDA fp32 image batches."""
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] <= _cabi.ATTACK_MAX_CHANNELS


def _f32(vals):
    """Python floats -> the fp32 values torch.as_tensor(..., dtype=float32) would hold."""
    return torch.tensor([float(v) for v in vals], dtype=torch.float32).tolist()


class _NormalizeFn(torch.autograd.Function):
    """(x - mean) / std as one kernel, differentiable w.r.t. x (grad / std, as autograd derives for the reference)."""

    @staticmethod
    def forward(ctx, x, mean, std):
        ctx.std = std
        return attack_step(_cabi.ATTACK_NORMALIZE, x.detach(), chan_consts=[mean, std])

    @staticmethod
    def backward(ctx, g):
        return g / _chan(ctx.std, g), None, None


_CHAN_CACHE = {}


def _chan(vals, like):
    """Per-channel constant as a (1,C,1,1) tensor on `like`'s device.  Cached: the attacks call this several
    times per iteration, and a fresh host->device copy each time would serialise the loop on the host (and
    cannot be captured into a CUDA graph)."""
    key = (tuple(float(v) for v in vals), like.dtype, like.device)
    t = _CHAN_CACHE.get(key)
    if t is None:
        t = torch.as_tensor(list(vals), dtype=like.dtype, device=like.device).view(1, -1, 1, 1)
        _CHAN_CACHE[key] = t
    return t


class _Normalizer:
    def __init__(self, mean, std):
        self.mean = tuple(mean) if mean is not None else (0., 0., 0.)
        self.std = tuple(std) if std is not None else (1., 1., 1.)

    def normalize(self, x):
        if _fusable(x):
            return _NormalizeFn.apply(x, _f32(self.mean), _f32(self.std))
        return (x - _chan(self.mean, x)) / _chan(self.std, x)

    def unnormalize(self, x):
        # transforms.Normalize(mean=[-m/s], std=[1/s]) of fgsm.py:27 / pgd.py:28
        inv_mean = [-m / s for m, s in zip(self.mean, self.std)]
        inv_std = [1 / s for s in self.std]
        if _fusable(x) and not x.requires_grad:
            return attack_step(_cabi.ATTACK_UNNORMALIZE, x, chan_consts=[_f32(inv_mean), _f32(inv_std)])
        return (x - _chan(inv_mean, x)) / _chan(inv_std, x)

    def consts(self):
        return [_f32(self.mean), _f32(self.std)]


class Attack(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model

    def _project(self, x):
        return torch.clamp(x, 0, 1)

    def _clamp(self, x, min, max):
        return torch.max(torch.min(x, max), min)

    def _eval_mode(self, models):
        was_training = models[0].training
        if was_training:
            for m in models:
                m.eval()
        return was_training

    def forward(self, *args, **kwargs):
        raise NotImplementedError


class Attack2Ensemble(Attack):
    def __init__(self, models):
        nn.Module.__init__(self)
        self.models = models


class Clean(Attack):
    def forward(self, x, y, kwargs):
        return x, y


class Clean2Ensemble(Attack2Ensemble):
    def forward(self, x, y, kwargs_arr):
        return x, y


def _input_gradient(loss, x):
    with input_grad_only():
        return torch.autograd.grad([loss], [x], create_graph=False, retain_graph=False)[0]


class FGSM(Attack):
    """Single signed-gradient step of size eps in [0,1] image space (fgsm.py:8-43)."""

    def __init__(self, model, eps=None, mean=None, std=None):
        super().__init__(model)
        self.eps = eps
        self.loss_fn = nn.CrossEntropyLoss()
        self.norm = _Normalizer(mean, std)

    def forward(self, x, y, kwargs):
        was_training = self._eval_mode([self.model])
        x01 = self.norm.unnormalize(x)
        xa = x01.clone().detach().requires_grad_(True)
        loss = self.loss_fn(self.model(self.norm.normalize(xa), **kwargs), y)
        grad = _input_gradient(loss, xa)
        if _fusable(xa):
            xa = attack_step(_cabi.ATTACK_FGSM_STEP, xa.detach(), grad=grad, eps=self.eps, normalize_out=True,
                             chan_consts=self.norm.consts())
        else:
            xa = self.norm.normalize(self._project(xa + self.eps * grad.sign())).detach()
        if was_training:
            self.model.train()
        return xa, y


def clamp(X, lower_limit, upper_limit):
    if not isinstance(upper_limit, torch.Tensor):
        upper_limit = torch.tensor(upper_limit, device=X.device, dtype=X.dtype)
    if not isinstance(lower_limit, torch.Tensor):
        lower_limit = torch.tensor(lower_limit, device=X.device, dtype=X.dtype)
    return torch.max(torch.min(X, upper_limit), lower_limit)


class FGSMRandom(Attack):
    """Random start in the eps-ball + one signed step alpha, in normalised space (fgsm.py:54-106)."""

    def __init__(self, model, alpha, epsilon=None, mu=None, std=None):
        super().__init__(model)
        self.scaled = (mu is not None) and (std is not None)
        self.mu, self.std_, self.alpha_, self.epsilon_ = mu, std, alpha, epsilon
        self.loss_fn = nn.CrossEntropyLoss()

    def _limits(self, x):
        if self.scaled:
            mu, std = _chan(self.mu, x), _chan(self.std_, x)
            return (0. - mu) / std, (1. - mu) / std, self.epsilon_ / std, self.alpha_ / std
        return 0., 1., self.epsilon_, self.alpha_

    def forward(self, x, y, kwargs, noise=None):
        was_training = self._eval_mode([self.model])
        u01 = torch.rand_like(x) if noise is None else noise.to(x)
        fused = _fusable(x) and not x.requires_grad
        if fused:
            consts = self._host_limits(x.shape[1])
            delta = attack_step(_cabi.ATTACK_FGSMR_INIT, u01, ref=x, chan_consts=consts).requires_grad_(True)
        else:
            lower, upper, epsilon, alpha = self._limits(x)
            delta = epsilon - (2 * epsilon) * u01                        # Uniform[-eps, eps]
            delta = clamp(delta, lower - x, upper - x).detach().requires_grad_(True)
        loss = self.loss_fn(self.model(x + delta, **kwargs), y)
        loss.backward()                                              # parameter grads accumulate too
        grad = delta.grad.detach()
        if fused:
            xa = attack_step(_cabi.ATTACK_FGSMR_STEP, delta.detach(), grad=grad, ref=x, normalize_out=True, chan_consts=consts)
        else:
            delta = clamp(delta.detach() + alpha * torch.sign(grad), -epsilon, epsilon)
            delta = clamp(delta, lower - x, upper - x).detach()
            xa = x + delta
        if was_training:
            self.model.train()
        return xa, y

    def _host_limits(self, C):
        """[lower, upper, eps, alpha] per channel as the fp32 values `_limits` computes on the device."""
        if self.scaled:
            mu = torch.tensor(list(self.mu), dtype=torch.float32)
            std = torch.tensor(list(self.std_), dtype=torch.float32)
            return [((0. - mu) / std).tolist(), ((1. - mu) / std).tolist(), (self.epsilon_ / std).tolist(),
                    (self.alpha_ / std).tolist()]
        return [[0.] * C, [1.] * C, _f32([self.epsilon_]) * C, _f32([self.alpha_]) * C]


class PGD(Attack):
    """n_iter signed steps of size lr, clamped to the eps-ball around x and to [0,1] (pgd.py:8-57)."""

    def __init__(self, model, eps=None, lr=None, n_iter=None, randomized_start=True, mean=None, std=None):
        super().__init__(model)
        self.eps, self.lr, self.n_iter, self.randomized_start = eps, lr, n_iter, randomized_start
        self.loss_fn = nn.CrossEntropyLoss()
        self.norm = _Normalizer(mean, std)

    def forward(self, x, y, kwargs, noise=None):
        was_training = self._eval_mode([self.model])
        x01 = self.norm.unnormalize(x)
        if self.randomized_start:
            start = torch.zeros_like(x01).uniform_(-self.eps, self.eps) if noise is None else noise.to(x01)
            xa = self._project(x01 + start).clone().detach()
        else:
            xa = x01.clone().detach()
        for i in range(self.n_iter):
            xa.requires_grad_(True)
            loss = self.loss_fn(self.model(self.norm.normalize(xa), **kwargs), y)
            grad = _input_gradient(loss, xa)
            if _fusable(xa):
                xa = attack_step(_cabi.ATTACK_PGD_STEP, xa.detach(), grad=grad, ref=x01, eps=self.eps, step=self.lr,
                                 normalize_out=(i == self.n_iter - 1), chan_consts=self.norm.consts())
                continue
            xa = self._project(self._clamp(xa + self.lr * grad.sign(), x01 - self.eps, x01 + self.eps))
            if i == self.n_iter - 1:
                xa = self.norm.normalize(xa)
            xa = xa.detach()
        if was_training:
            self.model.train()
        return xa, y


def ensemble_logits(models, x, kwargs_arr):
    """[model_i(x, **kwargs_i)] for the model-ensembling loop of fgsm.py:135-137.

    When every entry is the SAME network in the standalone regime and only the solver differs (how the
    reference's notebooks build solver ensembles of one trained model), the K forward passes are folded into
    ONE pass over a K-fold batch with a stacked solver axis in every ODE block (`solver_mode='stacked'`):
    slice k of the batch is integrated by solver k inside the same kernel launches."""
    from argparse import Namespace
    from ....sopa.src.solvers.rk_parametric import RKParametricSolver
    K = len(models)
    same_model = all(m is models[0] for m in models)
    solvers = []
    for kw in kwargs_arr:
        so, sv = kw.get("solver_options"), kw.get("solvers")
        if (set(kw) != {"solvers", "solver_options"} or getattr(so, "solver_mode", None) != "standalone"
                or not sv or not isinstance(sv[0], RKParametricSolver)):
            solvers = None
            break
        solvers.append(sv[0])
    stackable = (same_model and solvers is not None and 2 <= K <= 8 and x.is_cuda
                 and not (torch.is_grad_enabled() and any(s._params_need_grad() for s in solvers))
                 and all(s.n_stages == solvers[0].n_stages for s in solvers)
                 and all(torch.equal(s.host_time_grid(torch.tensor([0., 1.])), solvers[0].host_time_grid(torch.tensor([0., 1.])))
                         for s in solvers))
    if not stackable:
        return [m(x, **kw) for m, kw in zip(models, kwargs_arr)]
    xs = x.unsqueeze(0).expand(K, *x.shape).reshape(K * x.shape[0], *x.shape[1:])
    logits = models[0](xs, solvers=solvers, solver_options=Namespace(solver_mode="stacked"))
    return list(logits.view(K, x.shape[0], -1).unbind(0))


class FGSM2Ensemble(Attack2Ensemble):
    """FGSM on the NLL of the averaged softmax of several models / solvers (fgsm.py:109-155)."""

    def __init__(self, models, eps=None, mean=None, std=None):
        super().__init__(models)
        self.eps = eps
        self.loss_fn = nn.NLLLoss()
        self.norm = _Normalizer(mean, std)

    def forward(self, x, y, kwargs_arr):
        was_training = self._eval_mode(list(self.models))
        x01 = self.norm.unnormalize(x)
        xa = x01.clone().detach().requires_grad_(True)
        probs = 0
        for logits in ensemble_logits(list(self.models), self.norm.normalize(xa), kwargs_arr):
            probs = probs + torch.softmax(logits, dim=1)
        probs = probs / len(self.models)
        loss = self.loss_fn(torch.log(probs), y)
        grad = _input_gradient(loss, xa)
        if _fusable(xa):
            xa = attack_step(_cabi.ATTACK_FGSM_STEP, xa.detach(), grad=grad, eps=self.eps, normalize_out=True,
                             chan_consts=self.norm.consts())
        else:
            xa = self.norm.normalize(self._project(xa + self.eps * grad.sign())).detach()
        if was_training:
            for m in self.models:
                m.train()
        return xa, y
