"""Oracle: the attacks that consume the hot path's input gradients, as plain functions on a
`model(x) -> logits` callable.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates MegaAdversarial/src/attacks/fgsm.py:21-43 (FGSM), :88-106 (FGSMRandom),
pgd.py:23-57 (PGD) and the FGSM-random training step of examples/cifar10/train_and_attack.py:246-327
(with --opt-level O0 / loss-scale 1.0 apex.amp is an fp32 pass-through, :114-117).
Normalisation follows torchvision.transforms.Normalize: (x - mean[:,None,None]) / std[:,None,None].
"""
import torch
import torch.nn.functional as F


def _ch(v, x):
    return torch.as_tensor(list(v), dtype=x.dtype, device=x.device).view(1, -1, 1, 1)


def normalize(x, mean, std):
    return (x - _ch(mean, x)) / _ch(std, x)


def unnormalize(x, mean, std):
    return (x - _ch([-m / s for m, s in zip(mean, std)], x)) / _ch([1 / s for s in std], x)


def fgsm(model, x, y, eps, mean, std):
    x01 = unnormalize(x, mean, std)
    xa = x01.clone().detach().requires_grad_(True)
    loss = F.cross_entropy(model(normalize(xa, mean, std)), y)
    grad = torch.autograd.grad([loss], [xa])[0]
    xa = torch.clamp(xa + eps * grad.sign(), 0, 1)
    return normalize(xa, mean, std).detach()


def _clamp(X, lo, hi):
    if not isinstance(hi, torch.Tensor):
        hi = torch.tensor(hi, device=X.device, dtype=X.dtype)
    if not isinstance(lo, torch.Tensor):
        lo = torch.tensor(lo, device=X.device, dtype=X.dtype)
    return torch.max(torch.min(X, hi), lo)


def fgsm_random(model, x, y, alpha, epsilon, mu, std, u01):
    """`u01` = the U[0,1) draw (torch.rand_like(x) in the reference).  Calls loss.backward():
    gradients of `model`'s parameters accumulate, exactly as in fgsm.py:98."""
    mu_t, std_t = _ch(mu, x), _ch(std, x)
    lower, upper = (0. - mu_t) / std_t, (1. - mu_t) / std_t
    eps_t, alpha_t = epsilon / std_t, alpha / std_t
    delta = eps_t - (2 * eps_t) * u01
    delta = _clamp(delta, lower - x, upper - x).detach().requires_grad_(True)
    loss = F.cross_entropy(model(x + delta), y)
    loss.backward()
    grad = delta.grad.detach()
    delta = _clamp(delta.detach() + alpha_t * torch.sign(grad), -eps_t, eps_t)
    delta = _clamp(delta, lower - x, upper - x).detach()
    return x + delta


def pgd(model, x, y, eps, lr, n_iter, mean, std, start_noise=None):
    x01 = unnormalize(x, mean, std)
    if start_noise is not None:
        xa = torch.clamp(x01 + start_noise, 0, 1).clone().detach()
    else:
        xa = x01.clone().detach()
    for i in range(n_iter):
        xa.requires_grad_(True)
        loss = F.cross_entropy(model(normalize(xa, mean, std)), y)
        grad = torch.autograd.grad([loss], [xa])[0]
        xa = torch.max(torch.min(xa + lr * grad.sign(), x01 + eps), x01 - eps)
        xa = torch.clamp(xa, 0, 1)
        if i == n_iter - 1:
            xa = normalize(xa, mean, std)
        xa = xa.detach()
    return xa


def fgsm_2ensemble(models, x, y, eps, mean, std):
    """FGSM2Ensemble, fgsm.py:128-155: FGSM on NLL(log(mean_i softmax(model_i(x)))).
    `models` = list of callables x -> logits (same network, different solvers, in the reference's use)."""
    x01 = unnormalize(x, mean, std)
    xa = x01.clone().detach().requires_grad_(True)
    probs = 0
    for m in models:
        probs = probs + torch.softmax(m(normalize(xa, mean, std)), dim=1)
    probs = probs / len(models)
    loss = F.nll_loss(torch.log(probs), y)
    grad = torch.autograd.grad([loss], [xa])[0]
    xa = torch.clamp(xa + eps * grad.sign(), 0, 1)
    return normalize(xa, mean, std).detach()
