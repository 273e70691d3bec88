"""GPU: the fused callers either side of the path.

  * every fused attack step is BIT-IDENTICAL to the chain of torch calls the reference makes
    (MegaAdversarial/src/attacks/fgsm.py:27-40,93-105, pgd.py:28-53), in both memory formats, including zero gradients;
  * FusedSGD follows torch.optim.SGD (momentum 0.9, weight decay 5e-4: examples/cifar10/train_and_attack.py:98-99) and
    keeps parameters / gradients as views of its flat buffers.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

MEAN, STD = (0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)


def _chan(vals, like):
    return torch.as_tensor(list(vals), dtype=like.dtype, device=like.device).view(1, -1, 1, 1)


def _f32(vals):
    return torch.tensor([float(v) for v in vals], dtype=torch.float32).tolist()


def _clamp(x, lo, hi):
    return torch.max(torch.min(x, hi), lo)


@pytest.mark.parametrize("channels_last", [False, True])
def test_fused_attack_steps_bit_identical_to_torch_chain(channels_last):
    import metasolver_b200
    from metasolver_b200 import _cabi
    step = metasolver_b200.attack_step
    torch.manual_seed(0)
    B = 37
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    x01 = torch.rand(B, 3, 32, 32, device="cuda").contiguous(memory_format=fmt)
    x01[:, :, ::5, ::7] = 0.0                                                     # pixels exactly on the [0,1] box
    x01[:, :, 1::5, 2::7] = 1.0
    g = torch.randn_like(x01)
    g[g.abs() < 0.3] = 0.0                                                        # sign(0) = 0 must be reproduced
    eps, lr = 8 / 255, 2 / 255
    mean, std = _chan(MEAN, x01), _chan(STD, x01)
    xn = (x01 - mean) / std
    consts = [_f32(MEAN), _f32(STD)]
    # normalize / unnormalize
    assert torch.equal(step(_cabi.ATTACK_NORMALIZE, x01, chan_consts=consts), xn)
    inv_mean = [-m / s for m, s in zip(MEAN, STD)]
    inv_std = [1 / s for s in STD]
    ref = (xn - _chan(inv_mean, xn)) / _chan(inv_std, xn)
    assert torch.equal(step(_cabi.ATTACK_UNNORMALIZE, xn, chan_consts=[_f32(inv_mean), _f32(inv_std)]), ref)
    # FGSM
    xa = torch.clamp(x01 + 0.5 * eps * torch.randn_like(x01).sign(), 0, 1)
    ref = (torch.clamp(xa + eps * g.sign(), 0, 1) - mean) / std
    assert torch.equal(step(_cabi.ATTACK_FGSM_STEP, xa, grad=g, eps=eps, normalize_out=True, chan_consts=consts), ref)
    # PGD (inner and last iteration)
    inner = torch.clamp(_clamp(xa + lr * g.sign(), x01 - eps, x01 + eps), 0, 1)
    assert torch.equal(step(_cabi.ATTACK_PGD_STEP, xa, grad=g, ref=x01, eps=eps, step=lr, chan_consts=consts), inner)
    assert torch.equal(step(_cabi.ATTACK_PGD_STEP, xa, grad=g, ref=x01, eps=eps, step=lr, normalize_out=True,
                            chan_consts=consts), (inner - mean) / std)
    # FGSM-random in normalised space (per-channel limits) and in [0,1] space (scalar limits)
    u01 = torch.rand_like(x01)
    lower, upper, e_c, a_c = (0. - mean) / std, (1. - mean) / std, eps / std, (10 / 255) / std
    host = [t.reshape(-1).cpu().tolist() for t in (lower, upper, e_c, a_c)]
    d0 = _clamp(e_c - (2 * e_c) * u01, lower - xn, upper - xn)
    got0 = step(_cabi.ATTACK_FGSMR_INIT, u01, ref=xn, chan_consts=host)
    assert torch.equal(got0, d0)
    d1 = _clamp(_clamp(d0 + a_c * torch.sign(g), -e_c, e_c), lower - xn, upper - xn)
    assert torch.equal(step(_cabi.ATTACK_FGSMR_STEP, d0, grad=g, ref=xn, chan_consts=host), d1)
    assert torch.equal(step(_cabi.ATTACK_FGSMR_STEP, d0, grad=g, ref=xn, normalize_out=True, chan_consts=host), xn + d1)
    host = [[0.] * 3, [1.] * 3, _f32([eps]) * 3, _f32([10 / 255]) * 3]
    d0 = _clamp(eps - (2 * eps) * u01, 0. - x01, 1. - x01)
    assert torch.equal(step(_cabi.ATTACK_FGSMR_INIT, u01, ref=x01, chan_consts=host), d0)


def test_attack_step_rejects_bad_arguments():
    import metasolver_b200
    from metasolver_b200 import _cabi
    x = torch.rand(2, 3, 8, 8)
    with pytest.raises(RuntimeError):
        metasolver_b200.attack_step(_cabi.ATTACK_NORMALIZE, x, chan_consts=[[0.] * 3, [1.] * 3])      # CPU tensor
    xc = x.cuda()
    with pytest.raises(RuntimeError):
        metasolver_b200.attack_step(_cabi.ATTACK_PGD_STEP, xc, grad=xc)                              # ref missing
    with pytest.raises(RuntimeError):
        metasolver_b200.attack_step(17, xc)
    with pytest.raises(ValueError):
        metasolver_b200.attack_step(_cabi.ATTACK_NORMALIZE, torch.rand(2, 8, 4, 4, device="cuda"), chan_consts=[[0.], [1.]])
    assert metasolver_b200.attack_step(_cabi.ATTACK_NORMALIZE, torch.empty(0, 3, 8, 8, device="cuda"),
                                       chan_consts=[[0.] * 3, [1.] * 3]).numel() == 0


def test_fused_sgd_follows_torch_sgd():
    import metasolver_b200
    torch.manual_seed(1)
    shapes = [(64, 3, 3, 3), (64, 64, 3, 3), (128, 64, 1, 1), (10, 128), (10,)]
    ref_p = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    our_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    ref = torch.optim.SGD(ref_p, lr=0.05, momentum=0.9, weight_decay=5e-4)
    ours = metasolver_b200.FusedSGD(our_p, lr=0.05, momentum=0.9, weight_decay=5e-4)
    n = sum(p.numel() for p in our_p)
    assert ours.flat_param.numel() == n
    for it in range(4):
        ref.zero_grad()
        ours.zero_grad()
        for a, b in zip(ref_p, our_p):
            g = torch.randn_like(a)
            (a * g).sum().backward()
            (b * g).sum().backward()
            (b * g).sum().backward()                # accumulate twice (FGSM-random quirk), averaged away by grad_scale
        off = 0
        for b in our_p:                             # autograd accumulated IN PLACE into the flat buffer
            assert b.grad.data_ptr() == ours.flat_grad.data_ptr() + 4 * off
            assert b.data_ptr() == ours.flat_param.data_ptr() + 4 * off
            off += b.numel()
        if it == 2:
            ref.param_groups[0]["lr"] = ours.param_groups[0]["lr"] = 0.01       # scheduler-style lr change
        ref.step()
        ours.step(grad_scale=0.5)
        for a, b in zip(ref_p, our_p):
            torch.testing.assert_close(b.detach(), a.detach(), rtol=2e-6, atol=1e-7)
    with pytest.raises(RuntimeError):
        metasolver_b200.FusedSGD([torch.zeros(3, requires_grad=True)], lr=0.1)


def test_fused_sgd_sees_gradients_detached_by_zero_grad_set_to_none():
    """model.zero_grad(set_to_none=True) (the torch >= 2.0 default) detaches `.grad` from FusedSGD's flat buffer; step() /
    all_reduce() must gather the fresh gradients instead of applying stale zeros, and a parameter whose grad is None is
    skipped entirely (no weight decay, no momentum update) like torch.optim.SGD does."""
    import copy
    import torch.nn as nn
    from metasolver_b200 import FusedSGD
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(16, 32), nn.Linear(32, 8), nn.Linear(8, 4)).cuda()
    ref = copy.deepcopy(net)
    opt = FusedSGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    x = torch.randn(64, 16, device="cuda")
    for it in range(4):
        for m, o in ((net, opt), (ref, ropt)):
            m.zero_grad(set_to_none=True)                      # NOT FusedSGD.zero_grad: grads get detached from the flat buffer
            h = m[1](m[0](x))
            out = h.square().mean() if it == 2 else m[2](h).square().mean()      # iteration 2: the last layer gets no gradient
            out.backward()
            if o is opt:
                assert opt.all_reduce() == 1.0
            o.step()
        for a, b in zip(net.parameters(), ref.parameters()):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), it


def test_host_fed_loop_returns_every_result_in_order_one_call_late():
    import metasolver_b200 as msb
    seen = []

    def step(xd, yd):
        assert xd.is_cuda and yd.is_cuda
        seen.append(xd.data_ptr())
        return (xd * yd).sum()
    loop = msb.HostFedLoop(step, lag=1)
    n = 1 << 20                                         # large enough for the copies to take a while
    hosts = [(torch.full((n,), float(k)).pin_memory(), torch.full((n,), 0.5).pin_memory()) for k in range(6)]
    got = [loop(*h) for h in hosts]
    assert got[0] is None and got[1:] == [0.5 * n * k for k in range(5)]
    assert loop.drain() == [0.5 * n * 5] and loop.drain() == []
    assert len(set(seen)) == 2                          # two staging sets, alternating


def test_augment_normalize_kernel_matches_torchvision_bitwise():
    """sopa/src/models/odenet_cifar10/data.py:40-57: RandomCrop(32, padding=4) + RandomHorizontalFlip + ToTensor + Normalize
    per sample on the host -> one kernel over a uint8 dataset resident in HBM.  Same draws -> the same floats."""
    import metasolver_b200 as msb
    from metasolver_b200.sopa.src.models.odenet_cifar10.data import CIFAR_MEAN, CIFAR_STD
    tvf = pytest.importorskip("torchvision.transforms.functional")
    g = torch.Generator().manual_seed(5)
    N, B, pad = 40, 16, 4
    data = torch.randint(0, 256, (N, 32, 32, 3), generator=g, dtype=torch.uint8)
    index = torch.randperm(N, generator=g)[:B]
    dx = torch.randint(0, 2 * pad + 1, (B,), generator=g)
    dy = torch.randint(0, 2 * pad + 1, (B,), generator=g)
    flip = torch.rand(B, generator=g) < 0.5
    dx[0], dy[0], dx[1], dy[1] = 0, 0, 2 * pad, 2 * pad            # the extreme windows
    out = msb.augment_normalize(data.cuda(), index.cuda(), draws=(dx.cuda(), dy.cuda(), flip.cuda()), padding=pad)
    assert out.shape == (B, 3, 32, 32) and out.is_contiguous(memory_format=torch.channels_last)
    for b in range(B):
        img = data[index[b]].permute(2, 0, 1)                      # uint8 CHW, what PIL -> crop/flip sees
        img = tvf.pad(img, [pad, pad, pad, pad])                   # RandomCrop pads first (fill 0)
        img = tvf.crop(img, int(dy[b]), int(dx[b]), 32, 32)
        if bool(flip[b]):
            img = tvf.hflip(img)
        ref = tvf.normalize(img.float().div(255), CIFAR_MEAN, CIFAR_STD)     # ToTensor, Normalize
        assert torch.equal(out[b].cpu(), ref), b
    # evaluation transform: ToTensor + Normalize only
    ev = msb.augment_normalize(data.cuda(), train=False)
    ref = tvf.normalize(data.permute(0, 3, 1, 2).float().div(255), CIFAR_MEAN, CIFAR_STD)
    assert torch.equal(ev.cpu(), ref)
    # random draws on the device: deterministic under a seeded generator, every value a valid pixel or a padded one
    g1 = torch.Generator(device="cuda").manual_seed(3)
    g2 = torch.Generator(device="cuda").manual_seed(3)
    a1 = msb.augment_normalize(data.cuda(), generator=g1)
    a2 = msb.augment_normalize(data.cuda(), generator=g2)
    assert torch.equal(a1, a2) and not torch.equal(a1, ev)


def test_cyclic_lr_drives_fused_sgd_like_torch():
    """train_and_attack.py:480-505: SGD(momentum, weight_decay) + CyclicLR(cycle_momentum=True), 12 steps."""
    import metasolver_b200 as msb
    torch.manual_seed(0)
    w0 = torch.randn(4, 8, device="cuda")
    grads = [torch.randn(4, 8, device="cuda") for _ in range(12)]
    pt = torch.nn.Parameter(w0.clone())
    topt = torch.optim.SGD([pt], lr=0.1, momentum=0.9, weight_decay=5e-4)
    tsch = torch.optim.lr_scheduler.CyclicLR(topt, base_lr=1e-3, max_lr=0.1, step_size_up=4, mode="triangular2")
    pm = torch.nn.Parameter(w0.clone())
    mopt = msb.FusedSGD([pm], lr=0.1, momentum=0.9, weight_decay=5e-4)
    msch = msb.CyclicLR(mopt, base_lr=1e-3, max_lr=0.1, step_size_up=4, mode="triangular2")
    for gk in grads:
        pt.grad = gk.clone()
        pm.grad.copy_(gk)
        topt.step()
        tsch.step()
        mopt.step()
        msch.step()
        assert mopt.param_groups[0]["lr"] == topt.param_groups[0]["lr"]
    assert (pm.detach() - pt.detach()).abs().max().item() <= 1e-6 * pt.detach().abs().max().item()
