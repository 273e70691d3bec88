"""GPU parity of BASELINE config 3: solver ensembling and model ensembling over 4 RK2 u values (and RK4)
on a stacked solver axis -- the N solvers' stage evaluations run in ONE batched set of launches.
Checked against golden vectors of the real reference (sequential loop over solvers) and, bit for bit,
against this library's own one-solver-at-a-time integration."""
import os
import sys
from argparse import Namespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden, max_rel, ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_cases as cases  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _block(C):
    import metasolver_b200  # noqa: F401
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    return MetaODEBlock(PreBasicBlock2(C, norm_layer=Identity, act_layer=F.gelu)).cuda()


def _solvers(svs):
    from metasolver_b200.sopa.src.solvers.utils import create_solver
    out = [create_solver(*sv, torch.float32, "cuda") for sv in svs]
    for s in out:
        s.freeze_params()
    return out


@pytest.mark.parametrize("engine", ["tcgen05", "simt"])
@pytest.mark.parametrize("tag,shape,svs,weights", [
    ("c64_rk2x4_uniform", (64, 8, 32, 2), cases.C3_RK2_SOLVERS, None),
    ("c64_rk2x4_weighted", (64, 8, 32, 2), cases.C3_RK2_SOLVERS, cases.C3_WEIGHTS),
    ("c64_rk4x2_uniform", (64, 8, 32, 2), cases.C3_RK4_SOLVERS, None),
    ("c128_rk2x4_uniform", (128, 8, 16, 1), cases.C3_RK2_SOLVERS, None)])
def test_solver_ensemble_stacked_vs_reference_golden(tag, shape, svs, weights, engine):
    import metasolver_b200 as msb
    g = golden("ensemble_c3.npz")
    C = shape[0]
    x, w1, w2, r = [torch.from_numpy(a).cuda() for a in cases.ode_case_inputs(*shape)]
    blk = _block(C)
    with torch.no_grad():
        blk.rhs_func.conv1.weight.copy_(w1)
        blk.rhs_func.conv2.weight.copy_(w2)
    solvers = _solvers(svs)
    msb.set_default_engine(engine)
    try:
        before = msb.launch_count()
        x.requires_grad_(True)
        y = blk(x, solvers, Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=weights))
        (y * r).sum().backward()
        torch.cuda.synchronize()
        stacked_launches = msb.launch_count() - before
        gw1 = blk.rhs_func.conv1.weight.grad.cpu().numpy().reshape(-1)[::cases.WG_STRIDE]
        gw2 = blk.rhs_func.conv2.weight.grad.cpu().numpy().reshape(-1)[::cases.WG_STRIDE]
        nfe = blk.rhs_func.nfe
        # one solver alone on the same batch: the stacked pass must not cost more launches than that
        before = msb.launch_count()
        x2 = x.detach().clone().requires_grad_(True)
        y1 = blk(x2, solvers[:1], Namespace(solver_mode="standalone"))
        (y1 * r).sum().backward()
        torch.cuda.synchronize()
        single_launches = msb.launch_count() - before
    finally:
        msb.set_default_engine("auto")
    assert stacked_launches == single_launches, (stacked_launches, single_launches)
    assert nfe == int(g[tag + "_nfe"])
    assert max_rel(y.detach().cpu().numpy(), g[tag + "_y"]) <= TOL
    assert max_rel(x.grad.cpu().numpy(), g[tag + "_gx"]) <= TOL
    assert max_rel(gw1, g[tag + "_gw1"]) <= TOL
    assert max_rel(gw2, g[tag + "_gw2"]) <= TOL


def test_stacked_slices_bit_identical_to_single_solver_runs():
    """slice k of the stacked integration == solvers[k].integrate alone (forward AND input gradient)."""
    from metasolver_b200.sopa.src.solvers.rk_parametric import integrate_stacked
    x, w1, w2, r = [torch.from_numpy(a).cuda() for a in cases.ode_case_inputs(64, 8, 32, 2)]
    blk = _block(64)
    with torch.no_grad():
        blk.rhs_func.conv1.weight.copy_(w1)
        blk.rhs_func.conv2.weight.copy_(w2)
    solvers = _solvers(cases.C3_RK2_SOLVERS)
    t = torch.tensor([0., 1.])
    xs = x.clone().requires_grad_(True)
    ys = integrate_stacked(solvers, blk.rhs_func, xs, t)
    assert ys.shape == (4,) + tuple(x.shape)
    gsum = None
    for k, s in enumerate(solvers):
        xk = x.clone().requires_grad_(True)
        yk = s.integrate(blk.rhs_func, xk, t)[-1]
        assert torch.equal(yk, ys[k]), k
        (yk * r).sum().backward()
        gsum = xk.grad.clone() if gsum is None else gsum + xk.grad
    (ys * r.unsqueeze(0)).sum().backward()
    assert max_rel(xs.grad.cpu().numpy(), gsum.cpu().numpy()) <= 1e-6
    # unequal stage counts / grids cannot share the axis: the regime falls back to the sequential loop
    mixed = _solvers([cases.C3_RK2_SOLVERS[0], cases.C3_RK4_SOLVERS[0]])
    with pytest.raises(ValueError):
        integrate_stacked(mixed, blk.rhs_func, x, t)
    with torch.no_grad():
        y = blk(x, mixed, Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=None))
        ref = 0.5 * mixed[0].integrate(blk.rhs_func, x, t)[-1] + 0.5 * mixed[1].integrate(blk.rhs_func, x, t)[-1]
    assert torch.equal(y, ref)


def _model():
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import premetanode10
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    from oracle.models import det_premetanode10_params
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = premetanode10((Identity,) * 3, (lambda x: x,) * 3, (F.gelu,) * 3, in_planes=64, is_odenet=True)
    model.load_state_dict(det_premetanode10_params())
    return model.cuda().eval()


def _images():
    from oracle import det_uniform
    from oracle.models import CIFAR_MEAN, CIFAR_STD
    img = torch.from_numpy(det_uniform((4, 3, 32, 32), 920, 0.0, 1.0))
    x = (img - torch.tensor(CIFAR_MEAN).view(1, 3, 1, 1)) / torch.tensor(CIFAR_STD).view(1, 3, 1, 1)
    return x.cuda(), torch.tensor([3, 1, 4, 1]).cuda(), CIFAR_MEAN, CIFAR_STD


def test_premetanode10_solver_ensembling_vs_reference_golden():
    g = golden("ensemble_c3.npz")
    model = _model()
    x, labels, _, _ = _images()
    solvers = _solvers(cases.C3_RK2_SOLVERS)
    x.requires_grad_(True)
    logits = model(x, solvers, Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=None))
    F.cross_entropy(logits, labels).backward()
    assert max_rel(logits.detach().cpu().numpy(), g["model_solver_ens_logits"]) <= TOL
    assert (logits.argmax(1).cpu().numpy() == g["model_solver_ens_logits"].argmax(1)).all()
    assert max_rel(x.grad.cpu().numpy(), g["model_solver_ens_gx"]) <= TOL
    params = dict(model.named_parameters())
    for k in g.files:
        if k.startswith("model_solver_ens_g_"):
            got = params[k[len("model_solver_ens_g_"):]].grad.cpu().numpy().reshape(-1)[::cases.WG_STRIDE]
            assert max_rel(got, g[k]) <= TOL, (k, max_rel(got, g[k]))
    assert model.nfe == 2 * 4 * 2 * 8


def test_model_ensembling_stacked_vs_reference_golden():
    """FGSM2Ensemble over one network under 4 solvers: ONE forward over a 4-fold batch (solver_mode='stacked')."""
    import metasolver_b200 as msb
    from metasolver_b200.MegaAdversarial.src.attacks import FGSM2Ensemble, ensemble_logits
    g = golden("ensemble_c3.npz")
    model = _model()
    x, labels, mean, std = _images()
    solvers = _solvers(cases.C3_RK2_SOLVERS)
    kwargs_arr = [{"solvers": [s], "solver_options": Namespace(solver_mode="standalone")} for s in solvers]
    with torch.no_grad():
        before = msb.launch_count()
        stacked = ensemble_logits([model] * 4, x, kwargs_arr)
        n_stacked = msb.launch_count() - before
        before = msb.launch_count()
        loop = [model(x, **kw) for kw in kwargs_arr]
        n_loop = msb.launch_count() - before
    assert n_stacked * 4 == n_loop, (n_stacked, n_loop)
    for a, b in zip(stacked, loop):
        assert max_rel(a.cpu().numpy(), b.cpu().numpy()) <= 1e-6     # non-ODE layers (cuDNN/cuBLAS) may pick other algos
    probs = sum(torch.softmax(l, dim=1) for l in stacked) / 4
    assert max_rel(probs.cpu().numpy(), g["model_ens_probs"]) <= TOL
    assert (probs.argmax(1).cpu().numpy() == g["model_ens_probs"].argmax(1)).all()
    xa, _ = FGSM2Ensemble([model] * 4, eps=8 / 255., mean=mean, std=std)(x, labels, kwargs_arr)
    diff = np.abs(xa.cpu().numpy() - g["model_ens_fgsm_x"])
    step = 8 / 255. / max(std)
    assert float((diff > 0.25 * step).mean()) <= 2e-3      # sign(grad) flips only where |grad| ~ rounding noise
