"""CPU: pin the oracle against golden vectors produced by the real reference (bit-exact fp32)."""
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden, max_rel
import oracle
from oracle import butcher_tableau, det_normal, integrate, rhs_preact, rhs_postact, rhs_mnist, RhsCounter
from oracle.models import det_premetanode10_params, premetanode10_forward, CIFAR_MEAN, CIFAR_STD

sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), "golden"))
import make_golden_cases as cases  # noqa: E402


@pytest.mark.parametrize("idx", range(len(cases.TABLEAU_CASES)))
def test_tableau_bit_exact(idx):
    g = golden("tableaus.npz")
    m, p, u0, v0 = cases.TABLEAU_CASES[idx]
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        npdt = np.float32 if dt == torch.float32 else np.float64
        u = None if m == "euler" else npdt(u0)
        v = None if v0 == -1 else npdt(v0)
        tab = butcher_tableau(m, p, u, v, dt)
        assert np.array_equal(np.array(tab["c"]), g["%d_%s_c" % (idx, tag)])
        assert np.array_equal(np.array(tab["b"]), g["%d_%s_b" % (idx, tag)])
        assert np.array_equal(np.array(tab["w"]), g["%d_%s_w" % (idx, tag)])


def test_known_answer_tableaus():
    # order2stage2.py:6-17 Midpoint / Heun; order4stage4.py:6-17 classical RK4 and 3/8 rule
    t = butcher_tableau("rk2", "u", np.float32(0.5), None)
    assert t["c"] == [0.0, 0.5] and t["b"] == [0.0, 1.0] and t["w"][1][0] == 0.5
    t = butcher_tableau("rk2", "u", np.float32(1.0), None)
    assert t["c"] == [0.0, 1.0] and t["b"] == [0.5, 0.5] and t["w"][1][0] == 1.0
    t = butcher_tableau("rk4", "u2", np.float32(1 / 3.), None)
    np.testing.assert_allclose(t["b"], [1 / 6., 1 / 3., 1 / 3., 1 / 6.], rtol=3e-7)
    np.testing.assert_allclose(t["w"], [[0, 0, 0, 0], [.5, 0, 0, 0], [0, .5, 0, 0], [0, 0, 1, 0]], atol=3e-7)
    t = butcher_tableau("rk4", "uv", np.float32(1 / 3.), np.float32(2 / 3.))
    np.testing.assert_allclose(t["b"], [1 / 8., 3 / 8., 3 / 8., 1 / 8.], rtol=2e-6)
    np.testing.assert_allclose(t["w"], [[0, 0, 0, 0], [1 / 3., 0, 0, 0], [-1 / 3., 1, 0, 0], [1, -1, 1, 0]], atol=3e-6)


def test_time_grids_bit_exact():
    g = golden("grids.npz")
    t01 = torch.tensor([0, 1]).float()
    for n in (1, 2, 3, 5, 7, 8, 10, 16):
        assert np.array_equal(oracle.make_time_grid(t01, n_steps=n).numpy(), g["n%d" % n])
    for ss in (0.3, 0.125, 0.4):
        assert np.array_equal(oracle.make_time_grid(t01, step_size=ss).numpy(), g["ss%g" % ss])


def _grid(sv):
    return dict(n_steps=sv[2]) if sv[2] != -1 else dict(step_size=sv[3])


@pytest.mark.parametrize("case", cases.ODE_CASES, ids=[c[0] for c in cases.ODE_CASES])
def test_ode_block_bit_exact(case):
    name, C, H, W, B, kind, sv = case
    g = golden("ode_%s.npz" % name)
    x, w1, w2, r = [torch.from_numpy(a) for a in cases.ode_case_inputs(C, H, W, B)]
    x.requires_grad_(True); w1.requires_grad_(True); w2.requires_grad_(True)
    cnt = RhsCounter()
    rhs = (rhs_preact if kind == "preact" else rhs_postact)(w1, w2, "gelu", cnt)
    tab = butcher_tableau(sv[0], sv[1], None if sv[0] == "euler" else np.float32(sv[4]),
                          None if sv[5] == -1 else np.float32(sv[5]))
    y = integrate(tab, rhs, x, torch.tensor([0, 1]).float(), **_grid(sv))[-1]
    (y * r).sum().backward()
    assert cnt.nfe == int(g["nfe"])
    assert np.array_equal(y.detach().numpy(), g["y"])
    assert np.array_equal(x.grad.numpy(), g["gx"])
    assert np.array_equal(w1.grad.numpy().reshape(-1)[::cases.WG_STRIDE], g["gw1"])
    assert np.array_equal(w2.grad.numpy().reshape(-1)[::cases.WG_STRIDE], g["gw2"])


def test_mnist_ode_block_trained_weights():
    w = golden("mnist_odeblock_weights.npz")
    g = golden("mnist_odeblock.npz")
    np.testing.assert_allclose(g["logits_first4"], [1.00175, 0.14910, 2.99485, -0.86763], atol=2e-5)
    for tag, sv in (("rk2_u05_n8", ("rk2", "u", 8, -1, 0.5, -1)), ("rk4_u2_n2", ("rk4", "u2", 2, -1, 1 / 3., -1)),
                    ("euler_n4", ("euler", None, 4, -1, -1, -1))):
        p = {k: torch.from_numpy(w[k]).requires_grad_(True) for k in w.files}
        x = torch.from_numpy(g["feat"]).requires_grad_(True)
        tab = butcher_tableau(sv[0], sv[1], None if sv[0] == "euler" else np.float32(sv[4]), None)
        y = integrate(tab, rhs_mnist(p), x, torch.tensor([0, 1]).float(), n_steps=sv[2])[-1]
        r = torch.from_numpy(det_normal(tuple(y.shape), 77))
        (y * r).sum().backward()
        assert np.array_equal(y.detach().numpy(), g[tag + "_y"])
        assert np.array_equal(x.grad.numpy(), g[tag + "_gx"])
        assert np.array_equal(p["conv1_w"].grad.numpy().reshape(-1)[::cases.WG_STRIDE], g[tag + "_gconv1_w"])
        assert np.array_equal(p["conv2_b"].grad.numpy(), g[tag + "_gconv2_b"])
        assert np.array_equal(p["norm1_w"].grad.numpy(), g[tag + "_gnorm1_w"])
        assert np.array_equal(p["norm3_b"].grad.numpy(), g[tag + "_gnorm3_b"])


def test_regimes_bit_exact():
    g = golden("regimes.npz")
    x, w1, w2, r = [torch.from_numpy(a) for a in cases.ode_case_inputs(64, 8, 32, 2)]
    rhs = rhs_preact(w1, w2)
    svs = cases.REGIME_SOLVERS
    tabs = [butcher_tableau(s[0], s[1], np.float32(s[4]), None) for s in svs]
    grids = [_grid(s) for s in svs]
    with torch.no_grad():
        np.random.seed(123)
        ids = []
        for rep in range(3):
            rec = {}
            y = oracle.ode_block_forward(x, rhs, tabs, grids, "switch", switch_probs=[0.1, 0.2, 0.3, 0.4], record=rec)
            ids.append(rec["switch_solver_id"])
            assert np.array_equal(y.numpy(), g["switch_y%d" % rep])
        assert ids == list(g["switch_ids"])
        torch.manual_seed(5)
        y = oracle.ode_block_forward(x, rhs, tabs, grids, "ensemble", ensemble_prob=1.0,
                                     ensemble_weights=[0.4, 0.3, 0.2, 0.1])
        assert np.array_equal(y.numpy(), g["ens_weighted_y"])
        y = oracle.ode_block_forward(x, rhs, tabs, grids, "ensemble", ensemble_prob=1.0)
        assert np.array_equal(y.numpy(), g["ens_uniform_y"])
        y = oracle.ode_block_forward(x, rhs, tabs, grids, "ensemble", ensemble_prob=0.0)
        assert np.array_equal(y.numpy(), g["ens_tails_y"])


def test_premetanode10_whole_model():
    g = golden("premetanode10.npz")
    p = det_premetanode10_params()
    for v in p.values():
        v.requires_grad_(True)
    img = torch.from_numpy(oracle.det_uniform((4, 3, 32, 32), 900, 0.0, 1.0))
    mean = torch.tensor(CIFAR_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CIFAR_STD).view(1, 3, 1, 1)
    x = ((img - mean) / std).requires_grad_(True)
    tab = butcher_tableau("rk2", "u", np.float32(0.5), None)
    taps = {}
    logits = premetanode10_forward(p, x, tab, dict(n_steps=8), taps=taps)
    loss = F.cross_entropy(logits, torch.tensor([3, 1, 4, 1]))
    loss.backward()
    assert np.array_equal(logits.detach().numpy(), g["logits"])
    for k, v in taps.items():
        assert np.array_equal(v.detach().numpy(), g["odeblock_" + k])
    assert np.array_equal(x.grad.numpy(), g["gx"])
    for k in g.files:
        if k.startswith("g_"):
            assert np.array_equal(p[k[2:]].grad.numpy().reshape(-1)[::cases.WG_STRIDE], g[k]), k


def test_numpy_direct_conv_pins_conv_semantics():
    """Independent fp64 direct 3x3 cross-correlation (pad 1) pins what F.conv2d means here."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 3, 5, 6)); w = rng.standard_normal((4, 3, 3, 3))
    ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), None, 1, 1).numpy()
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    out = np.zeros((1, 4, 5, 6))
    for r in range(3):
        for s in range(3):
            out += np.einsum("oc,nchw->nohw", w[:, :, r, s], xp[:, :, r:r + 5, s:s + 6])
    assert max_rel(out, ref) < 1e-14


def test_attacks_and_fgsm_random_train_step_bit_exact():
    """Oracle restatement of FGSM / PGD / FGSM-random + the published training step vs the real reference."""
    from oracle import attacks as oa
    g = golden("attacks.npz")
    p = det_premetanode10_params()
    for v in p.values():
        v.requires_grad_(True)
    img = torch.from_numpy(oracle.det_uniform((8, 3, 32, 32), 910, 0.0, 1.0))
    mean = torch.tensor(CIFAR_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CIFAR_STD).view(1, 3, 1, 1)
    x = (img - mean) / std
    y = torch.tensor([3, 1, 4, 1, 5, 9, 2, 6])
    tab = butcher_tableau("rk2", "u", np.float32(0.5), None)
    model = lambda inp: premetanode10_forward(p, inp, tab, dict(n_steps=8))
    with torch.no_grad():
        assert np.array_equal(model(x).numpy(), g["clean_logits"])
    xf = oa.fgsm(model, x, y, 8 / 255., CIFAR_MEAN, CIFAR_STD)
    assert np.array_equal(xf.numpy(), g["fgsm_x"])
    xp = oa.pgd(model, x, y, 8 / 255., 2 / 255., 7, CIFAR_MEAN, CIFAR_STD, torch.from_numpy(g["pgd_start"]))
    assert np.array_equal(xp.numpy(), g["pgd_x"])
    with torch.no_grad():
        assert np.array_equal(model(xp).numpy(), g["pgd_logits"])
    for v in p.values():
        v.grad = None
    xr = oa.fgsm_random(model, x, y, 10 / 255., 8 / 255., CIFAR_MEAN, CIFAR_STD, torch.from_numpy(g["fgsmr_u01"]))
    assert np.array_equal(xr.numpy(), g["fgsmr_x"])
    loss = F.cross_entropy(model(xr), y)
    loss.backward()
    assert float(loss) == float(g["train_loss"])
    for k in g.files:
        if k.startswith("train_g_"):
            assert np.array_equal(p[k[8:]].grad.numpy().reshape(-1)[::cases.WG_STRIDE], g[k]), k


# ---------------------------------------------------------------- BASELINE config 3: ensembles
def _c3_tabs(svs):
    tabs = []
    for s in svs:
        v = None if s[5] == -1 else np.float32(s[5])
        tabs.append(butcher_tableau(s[0], s[1], np.float32(s[4]), v))
    return tabs, [_grid(s) for s in svs]


@pytest.mark.parametrize("tag,shape,svs,weights", [
    ("c64_rk2x4_uniform", (64, 8, 32, 2), cases.C3_RK2_SOLVERS, None),
    ("c64_rk2x4_weighted", (64, 8, 32, 2), cases.C3_RK2_SOLVERS, cases.C3_WEIGHTS),
    ("c64_rk4x2_uniform", (64, 8, 32, 2), cases.C3_RK4_SOLVERS, None),
    ("c128_rk2x4_uniform", (128, 8, 16, 1), cases.C3_RK2_SOLVERS, None)])
def test_solver_ensemble_block_bit_exact(tag, shape, svs, weights):
    g = golden("ensemble_c3.npz")
    x, w1, w2, r = [torch.from_numpy(a) for a in cases.ode_case_inputs(*shape)]
    x.requires_grad_(True); w1.requires_grad_(True); w2.requires_grad_(True)
    tabs, grids = _c3_tabs(svs)
    cnt = RhsCounter()
    y = oracle.ode_block_forward(x, rhs_preact(w1, w2, counter=cnt), tabs, grids, "ensemble", ensemble_prob=1.0,
                                 ensemble_weights=weights)
    (y * r).sum().backward()
    assert np.array_equal(y.detach().numpy(), g[tag + "_y"])
    assert np.array_equal(x.grad.numpy(), g[tag + "_gx"])
    assert np.array_equal(w1.grad.numpy().reshape(-1)[::cases.WG_STRIDE], g[tag + "_gw1"])
    assert np.array_equal(w2.grad.numpy().reshape(-1)[::cases.WG_STRIDE], g[tag + "_gw2"])
    assert cnt.nfe == int(g[tag + "_nfe"])


def test_model_solver_and_model_ensembling_bit_exact():
    from oracle import attacks as oa
    g = golden("ensemble_c3.npz")
    p = det_premetanode10_params()
    for v in p.values():
        v.requires_grad_(True)
    img = torch.from_numpy(oracle.det_uniform((4, 3, 32, 32), 920, 0.0, 1.0))
    mean = torch.tensor(CIFAR_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CIFAR_STD).view(1, 3, 1, 1)
    labels = torch.tensor([3, 1, 4, 1])
    tabs, grids = _c3_tabs(cases.C3_RK2_SOLVERS)
    x = ((img - mean) / std).requires_grad_(True)
    logits = premetanode10_forward(p, x, tabs, grids)
    F.cross_entropy(logits, labels).backward()
    assert np.array_equal(logits.detach().numpy(), g["model_solver_ens_logits"])
    assert np.array_equal(x.grad.numpy(), g["model_solver_ens_gx"])
    for k in g.files:
        if k.startswith("model_solver_ens_g_"):
            name = k[len("model_solver_ens_g_"):]
            assert np.array_equal(p[name].grad.numpy().reshape(-1)[::cases.WG_STRIDE], g[k]), k
    # model ensembling: the same network under each of the 4 solvers, softmax-averaged (fgsm.py:135-143)
    x = (img - mean) / std
    models = [lambda inp, t=t, gr=gr: premetanode10_forward(p, inp, t, gr) for t, gr in zip(tabs, grids)]
    with torch.no_grad():
        probs = 0
        for m in models:
            probs = probs + torch.softmax(m(x), dim=1)
        assert np.array_equal((probs / len(models)).numpy(), g["model_ens_probs"])
    xa = oa.fgsm_2ensemble(models, x, labels, 8 / 255., CIFAR_MEAN, CIFAR_STD)
    assert np.array_equal(xa.numpy(), g["model_ens_fgsm_x"])


def test_non_ode_layer_gradients_bit_exact():
    """stem pre-activation, strided residual block output and the gradients of both residual blocks."""
    g = golden("premetanode10_resgrads.npz")
    p = det_premetanode10_params()
    for v in p.values():
        v.requires_grad_(True)
    img = torch.from_numpy(oracle.det_uniform((4, 3, 32, 32), 900, 0.0, 1.0))
    mean = torch.tensor(CIFAR_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CIFAR_STD).view(1, 3, 1, 1)
    x = ((img - mean) / std)
    tab = butcher_tableau("rk2", "u", np.float32(0.5), None)
    logits = premetanode10_forward(p, x, tab, dict(n_steps=8))
    F.cross_entropy(logits, torch.tensor([3, 1, 4, 1])).backward()
    assert np.array_equal(F.conv2d(x, p["conv1.weight"], None, 1, 1).detach().numpy(), g["stem_preact"])
    for k in g.files:
        if k.startswith("g_"):
            assert np.array_equal(p[k[2:]].grad.numpy().reshape(-1)[::cases.WG_STRIDE], g[k]), k


@pytest.mark.parametrize("case", cases.SOLVER_GRAD_CASES, ids=[c[0] for c in cases.SOLVER_GRAD_CASES])
def test_solver_parameter_gradients(case):
    """dL/du, dL/dv after unfreeze_params() (order2stage2.py:104-109): the oracle differentiates through the same
    closed-form tableau; fp32 must reproduce the reference's fp32 run, fp64 its fp64 run."""
    from oracle.tableau import butcher_tableau_tensors
    name, C, H, W, B, kind, sv = case
    method, param, n_steps, step_size, u0, v0 = sv
    g = golden("solver_grads.npz")
    for dt, tag, tol in ((torch.float32, "f32", 0.0), (torch.float64, "f64", 1e-9)):
        x, w1, w2, r = [torch.from_numpy(a).to(dt) for a in cases.ode_case_inputs(C, H, W, B)]
        u = torch.tensor((u0,), dtype=dt, requires_grad=True)
        v = torch.tensor((v0,), dtype=dt, requires_grad=True) if v0 != -1 else None
        tab = butcher_tableau_tensors(method, param, u, v, dt)
        rhs = (rhs_preact if kind == "preact" else rhs_postact)(w1, w2)
        y = integrate(tab, rhs, x, torch.tensor([0., 1.]), n_steps=n_steps)[-1]
        (y * r).sum().backward()
        for p, key in ((u, "du"), (v, "dv")):
            if p is None:
                continue
            ref = g["%s_%s_%s" % (name, tag, key)]
            got = p.grad.numpy()
            if tol == 0.0:
                assert np.array_equal(got, ref), (name, tag, key, got, ref)
            else:
                assert abs(float(got[0]) - float(ref[0])) <= tol * max(1.0, abs(float(ref[0]))), (name, tag, key, got, ref)


@pytest.mark.parametrize("case", cases.MULTITIME_CASES, ids=[c[0] for c in cases.MULTITIME_CASES])
def test_integrate_with_interior_output_times_bit_exact(case):
    """rk_parametric.py:104-123: solutions at interior times are linear interpolations between grid points."""
    name, C, H, W, B, sv, times = case
    g = golden("multitime.npz")
    x, w1, w2, r = [torch.from_numpy(a) for a in cases.ode_case_inputs(C, H, W, B)]
    x.requires_grad_(True); w1.requires_grad_(True); w2.requires_grad_(True)
    cnt = RhsCounter()
    tab = butcher_tableau(sv[0], sv[1], np.float32(sv[4]), None if sv[5] == -1 else np.float32(sv[5]))
    ys = integrate(tab, rhs_preact(w1, w2, "gelu", cnt), x, torch.tensor(times), n_steps=sv[2])
    sum(((k + 1.0) * ys[k] * r).sum() for k in range(1, len(times))).backward()
    assert cnt.nfe == int(g[name + "_nfe"])
    assert np.array_equal(ys.detach().numpy()[1:, :, ::3], g[name + "_y"])
    assert np.array_equal(x.grad.numpy(), g[name + "_gx"])
    assert np.array_equal(w1.grad.numpy().reshape(-1)[::cases.WG_STRIDE], g[name + "_gw1"])


@pytest.mark.parametrize("case", cases.MNIST_SOLVER_GRAD_CASES, ids=[c[0] for c in cases.MNIST_SOLVER_GRAD_CASES])
def test_solver_parameter_gradients_mnist(case):
    """Time-dependent right-hand side: dL/du also flows through the nodes c_i (t_i = t_n + c_i dt, order2stage2.py:81-86)."""
    from oracle.tableau import butcher_tableau_tensors
    tag, sv = case
    method, param, n_steps, step_size, u0, v0 = sv
    g = golden("solver_grads_mnist.npz")
    w = golden("mnist_odeblock_weights.npz")
    feat = golden("mnist_odeblock.npz")["feat"]
    for dt, dn, tol in ((torch.float32, "f32", 2e-6), (torch.float64, "f64", 1e-10)):
        po = {k: torch.from_numpy(w[k]).to(dt) for k in w.files}
        x = torch.from_numpy(feat).to(dt).requires_grad_(True)
        u = torch.tensor((u0,), dtype=dt, requires_grad=True)
        v = torch.tensor((v0,), dtype=dt, requires_grad=True) if v0 != -1 else None
        tab = butcher_tableau_tensors(method, param, u, v, dt)
        y = integrate(tab, rhs_mnist(po), x, torch.tensor([0., 1.]), n_steps=n_steps)[-1]
        r = torch.from_numpy(det_normal(tuple(y.shape), 77)).to(dt)
        (y * r).sum().backward()
        for p, key in ((u, "du"), (v, "dv")):
            if p is None:
                continue
            ref, got = float(g["%s_%s_%s" % (tag, dn, key)][0]), float(p.grad[0])
            assert abs(got - ref) <= tol * abs(ref), (tag, dn, key, got, ref)
        if dn == "f32":
            assert max_rel(x.grad.numpy(), g[tag + "_f32_gx"]) <= 1e-6


@pytest.mark.parametrize("case", cases.GN_CASES, ids=[c[0] for c in cases.GN_CASES])
def test_group_norm_ode_block_bit_exact(case):
    """PreBasicBlock2 with the per-sample normalisations 'GN' / 'LN' / 'IN' (cifar10/utils.py:26-36) inside the ODE block."""
    from oracle import rhs_preact_gn
    name, C, H, W, B, norm_key, groups, sv = case
    g = golden("gn_blocks.npz")
    x, w1, w2, r = [torch.from_numpy(a) for a in cases.ode_case_inputs(C, H, W, B)]
    G = {"GN": groups, "LN": 1, "IN": C}[norm_key]
    p = dict(conv1_w=w1, conv2_w=w2)
    for k in (0, 1):
        gw, gb = cases.gn_affine(C, k) if norm_key != "IN" else (np.ones(C, np.float32), np.zeros(C, np.float32))
        p["norm%d_w" % (k + 1)], p["norm%d_b" % (k + 1)] = torch.from_numpy(gw), torch.from_numpy(gb)
    for t in list(p.values()) + [x]:
        t.requires_grad_(True)
    cnt = RhsCounter()
    tab = butcher_tableau(sv[0], sv[1], None if sv[0] == "euler" else np.float32(sv[4]), None if sv[5] == -1 else np.float32(sv[5]))
    y = integrate(tab, rhs_preact_gn(p, G, 1e-5, "gelu", cnt, instance_norm=(norm_key == "IN")), x, torch.tensor([0., 1.]), n_steps=sv[2])[-1]
    (y * r).sum().backward()
    assert cnt.nfe == int(g[name + "_nfe"])
    assert np.array_equal(y.detach().numpy(), g[name + "_y"])
    assert np.array_equal(x.grad.numpy(), g[name + "_gx"])
    assert np.array_equal(p["conv1_w"].grad.numpy().reshape(-1)[::cases.WG_STRIDE], g[name + "_gw1"])
    if norm_key != "IN":
        assert np.array_equal(p["norm2_b"].grad.numpy(), g[name + "_gnorm2_b"])


@pytest.mark.parametrize("case", cases.GN_POST_CASES, ids=[c[0] for c in cases.GN_POST_CASES])
def test_group_norm_postact_ode_block_bit_exact(case):
    """BasicBlock2 (post-activation right-hand side, cifar10/layers.py:108-121) with the per-sample normalisations."""
    from oracle import rhs_postact_gn
    name, C, H, W, B, norm_key, groups, sv = case
    g = golden("gn_post_blocks.npz")
    x, w1, w2, r = [torch.from_numpy(a) for a in cases.ode_case_inputs(C, H, W, B)]
    G = {"GN": groups, "LN": 1, "IN": C}[norm_key]
    p = dict(conv1_w=w1, conv2_w=w2)
    for k in (0, 1):
        gw, gb = cases.gn_affine(C, k) if norm_key != "IN" else (np.ones(C, np.float32), np.zeros(C, np.float32))
        p["norm%d_w" % (k + 1)], p["norm%d_b" % (k + 1)] = torch.from_numpy(gw), torch.from_numpy(gb)
    for t in list(p.values()) + [x]:
        t.requires_grad_(True)
    cnt = RhsCounter()
    tab = butcher_tableau(sv[0], sv[1], None if sv[0] == "euler" else np.float32(sv[4]), None if sv[5] == -1 else np.float32(sv[5]))
    y = integrate(tab, rhs_postact_gn(p, G, 1e-5, "gelu", cnt, instance_norm=(norm_key == "IN")), x, torch.tensor([0., 1.]), n_steps=sv[2])[-1]
    (y * r).sum().backward()
    assert cnt.nfe == int(g[name + "_nfe"])
    assert np.array_equal(y.detach().numpy(), g[name + "_y"])
    assert np.array_equal(x.grad.numpy(), g[name + "_gx"])
    assert np.array_equal(p["conv1_w"].grad.numpy().reshape(-1)[::cases.WG_STRIDE], g[name + "_gw1"])
    if norm_key != "IN":
        assert np.array_equal(p["norm2_b"].grad.numpy(), g[name + "_gnorm2_b"])
