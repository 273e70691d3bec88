"""CPU: host side of the product -- solver objects (tableaus, grids, smoothing API) and the C-ABI
library (loads, exports every symbol the header declares).  No GPU compute here."""
import copy
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import golden, ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_cases as cases  # noqa: E402
import metasolver_b200  # noqa: E402
from metasolver_b200 import _cabi  # noqa: E402
from metasolver_b200.sopa.src.solvers.utils import (create_solver, noise_params,  # noqa: E402
                                                    create_solver_ensemble_by_noising_params)
from metasolver_b200.sopa.src.solvers.rk_parametric_order2stage2 import RKOrder2Stage2  # noqa: E402


@pytest.mark.parametrize("idx", range(len(cases.TABLEAU_CASES)))
def test_solver_tableau_bit_exact_vs_reference(idx):
    g = golden("tableaus.npz")
    m, p, u0, v0 = cases.TABLEAU_CASES[idx]
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        s = create_solver(m, p, 4, -1, u0, v0, dt, "cpu")
        t = s.host_tableau()
        assert np.array_equal(np.array(t["c"]), g["%d_%s_c" % (idx, tag)])
        assert np.array_equal(np.array(t["b"]), g["%d_%s_b" % (idx, tag)])
        assert np.array_equal(np.array(t["w"]), g["%d_%s_w" % (idx, tag)])
        c, w, b = s.build_ButcherTableau(return_tableau=True)
        assert len(c) == len(b) == len(w) == t["stages"]


def test_time_grids_bit_exact_vs_reference():
    g = golden("grids.npz")
    t01 = torch.tensor([0, 1]).float()
    for n in (1, 2, 3, 5, 7, 8, 10, 16):
        s = create_solver("rk2", "u", n, -1, 0.5, -1, torch.float32, "cpu")
        assert np.array_equal(s.host_time_grid(t01).numpy(), g["n%d" % n])
    for ss in (0.3, 0.125, 0.4):
        s = create_solver("rk2", "u", -1, ss, 0.5, -1, torch.float32, "cpu")
        assert np.array_equal(s.host_time_grid(t01).numpy(), g["ss%g" % ss])
    # grid_constructor stays externally assignable (sopa/src/models/odenet_mnist/metrics.py:35)
    s.grid_constructor = s._grid_constructor_from_n_steps(5)
    assert np.array_equal(s.host_time_grid(t01).numpy(), g["n5"])


def test_solver_api_surface():
    with pytest.raises(ValueError):
        RKOrder2Stage2(parameterization="uv", u0=0.5, dtype=torch.float32, n_steps=2)
    with pytest.raises(ValueError):
        create_solver("rk2", "u", 4, 0.25, 0.5, -1, torch.float32, "cpu")     # n_steps and step_size together
    s = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, "cpu")
    assert s.order == 2 and isinstance(s.u, torch.nn.Parameter) and s.u.requires_grad
    s.freeze_params()
    assert not s.u.requires_grad and float(s.b2) == 1.0 and float(s.w21) == 0.5 and s.v is None
    # solver smoothing (examples/cifar10/train_and_attack.py:266-273, 320-323)
    torch.manual_seed(0)
    s.u, s.v = noise_params(s.u0, s.v0, std=0.0125, bernoulli_p=1.0, noise_type="normal")
    s.build_ButcherTableau()
    u = float(s.u)
    assert abs(u - 0.5) < 0.025 and abs(s.host_tableau()["b"][1] - np.float32(1.0) / (np.float32(2) * np.float32(u))) == 0
    s.u, s.v = s.u0, s.v0
    s.build_ButcherTableau()
    assert s.host_tableau()["c"][1] == 0.5
    ens = create_solver_ensemble_by_noising_params(s, 3, dict(std=0.2, noise_type="normal"))
    assert len(ens) == 3 and ens[0] is s and ens[1].host_tableau()["c"][1] != 0.5
    assert copy.deepcopy(s).host_tableau() == s.host_tableau()
    assert create_solver("euler", None, 2, -1, -1, -1, torch.float32, "cpu").order == 1
    assert create_solver("rk4", "uv", 2, -1, 1 / 3., 2 / 3., torch.float32, "cpu").order == 4


def test_cpu_tensors_are_refused_loudly():
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import MetaODEBlock, PreBasicBlock2
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    import torch.nn.functional as F
    from argparse import Namespace
    blk = MetaODEBlock(PreBasicBlock2(64, norm_layer=Identity, act_layer=F.gelu))
    s = create_solver("rk2", "u", 2, -1, 0.5, -1, torch.float32, "cpu")
    s.freeze_params()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk(torch.zeros(1, 64, 4, 32), [s], Namespace(solver_mode="standalone"))
    with pytest.raises(RuntimeError):
        blk.rhs_func(0.0, torch.zeros(1, 64, 4, 32))


def test_cabi_library_exports_header_symbols():
    header = open(os.path.join(ROOT, "include", "metasolver_b200.h")).read()
    declared = set(re.findall(r"\b(msb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    lib = _cabi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.msb_abi_version() == _cabi.ABI_VERSION == 6
    assert lib.msb_shape_supports_tcgen05(64, 32, 32) in (0, 1)
    # descriptor validation is host-only: bad stage count must be refused with a message
    d = _cabi.MsbOdeDesc()
    d.stages = 9
    assert lib.msb_odeblock_tape_bytes(ctypes.byref(d)) == 0
    assert b"stages" in lib.msb_last_error() or b"rhs" in lib.msb_last_error()


def test_binding_constants_match_the_header():
    header = open(os.path.join(ROOT, "include", "metasolver_b200.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(MSB_[A-Z_]+)\s+(\d+)", header)}
    assert defs["MSB_ABI_VERSION"] == _cabi.ABI_VERSION
    assert (defs["MSB_PEER_MAX_RANKS"], defs["MSB_PEER_HANDLE_BYTES"], defs["MSB_PEER_HEADER_BYTES"]) == (
        _cabi.PEER_MAX_RANKS, _cabi.PEER_HANDLE_BYTES, _cabi.PEER_HEADER_BYTES)
    assert defs["MSB_ATTACK_MAX_CHANNELS"] == _cabi.ATTACK_MAX_CHANNELS
    # the peer exchange validates its arguments on the host before any CUDA call
    lib = _cabi.lib()
    bases = (ctypes.c_void_p * 2)(None, None)
    assert lib.msb_peer_allreduce_sgd(bases, 2, 5, 0, 16, 16, 1, None, None, 0.0, 0.0, 0.0, 1.0, 0, 0, None) == -1
    assert b"bad arguments" in lib.msb_last_error()
    assert lib.msb_peer_allreduce_sgd(bases, 2, 0, 0, 16, -1, 1, None, None, 0.0, 0.0, 0.0, 1.0, 0, 0, None) == -1
    assert b"result" in lib.msb_last_error()
    assert lib.msb_peer_allreduce_sgd(bases, 2, 0, 0, 16, 16, 1, None, None, 0.0, 0.0, 0.0, 1.0, 0, 0, None) == -1
    assert b"null" in lib.msb_last_error()


def test_binding_struct_layouts_match_the_library():
    lib = _cabi.lib()
    for which, cls in enumerate((_cabi.MsbOdeDesc, _cabi.MsbTableau, _cabi.MsbMnistParams, _cabi.MsbMnistGrads,
                                 _cabi.MsbDownDesc)):
        assert lib.msb_sizeof(which) == ctypes.sizeof(cls), cls.__name__
    assert lib.msb_sizeof(99) == 0


def test_descriptor_validation_is_host_side_and_loud():
    """Bad descriptors are refused before any CUDA call (size queries return 0 and set the error text)."""
    lib = _cabi.lib()
    d = _cabi.MsbDownDesc()
    d.act, d.engine, d.batch, d.height, d.width, d.in_channels, d.out_channels = 1, 2, 4, 32, 32, 64, 96
    assert lib.msb_downblock_workspace_bytes(ctypes.byref(d)) == 0
    assert b"out_channels" in lib.msb_last_error()
    d.out_channels, d.height = 128, 31
    assert lib.msb_downblock_tape_bytes(ctypes.byref(d)) == 0
    assert b"even" in lib.msb_last_error()
    d.height = 32
    assert lib.msb_downblock_tape_bytes(ctypes.byref(d)) > 0
    o = _cabi.MsbOdeDesc()
    o.rhs_kind, o.act, o.engine, o.batch, o.height, o.width, o.channels, o.n_steps, o.stages = 0, 1, 2, 6, 8, 8, 16, 2, 2
    grid = (ctypes.c_float * 3)(0.0, 0.5, 1.0)
    o.time_grid = ctypes.cast(grid, ctypes.POINTER(ctypes.c_float))
    assert lib.msb_odeblock_tape_bytes(ctypes.byref(o)) > 0
    o.n_solvers = 4                                   # 6 images do not split into 4 solver slices; no tableaus given
    assert lib.msb_odeblock_tape_bytes(ctypes.byref(o)) == 0
    assert b"solver" in lib.msb_last_error()
    o.n_solvers, o.rhs_kind = 0, 2                    # MNIST tape: five tensors per stage evaluation
    mn = lib.msb_odeblock_tape_bytes(ctypes.byref(o))
    o.rhs_kind = 0
    assert mn * 4 == lib.msb_odeblock_tape_bytes(ctypes.byref(o)) * 5
    assert lib.msb_stem_backward_workspace_bytes(64) > 0


def test_package_detrand_equals_the_oracle_generator_and_shards():
    """metasolver_b200.detrand (synthetic evaluation sets of scripts/eval_pgd_sweep.py, product side) generates the same
    tensors as oracle.detrand (test side), any slice of a tensor can be generated on its own, and the deterministic
    premetanode10 weights equal the oracle's recipe."""
    import torch.nn.functional as F
    from metasolver_b200 import detrand
    from metasolver_b200.sopa.src.models.odenet_cifar10.layers import premetanode10
    from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
    from metasolver_b200.sopa.src.models.odenet_cifar10.data import CIFAR_MEAN, CIFAR_STD, augment_batch
    import oracle
    from oracle.models import det_premetanode10_params
    a = detrand.uniform((7, 3, 4, 5), 9100, 0.0, 1.0)
    assert np.array_equal(a, oracle.det_uniform((7, 3, 4, 5), 9100, 0.0, 1.0))
    per = 3 * 4 * 5
    assert np.array_equal(a[2:5], detrand.uniform((3, 3, 4, 5), 9100, 0.0, 1.0, offset=2 * per))
    model = premetanode10((Identity,) * 3, (lambda x: x,) * 3, (F.gelu,) * 3, in_planes=64, is_odenet=True)
    sd, ref = detrand.premetanode10_state_dict(model), det_premetanode10_params()
    assert list(sd) == list(ref)
    for k in sd:
        assert torch.equal(sd[k], ref[k]), k
    assert (CIFAR_MEAN, CIFAR_STD) == (oracle.models.CIFAR_MEAN, oracle.models.CIFAR_STD)
    # batched RandomCrop(32, padding=4) + RandomHorizontalFlip (data.py:41-43): every output is a shifted / mirrored window
    g = torch.Generator().manual_seed(3)
    img = torch.rand(16, 3, 32, 32)
    out = augment_batch(img, generator=g)
    assert out.shape == img.shape
    pad = torch.nn.functional.pad(img, (4, 4, 4, 4))
    for n in range(16):
        wins = [pad[n, :, dy:dy + 32, dx:dx + 32] for dy in range(9) for dx in range(9)]
        assert any(torch.equal(out[n], w) or torch.equal(out[n], w.flip(-1)) for w in wins), n


def test_noise_samplers_match_reference_golden():
    """noise_params / sample_solver_by_noising_params / create_solver_ensemble_by_noising_params make the
    same torch CPU RNG calls as the reference (solvers/utils.py:60-117): same drawn u / v and bit-identical rebuilt
    tableaus under the same seed, for float32 and float64 solvers (tests/golden/make_golden_noise.py)."""
    import contextlib
    import io
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_noise as mg          # case table only; its reference imports are not needed here
    from metasolver_b200.sopa.src.solvers.utils import (sample_solver_by_noising_params,
                                                        create_solver_ensemble_by_noising_params)
    g = golden("noise_samplers.npz")

    def tab(s):
        c, w, b = s.build_ButcherTableau(return_tableau=True)
        flat = [float(x) for x in c] + [float(x) for x in b]
        for row in w:
            flat += [float(x) for x in np.atleast_1d(row.detach().numpy())]
        return np.asarray(flat, dtype=np.float64)
    for name, args, kw, seed in mg.CASES:
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            solver = create_solver(*args, dt, "cpu")
            solver.freeze_params()
            torch.manual_seed(seed)
            with contextlib.redirect_stdout(io.StringIO()):
                for i in range(mg.N_DRAWS):
                    s2 = sample_solver_by_noising_params(solver, **kw)
                    key = "%s_%s" % (name, tag)
                    assert float(s2.u) == g[key + "_u"][i], (key, i)
                    if s2.v is not None:
                        assert float(s2.v) == g[key + "_v"][i], (key, i)
                    assert str(s2.u.dtype) == str(g[key + "_udtype"][0]), key
                    assert np.array_equal(tab(s2), g[key + "_tab"][i]), (key, i)
                ens = create_solver_ensemble_by_noising_params(solver, ensemble_size=4, kwargs_noise=kw)
            assert np.array_equal(np.asarray([float(e.u) for e in ens]), g[key + "_ens_u"]), key
            assert np.array_equal(np.stack([tab(e) for e in ens]), g[key + "_ens_tab"]), key


@pytest.mark.parametrize("mode,kw", [("triangular", {}), ("triangular2", {}), ("exp_range", {"gamma": 0.99}),
                                     ("triangular", {"step_size_down": 7, "cycle_momentum": False})])
def test_cyclic_lr_matches_torch_scheduler(mode, kw):
    """examples/cifar10/train_and_attack.py:500-505: optim.lr_scheduler.CyclicLR(base_lr, max_lr, step_size_up, mode,
    cycle_momentum).  torch's class only accepts torch optimizers; ours drives FusedSGD.param_groups.  Same values."""
    from metasolver_b200.train_ops import CyclicLR

    class Opt:                                   # what FusedSGD exposes to a scheduler
        def __init__(self):
            self.param_groups = [{"lr": 0.1, "momentum": 0.9, "weight_decay": 5e-4}]
    p = torch.nn.Parameter(torch.zeros(1))
    topt = torch.optim.SGD([p], lr=0.1, momentum=0.9)
    tsch = torch.optim.lr_scheduler.CyclicLR(topt, base_lr=1e-4, max_lr=0.05, step_size_up=5, mode=mode, **kw)
    mine_opt = Opt()
    mine = CyclicLR(mine_opt, base_lr=1e-4, max_lr=0.05, step_size_up=5, mode=mode, **kw)
    for it in range(40):
        assert mine.get_last_lr()[0] == tsch.get_last_lr()[0], it
        assert mine_opt.param_groups[0]["lr"] == topt.param_groups[0]["lr"], it
        assert mine_opt.param_groups[0]["momentum"] == topt.param_groups[0]["momentum"], it
        topt.step()
        tsch.step()
        mine.step()
