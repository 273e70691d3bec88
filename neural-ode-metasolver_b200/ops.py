"""torch.autograd binding of the fused ODE-block kernels (C ABI: include/metasolver_b200.h).

`ode_block_integrate` is what `RKParametricSolver.integrate` (sopa/src/solvers/rk_parametric.py)
calls for a recognised right-hand side: one C call enqueues the whole steps x stages loop on the
current CUDA stream; the backward is the matching fused discretize-then-optimize pass.

Tensors cross the boundary as raw device pointers.  Activations are channels-last (NHWC) inside;
the returned tensor keeps the logical NCHW shape with channels_last strides, so callers written
against the reference keep working unchanged.
"""
import contextlib
import ctypes

import torch

from . import _cabi

# input-gradient-only mode is PROCESS-global on purpose: autograd runs Function.backward on its own engine
# thread, where a thread-local set by the caller of autograd.grad() would not be visible.
_input_only_depth = [0]
_default_engine = ["auto"]
_library_fallback = [False]


def set_library_fallback(allow):
    """Layers of the model zoo AROUND the ODE blocks (stem, residual blocks, head) run on this library's kernels for
    the published configurations.  For anything else (BatchNorm residual blocks, the post-activation `BasicBlock`,
    weight-normed convolutions, CPU tensors ...) they raise, like the ODE blocks do -- unless the caller explicitly opts
    in to PyTorch's own (cuDNN / ATen) kernels for those layers with `set_library_fallback(True)`.  The ODE blocks
    themselves never fall back."""
    _library_fallback[0] = bool(allow)


def library_fallback_allowed():
    return _library_fallback[0]


def require_fallback(what, why):
    if not _library_fallback[0]:
        raise NotImplementedError(
            "metasolver_b200: %s is not implemented on the library's kernels (%s).  There is no silent cuDNN / CPU "
            "fallback; call metasolver_b200.set_library_fallback(True) to run this layer with PyTorch's own kernels."
            % (what, why))


def set_default_engine(name):
    """'auto' (tcgen05 where the shape is covered, SIMT otherwise), 'tcgen05' or 'simt'."""
    if name not in _cabi.ENGINES:
        raise ValueError("unknown engine %r" % (name,))
    _default_engine[0] = name


def launch_count():
    """Kernels launched by libmetasolver_b200.so since it was loaded."""
    return int(_cabi.lib().msb_launch_count())


def set_option(name, value):
    """Set a process-wide tuning option of the library (include/metasolver_b200.h: msb_set_option)."""
    _cabi.check(_cabi.lib().msb_set_option(name.encode(), int(value)), "set_option")


def get_option(name):
    v = ctypes.c_int()
    _cabi.check(_cabi.lib().msb_get_option(name.encode(), ctypes.byref(v)), "get_option")
    return v.value


def profile_enable(on=True):
    """Record CUDA events around every convolution-engine launch (clears earlier records)."""
    _cabi.lib().msb_profile_enable(1 if on else 0)


def profile_read(kind):
    """kind 0 = fwd/dgrad convolutions, 1 = wgrad GEMMs -> (total ms, total algorithmic flops, launches)."""
    ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    _cabi.check(_cabi.lib().msb_profile_read(kind, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n)), "profile_read")
    return ms.value, fl.value, n.value


def profile_read_executed(kind):
    """Executed tensor-core flops (algorithmic x hi/lo products formed) of the recorded launches of `kind`."""
    fl = ctypes.c_double()
    _cabi.check(_cabi.lib().msb_profile_read_executed(kind, ctypes.byref(fl)), "profile_read_executed")
    return fl.value


@contextlib.contextmanager
def input_grad_only():
    """Inside this context backward passes skip the weight gradients (FGSM / PGD input-gradient
    mode, MegaAdversarial/src/attacks/pgd.py:44-46 uses autograd.grad w.r.t. the input only)."""
    _input_only_depth[0] += 1
    try:
        yield
    finally:
        _input_only_depth[0] -= 1


class OdeProblem:
    """Host-side description of one integration: tableau, time grid, RHS family."""

    def __init__(self, rhs_kind, act, tableau, time_grid, engine=None):
        self.rhs_kind = rhs_kind
        self.act = act
        # dict(stages, c, b, w) of python floats -- or a list of them (stacked solver axis: slice s of
        # the batch is integrated with tableau s; all must share the number of stages)
        self.tableaus = list(tableau) if isinstance(tableau, (list, tuple)) else [tableau]
        self.tableau = self.tableaus[0]
        if len(self.tableaus) > _cabi.MSB_MAX_SOLVERS:
            raise ValueError("at most %d stacked solvers per launch" % _cabi.MSB_MAX_SOLVERS)
        if any(t["stages"] != self.tableau["stages"] for t in self.tableaus):
            raise ValueError("stacked solvers must have the same number of stages")
        self.time_grid = [float(v) for v in time_grid]
        self.engine = _cabi.ENGINES[engine or _default_engine[0]]
        if len(self.time_grid) < 2:
            raise ValueError("time grid needs at least two points")

    def desc(self, x_shape, save_tape):
        B, C, H, W = x_shape
        d = _cabi.MsbOdeDesc()
        d.rhs_kind, d.act, d.engine = self.rhs_kind, self.act, self.engine
        d.batch, d.height, d.width, d.channels = B, H, W, C
        d.n_steps = len(self.time_grid) - 1
        s = self.tableau["stages"]
        d.stages = s
        for i in range(s):
            d.c[i] = self.tableau["c"][i]
            d.b[i] = self.tableau["b"][i]
            for j in range(s):
                d.w[i * _cabi.MSB_MAX_STAGES + j] = self.tableau["w"][i][j]
        grid = (ctypes.c_float * len(self.time_grid))(*self.time_grid)
        d.time_grid = ctypes.cast(grid, ctypes.POINTER(ctypes.c_float))
        d.save_tape = 1 if save_tape else 0
        d._keepalive = [grid]
        K = len(self.tableaus)
        if K > 1:
            if B % K:
                raise ValueError("batch %d is not divisible into %d solver slices" % (B, K))
            tabs = (_cabi.MsbTableau * K)()
            for q, t in enumerate(self.tableaus):
                for i in range(s):
                    tabs[q].c[i] = t["c"][i]
                    tabs[q].b[i] = t["b"][i]
                    for j in range(s):
                        tabs[q].w[i * _cabi.MSB_MAX_STAGES + j] = t["w"][i][j]
            d.n_solvers = K
            d.solver_tableaus = ctypes.cast(tabs, ctypes.POINTER(_cabi.MsbTableau))
            d._keepalive.append(tabs)
        return d


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _wants_tape(*tensors):
    """Record a tape only when a backward pass can follow.  Decided by the CALLER of Function.apply: inside
    Function.forward grad mode is always off and ctx.needs_input_grad stays True for trainable parameters even under
    torch.no_grad(), so an evaluation forward of a trainable model would otherwise allocate and write the whole tape."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _check_inputs(x, w1, w2):
    if not x.is_cuda:
        raise RuntimeError("metasolver_b200: the ODE-block path runs on CUDA only (got a %s tensor); "
                           "there is no CPU fallback" % x.device)
    if x.dtype != torch.float32 or w1.dtype != torch.float32 or w2.dtype != torch.float32:
        raise RuntimeError("metasolver_b200: float32 tensors required")
    if x.dim() != 4:
        raise RuntimeError("metasolver_b200: expected a (B, C, H, W) state")
    C = x.shape[1]
    for w in (w1, w2):
        if tuple(w.shape) != (C, C, 3, 3):
            raise RuntimeError("metasolver_b200: conv weight must be (%d, %d, 3, 3), got %s" % (C, C, tuple(w.shape)))


class _OdeBlockFn(torch.autograd.Function):
    """coef (optional): a host float64 tensor [b_1..b_4, w_11..w_44, c_1..c_4] built DIFFERENTIABLY from solver.u / solver.v
    (RKParametricSolver._tableau_torch).  Its values are not used -- the kernels take the bit-exact numpy tableau in
    `prob` -- but when it requires grad the backward pass also reduces dL/db_i, dL/dw_ij on the device and returns
    them as coef's gradient, which autograd chains to u and v (unfreeze_params())."""

    @staticmethod
    def forward(ctx, x, w1, w2, prob, coef=None, save=True):
        _check_inputs(x, w1, w2)
        lib = _cabi.lib()
        dev = x.device
        need_grad = bool(save)
        ctx.coef_shape = None if coef is None else tuple(coef.shape)
        with torch.cuda.device(dev):
            xc = x.detach().contiguous(memory_format=torch.channels_last)
            w1c, w2c = w1.detach().contiguous(), w2.detach().contiguous()
            d = prob.desc(tuple(x.shape), need_grad)
            ws_bytes = lib.msb_odeblock_workspace_bytes(ctypes.byref(d))
            if ws_bytes == 0:
                _cabi.check(-1, "odeblock workspace query")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            tape = None
            tape_bytes = 0
            if need_grad:
                tape_bytes = lib.msb_odeblock_tape_bytes(ctypes.byref(d))
                tape = torch.empty(tape_bytes, dtype=torch.uint8, device=dev)
            y = torch.empty_like(xc)          # channels_last strides, logical NCHW
            rc = lib.msb_odeblock_forward(ctypes.byref(d), _ptr(xc), _ptr(w1c), _ptr(w2c), None, _ptr(y),
                                          _ptr(ws), ws_bytes, _ptr(tape), tape_bytes, _stream(dev))
            _cabi.check(rc, "odeblock forward")
        ctx.prob = prob
        ctx.tape = tape
        ctx.tape_bytes = tape_bytes
        ctx.shape = tuple(x.shape)
        ctx.save_for_backward(w1c, w2c)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _cabi.lib()
        w1c, w2c = ctx.saved_tensors
        if ctx.tape is None:
            raise RuntimeError("metasolver_b200: backward called but no tape was recorded")
        dev = gy.device
        need_w = (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) and not (_input_only_depth[0] > 0)
        need_coef = ctx.coef_shape is not None and ctx.needs_input_grad[4]
        gcoef = None
        with torch.cuda.device(dev):
            gyc = gy.contiguous(memory_format=torch.channels_last)
            d = ctx.prob.desc(ctx.shape, True)
            gx = torch.empty_like(gyc)
            gw1 = torch.empty_like(w1c) if need_w else None
            gw2 = torch.empty_like(w2c) if need_w else None
            if need_coef:
                ws_bytes = lib.msb_odeblock_bwd_workspace_bytes_tableau(ctypes.byref(d))
                if ws_bytes == 0:
                    _cabi.check(-1, "odeblock tableau-gradient workspace query")
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                gtab = torch.zeros(len(ctx.prob.tableaus) * _cabi.TABLEAU_GRAD_DOUBLES, dtype=torch.float64, device=dev)
                rc = lib.msb_odeblock_backward_tableau(ctypes.byref(d), _ptr(gyc), _ptr(w1c), _ptr(w2c), _ptr(ctx.tape),
                                                       ctx.tape_bytes, _ptr(gx), _ptr(gw1), _ptr(gw2), _ptr(gtab), _ptr(ws),
                                                       ws_bytes, _stream(dev))
                _cabi.check(rc, "odeblock backward (tableau gradients)")
                gcoef = gtab.cpu().reshape(ctx.coef_shape)       # 20 doubles to the host (the solver scalars live there)
            else:
                ws_bytes = lib.msb_odeblock_bwd_workspace_bytes(ctypes.byref(d))
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                rc = lib.msb_odeblock_backward(ctypes.byref(d), _ptr(gyc), _ptr(w1c), _ptr(w2c), _ptr(ctx.tape),
                                               ctx.tape_bytes, _ptr(gx), _ptr(gw1), _ptr(gw2), _ptr(ws), ws_bytes,
                                               _stream(dev))
                _cabi.check(rc, "odeblock backward")
        ctx.tape = None       # the tape is large; release it as soon as it has been consumed
        return gx, gw1, gw2, None, gcoef, None


def _check_coef(prob, coef, stacked_ok=False):
    K = len(prob.tableaus)
    if K != 1 and not stacked_ok:
        raise NotImplementedError("metasolver_b200: gradients w.r.t. solver parameters on a stacked solver axis are "
                                  "implemented for the normalisation-free CIFAR right-hand sides only")
    if coef.dtype != torch.float64 or coef.is_cuda or coef.numel() != K * _cabi.TABLEAU_GRAD_DOUBLES:
        raise ValueError("tableau_coef must be a host float64 tensor of %d x %d elements" % (K, _cabi.TABLEAU_GRAD_DOUBLES))


def ode_block_integrate(x, w1, w2, tableau, time_grid, rhs_kind=_cabi.RHS_PREACT_NF, act=_cabi.ACT_GELU_ERF,
                        engine=None, tableau_coef=None):
    """y(t_end) of dy/dt = f(y) with f = conv2(act(conv1(act(y)))) integrated on `time_grid`
    by the explicit RK method `tableau`; differentiable w.r.t. x, w1, w2.

    `tableau` may be a list of K tableaus (same stage count): the batch is then K equal slices along
    dim 0 and slice s is integrated by solver s, all inside the same kernel launches (stacked solver axis).
    `tableau_coef` (host float64, K x 20, built differentiably from the solvers' u / v): when it requires grad the
    backward pass also returns dL/d(b, w, c) per solver, reduced over that solver's slice."""
    prob = OdeProblem(rhs_kind, act, tableau, time_grid, engine)
    if tableau_coef is not None:
        _check_coef(prob, tableau_coef, stacked_ok=True)
    return _OdeBlockFn.apply(x, w1, w2, prob, tableau_coef, _wants_tape(x, w1, w2, tableau_coef))


def ode_block_integrate_stacked(x, w1, w2, tableaus, time_grid, rhs_kind=_cabi.RHS_PREACT_NF,
                                act=_cabi.ACT_GELU_ERF, engine=None, tableau_coef=None):
    """Solver ensembling on a stacked solver axis: integrate the SAME state x (B,C,H,W) with each of the K
    solvers in one batched set of launches -> (K, B, C, H, W).  Replaces the reference's sequential loop over
    solvers (cifar10/layers.py:198-203); each slice is bit-identical to the one-solver call."""
    K = len(tableaus)
    xs = x.unsqueeze(0).expand(K, *x.shape).reshape(K * x.shape[0], *x.shape[1:])
    y = ode_block_integrate(xs, w1, w2, list(tableaus), time_grid, rhs_kind=rhs_kind, act=act, engine=engine,
                            tableau_coef=tableau_coef)
    return y.view(K, *x.shape)


_MNIST_KEYS = ("norm1_w", "norm1_b", "norm2_w", "norm2_b", "norm3_w", "norm3_b", "conv1_w", "conv1_b", "conv2_w", "conv2_b")
_GN_KEYS = ("norm1_w", "norm1_b", "norm2_w", "norm2_b", "conv1_w", "conv2_w")      # MSB_RHS_PREACT_GN: no biases, two norms


def _gn_params_struct(keep, groups, eps, cls=None):
    """MsbMnistParams / MsbMnistGrads filled from the tensors present in `keep` (absent ones stay NULL)."""
    mp = (cls or _cabi.MsbMnistParams)()
    for i in range(3):
        for kind, arr in (("w", mp.norm_w), ("b", mp.norm_b)):
            t = keep.get("norm%d_%s" % (i + 1, kind))
            arr[i] = t.data_ptr() if t is not None else None
    for i in range(2):
        for kind, arr in (("w", mp.conv_w), ("b", mp.conv_b)):
            t = keep.get("conv%d_%s" % (i + 1, kind))
            arr[i] = t.data_ptr() if t is not None else None
    if cls is None:
        mp.groups, mp.eps = groups, eps
    return mp


class _GnOdeBlockFn(torch.autograd.Function):
    """ODE block with a GroupNorm right-hand side (MNIST: MSB_RHS_MNIST_GN_T; CIFAR 'GN'/'LN'/'IN': MSB_RHS_PREACT_GN);
    backward = fused discretize-then-optimize pass: gradients w.r.t. x and every RHS parameter (and, for the MNIST
    family, the Butcher coefficients through `coef`, see _OdeBlockFn)."""

    @staticmethod
    def forward(ctx, x, prob, groups, eps, keys, coef, save, *params):
        lib = _cabi.lib()
        dev = x.device
        need_grad = bool(save)                                                    # x, coef or any parameter (see _wants_tape)
        ctx.coef_shape = None if coef is None else tuple(coef.shape)
        with torch.cuda.device(dev):
            xc = x.detach().contiguous(memory_format=torch.channels_last)
            keep = {k: v.detach().contiguous() for k, v in zip(keys, params)}
            mp = _gn_params_struct(keep, groups, eps)
            d = prob.desc(tuple(x.shape), need_grad)
            ws_bytes = lib.msb_odeblock_workspace_bytes(ctypes.byref(d))
            if ws_bytes == 0:
                _cabi.check(-1, "odeblock workspace query")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            tape, tape_bytes = None, 0
            if need_grad:
                tape_bytes = lib.msb_odeblock_tape_bytes(ctypes.byref(d))
                tape = torch.empty(tape_bytes, dtype=torch.uint8, device=dev)
            y = torch.empty_like(xc)
            rc = lib.msb_odeblock_forward(ctypes.byref(d), _ptr(xc), None, None, ctypes.byref(mp), _ptr(y), _ptr(ws),
                                          ws_bytes, _ptr(tape), tape_bytes, _stream(dev))
            _cabi.check(rc, "odeblock forward (GroupNorm right-hand side)")
        ctx.prob, ctx.groups, ctx.eps, ctx.tape, ctx.tape_bytes, ctx.shape = prob, groups, eps, tape, tape_bytes, tuple(x.shape)
        ctx.keys = tuple(keys)
        ctx.save_for_backward(*[keep[k] for k in keys])
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _cabi.lib()
        if ctx.tape is None:
            raise RuntimeError("metasolver_b200: backward called but no tape was recorded")
        keys = ctx.keys
        keep = dict(zip(keys, ctx.saved_tensors))
        dev = gy.device
        need_w = any(ctx.needs_input_grad[7:]) and not (_input_only_depth[0] > 0)
        need_coef = ctx.coef_shape is not None and ctx.needs_input_grad[5]
        gcoef = None
        with torch.cuda.device(dev):
            gyc = gy.contiguous(memory_format=torch.channels_last)
            mp = _gn_params_struct(keep, ctx.groups, ctx.eps)
            d = ctx.prob.desc(ctx.shape, True)
            ws_bytes = (lib.msb_odeblock_bwd_workspace_bytes_tableau if need_coef else lib.msb_odeblock_bwd_workspace_bytes)(
                ctypes.byref(d))
            if ws_bytes == 0:
                _cabi.check(-1, "odeblock backward workspace query (GroupNorm right-hand side)")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            gx = torch.empty_like(gyc)
            grads, gstruct = None, None
            if need_w:
                grads = {k: torch.zeros_like(v) for k, v in keep.items()}
                gstruct = _gn_params_struct(grads, None, None, cls=_cabi.MsbMnistGrads)
            if need_coef:
                gtab = torch.zeros(len(ctx.prob.tableaus) * _cabi.TABLEAU_GRAD_DOUBLES, dtype=torch.float64, device=dev)
                rc = lib.msb_odeblock_backward_mnist_tableau(ctypes.byref(d), _ptr(gyc), ctypes.byref(mp), _ptr(ctx.tape),
                                                             ctx.tape_bytes, _ptr(gx), ctypes.byref(gstruct) if need_w else None,
                                                             _ptr(gtab), _ptr(ws), ws_bytes, _stream(dev))
                _cabi.check(rc, "odeblock backward (GroupNorm right-hand side, tableau gradients)")
                gcoef = gtab.cpu().reshape(ctx.coef_shape)
            else:
                rc = lib.msb_odeblock_backward_mnist(ctypes.byref(d), _ptr(gyc), ctypes.byref(mp), _ptr(ctx.tape), ctx.tape_bytes,
                                                     _ptr(gx), ctypes.byref(gstruct) if need_w else None, _ptr(ws), ws_bytes,
                                                     _stream(dev))
                _cabi.check(rc, "odeblock backward (GroupNorm right-hand side)")
        ctx.tape = None
        return (gx, None, None, None, None, gcoef, None) + (tuple(grads[k] for k in keys) if need_w else (None,) * len(keys))


def ode_block_integrate_mnist(x, params, tableau, time_grid, groups, eps=1e-5, tableau_coef=None):
    """MNIST right-hand side (GroupNorm / ReLU / time-concatenated convs, mnist/layers.py:158-171).
    `params`: dict(norm{1,2,3}_{w,b}, conv{1,2}_{w,b}) of fp32 CUDA tensors.  Differentiable w.r.t. x and all params."""
    if not x.is_cuda:
        raise RuntimeError("metasolver_b200: the ODE-block path runs on CUDA only (got a %s tensor); "
                           "there is no CPU fallback" % x.device)
    prob = OdeProblem(_cabi.RHS_MNIST_GN_T, _cabi.ACT_RELU, tableau, time_grid, "simt")
    if tableau_coef is not None:
        _check_coef(prob, tableau_coef)
    ps = [params[k] for k in _MNIST_KEYS]
    return _GnOdeBlockFn.apply(x, prob, groups, eps, _MNIST_KEYS, tableau_coef, _wants_tape(x, tableau_coef, *ps), *ps)


def ode_block_integrate_gn(x, params, tableau, time_grid, groups, eps=1e-5, act=_cabi.ACT_GELU_ERF, engine=None,
                           rhs_kind=_cabi.RHS_PREACT_GN):
    """CIFAR right-hand sides with GroupNorm: pre-activation conv2(act(GN2(conv1(act(GN1(x)))))) (cifar10/layers.py:148-161)
    or, with rhs_kind = RHS_POSTACT_GN, post-activation act(GN2(conv2(act(GN1(conv1(x)))))) (:108-121), with the 'GN' /
    'LN' / 'IN' normalisations of cifar10/utils.py:26-36.  `params`: dict(norm{1,2}_{w,b}, conv{1,2}_w)."""
    if not x.is_cuda:
        raise RuntimeError("metasolver_b200: the ODE-block path runs on CUDA only (got a %s tensor); "
                           "there is no CPU fallback" % x.device)
    C = x.shape[1]
    for k in ("conv1_w", "conv2_w"):
        if tuple(params[k].shape) != (C, C, 3, 3):
            raise RuntimeError("metasolver_b200: conv weight must be (%d, %d, 3, 3), got %s" % (C, C, tuple(params[k].shape)))
    if C % groups:
        raise RuntimeError("metasolver_b200: %d channels are not divisible into %d groups" % (C, groups))
    if rhs_kind not in (_cabi.RHS_PREACT_GN, _cabi.RHS_POSTACT_GN):
        raise ValueError("ode_block_integrate_gn: rhs_kind must be RHS_PREACT_GN or RHS_POSTACT_GN")
    prob = OdeProblem(rhs_kind, act, tableau, time_grid, engine)
    ps = [params[k] for k in _GN_KEYS]
    return _GnOdeBlockFn.apply(x, prob, groups, eps, _GN_KEYS, None, _wants_tape(x, *ps), *ps)


# --------------------------------------------------------------------------- non-ODE layers
def _input_only():
    return (_input_only_depth[0] > 0)


class _StemFn(torch.autograd.Function):
    """y = act(conv3x3(x, w)), 3 -> C channels (MetaNODE stem, cifar10/layers.py:411-413)."""

    @staticmethod
    def forward(ctx, x, w, act, save=True):
        lib = _cabi.lib()
        dev = x.device
        need_grad = bool(save)
        B, _, H, W = x.shape
        C = w.shape[0]
        with torch.cuda.device(dev):
            xc = x.detach().contiguous(memory_format=torch.channels_last)
            wc = w.detach().contiguous()
            y = torch.empty((B, C, H, W), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
            dact = torch.empty_like(y) if need_grad else None
            _cabi.check(lib.msb_stem_forward(_ptr(xc), _ptr(wc), act, _ptr(y), _ptr(dact), B, H, W, C, _stream(dev)),
                        "stem forward")
        ctx.save_for_backward(xc, wc, dact)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _cabi.lib()
        xc, wc, dact = ctx.saved_tensors
        dev = gy.device
        B, _, H, W = xc.shape
        C = wc.shape[0]
        need_w = ctx.needs_input_grad[1] and not _input_only()
        need_x = ctx.needs_input_grad[0]
        with torch.cuda.device(dev):
            gyc = gy.contiguous(memory_format=torch.channels_last)
            ws_bytes = lib.msb_stem_backward_workspace_bytes(C)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            gw = torch.empty_like(wc) if need_w else None
            gx = torch.empty_like(xc) if need_x else None
            _cabi.check(lib.msb_stem_backward(_ptr(gyc), _ptr(dact), _ptr(xc), _ptr(wc), _ptr(gw), _ptr(gx), B, H, W, C,
                                              _ptr(ws), ws_bytes, _stream(dev)), "stem backward")
        return gx, gw, None, None


def stem_conv_act(x, w, act=_cabi.ACT_GELU_ERF):
    """act(conv2d(x, w, stride 1, padding 1)) for a 3-channel input; differentiable w.r.t. x and w."""
    if not x.is_cuda:
        raise RuntimeError("metasolver_b200: CUDA tensors required (got %s)" % x.device)
    if x.dtype != torch.float32 or w.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3 \
            or tuple(w.shape[1:]) != (3, 3, 3):
        raise RuntimeError("metasolver_b200: stem expects fp32 (B,3,H,W) input and (C,3,3,3) weight")
    return _StemFn.apply(x, w, act, _wants_tape(x, w))


class _DownBlockFn(torch.autograd.Function):
    """Strided pre-activation residual block (PreBasicBlock, stride 2, 1x1 shortcut; cifar10/layers.py:54-81)."""

    @staticmethod
    def _desc(x_shape, co, act, engine, save):
        d = _cabi.MsbDownDesc()
        d.act, d.engine = act, engine
        d.batch, d.in_channels, d.height, d.width = x_shape
        d.out_channels = co
        d.save_tape = 1 if save else 0
        return d

    @staticmethod
    def forward(ctx, x, w1, w2, wsc, act, engine, save=True):
        lib = _cabi.lib()
        dev = x.device
        need_grad = bool(save)
        B, Ci, H, W = x.shape
        Co = w1.shape[0]
        with torch.cuda.device(dev):
            xc = x.detach().contiguous(memory_format=torch.channels_last)
            ws_ = [w.detach().contiguous() for w in (w1, w2, wsc)]
            d = _DownBlockFn._desc(tuple(x.shape), Co, act, engine, need_grad)
            ws_bytes = lib.msb_downblock_workspace_bytes(ctypes.byref(d))
            if ws_bytes == 0:
                _cabi.check(-1, "downblock workspace query")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            tape, tape_bytes = None, 0
            if need_grad:
                tape_bytes = lib.msb_downblock_tape_bytes(ctypes.byref(d))
                tape = torch.empty(tape_bytes, dtype=torch.uint8, device=dev)
            y = torch.empty((B, Co, H // 2, W // 2), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
            _cabi.check(lib.msb_downblock_forward(ctypes.byref(d), _ptr(xc), _ptr(ws_[0]), _ptr(ws_[1]), _ptr(ws_[2]),
                                                  _ptr(y), _ptr(ws), ws_bytes, _ptr(tape), tape_bytes, _stream(dev)),
                        "downblock forward")
        ctx.tape, ctx.tape_bytes, ctx.shape, ctx.co, ctx.act, ctx.engine = tape, tape_bytes, tuple(x.shape), Co, act, engine
        ctx.save_for_backward(*ws_)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _cabi.lib()
        w1, w2, wsc = ctx.saved_tensors
        if ctx.tape is None:
            raise RuntimeError("metasolver_b200: backward called but no tape was recorded")
        dev = gy.device
        need_w = any(ctx.needs_input_grad[1:4]) and not _input_only()
        with torch.cuda.device(dev):
            gyc = gy.contiguous(memory_format=torch.channels_last)
            d = _DownBlockFn._desc(ctx.shape, ctx.co, ctx.act, ctx.engine, True)
            ws_bytes = lib.msb_downblock_bwd_workspace_bytes(ctypes.byref(d))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            B, Ci, H, W = ctx.shape
            gx = torch.empty((B, Ci, H, W), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
            gws = [torch.empty_like(w) if need_w else None for w in (w1, w2, wsc)]
            _cabi.check(lib.msb_downblock_backward(ctypes.byref(d), _ptr(gyc), _ptr(w1), _ptr(w2), _ptr(wsc), _ptr(ctx.tape),
                                                   ctx.tape_bytes, _ptr(gx), _ptr(gws[0]), _ptr(gws[1]), _ptr(gws[2]),
                                                   _ptr(ws), ws_bytes, _stream(dev)), "downblock backward")
        ctx.tape = None
        return gx, gws[0], gws[1], gws[2], None, None, None


def resblock_down(x, w1, w2, wsc, act=_cabi.ACT_GELU_ERF, engine=None):
    """conv2(act(conv1_stride2(act(x)))) + conv1x1_stride2(x); w1 (2C,C,3,3), w2 (2C,2C,3,3), wsc (2C,C,1,1)."""
    if not x.is_cuda:
        raise RuntimeError("metasolver_b200: CUDA tensors required (got %s)" % x.device)
    C = x.shape[1]
    if tuple(w1.shape) != (2 * C, C, 3, 3) or tuple(w2.shape) != (2 * C, 2 * C, 3, 3) or tuple(wsc.shape) != (2 * C, C, 1, 1):
        raise RuntimeError("metasolver_b200: strided block weight shapes do not match a %d -> %d block" % (C, 2 * C))
    return _DownBlockFn.apply(x, w1, w2, wsc, act, _cabi.ENGINES[engine or _default_engine[0]], _wants_tape(x, w1, w2, wsc))


class _PoolFcFn(torch.autograd.Function):
    """logits = Linear(mean over pixels(x)): the head of MetaNODE (cifar10/layers.py:390-392, 425) as one kernel."""

    @staticmethod
    def forward(ctx, x, w, b):
        lib = _cabi.lib()
        dev = x.device
        B, C, H, W = x.shape
        K = w.shape[0]
        with torch.cuda.device(dev):
            xc = x.detach().contiguous(memory_format=torch.channels_last)
            wc = w.detach().contiguous()
            bc = b.detach().contiguous() if b is not None else None
            pooled = torch.empty((B, C), dtype=torch.float32, device=dev)
            logits = torch.empty((B, K), dtype=torch.float32, device=dev)
            _cabi.check(lib.msb_pool_fc_forward(_ptr(xc), _ptr(wc), _ptr(bc), _ptr(pooled), _ptr(logits), B, H * W, C, K,
                                                _stream(dev)), "pool_fc forward")
        ctx.save_for_backward(wc, pooled)
        ctx.geom = (B, C, H, W, K, b is not None)
        return logits

    @staticmethod
    def backward(ctx, g):
        lib = _cabi.lib()
        wc, pooled = ctx.saved_tensors
        B, C, H, W, K, has_b = ctx.geom
        dev = g.device
        need_x = ctx.needs_input_grad[0]
        need_w = (ctx.needs_input_grad[1] or (has_b and ctx.needs_input_grad[2])) and not _input_only()
        with torch.cuda.device(dev):
            gc = g.contiguous()
            dx = torch.empty((B, C, H, W), dtype=torch.float32, device=dev, memory_format=torch.channels_last) if need_x else None
            dw = torch.empty_like(wc) if need_w else None
            db = torch.empty(K, dtype=torch.float32, device=dev) if (need_w and has_b) else None
            _cabi.check(lib.msb_pool_fc_backward(_ptr(gc), _ptr(wc), _ptr(pooled), _ptr(dx), _ptr(dw), _ptr(db), B, H * W, C, K,
                                                 _stream(dev)), "pool_fc backward")
        return dx, dw, db


def pool_fc(x, w, b=None):
    """AdaptiveAvgPool2d((1,1)) + Flatten + Linear on a (B,C,H,W) CUDA fp32 map; differentiable w.r.t. x, w, b."""
    if not x.is_cuda:
        raise RuntimeError("metasolver_b200: CUDA tensors required (got %s)" % x.device)
    if x.dtype != torch.float32 or x.dim() != 4 or w.dim() != 2 or w.shape[1] != x.shape[1] or x.shape[1] % 4:
        raise RuntimeError("metasolver_b200: pool_fc expects an fp32 (B,C,H,W) map with C %% 4 == 0 and a (K,C) weight")
    return _PoolFcFn.apply(x, w, b)


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        lib = _cabi.lib()
        dev = logits.device
        B, K = logits.shape
        with torch.cuda.device(dev):
            z = logits.detach().contiguous()
            y = labels.contiguous()
            loss = torch.empty((), dtype=torch.float32, device=dev)
            lse = torch.empty(B, dtype=torch.float32, device=dev)
            _cabi.check(lib.msb_cross_entropy_forward(_ptr(z), _ptr(y), _ptr(loss), _ptr(lse), B, K, _stream(dev)),
                        "cross_entropy forward")
        ctx.save_for_backward(z, y, lse)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _cabi.lib()
        z, y, lse = ctx.saved_tensors
        dev = z.device
        B, K = z.shape
        with torch.cuda.device(dev):
            gc = g.reshape(1).to(torch.float32).contiguous()
            dz = torch.empty_like(z)
            _cabi.check(lib.msb_cross_entropy_backward(_ptr(z), _ptr(y), _ptr(lse), _ptr(gc), _ptr(dz), B, K, _stream(dev)),
                        "cross_entropy backward")
        return dz, None


def cross_entropy(logits, labels):
    """F.cross_entropy(logits, labels) (mean reduction) for CUDA fp32 (B,K) logits and int64 (B,) labels, as one kernel
    each way (the loss of examples/cifar10/train_and_attack.py:303-311 and of the attacks, fgsm.py:33 / pgd.py:43)."""
    if not logits.is_cuda or logits.dtype != torch.float32 or logits.dim() != 2:
        raise RuntimeError("metasolver_b200: cross_entropy expects CUDA fp32 (B,K) logits")
    if labels.dtype != torch.int64 or labels.shape != (logits.shape[0],) or labels.device != logits.device:
        raise RuntimeError("metasolver_b200: cross_entropy expects int64 (B,) labels on the logits' device")
    return _CrossEntropyFn.apply(logits, labels)


# --------------------------------------------------------------------------- single-kernel entry points
def act_split(x_cl, act=_cabi.ACT_NONE, want_dact=False):
    """x_cl: (B,C,H,W) channels_last fp32 -> (split uint16-view tensor [B,H,2,W,C], dact or None)."""
    lib = _cabi.lib()
    B, C, H, W = x_cl.shape
    assert x_cl.is_contiguous(memory_format=torch.channels_last)
    split = torch.empty((B, H, 2, W, C), dtype=torch.bfloat16, device=x_cl.device)
    dact = torch.empty_like(x_cl) if want_dact else None
    with torch.cuda.device(x_cl.device):
        _cabi.check(lib.msb_act_split(_ptr(x_cl), act, _ptr(split), _ptr(dact), B, H, W, C, _stream(x_cl.device)),
                    "act_split")
    return split, dact


def conv3x3(split, w, transpose=False, engine="auto"):
    """split: [B,H,2,W,C] bf16 hi/lo operand, w: (C,C,3,3) fp32 -> (B,C,H,W) channels_last fp32."""
    lib = _cabi.lib()
    B, H, _, W, C = split.shape
    dev = split.device
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
    ws_bytes = lib.msb_conv3x3_workspace_bytes(C)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(lib.msb_conv3x3(_ptr(split), _ptr(w.contiguous()), _ptr(out), 1 if transpose else 0,
                                    _cabi.ENGINES[engine], B, H, W, C, _ptr(ws), ws_bytes, _stream(dev)), "conv3x3")
    return out


def wgrad3x3(split_gout, split_in, engine="auto"):
    lib = _cabi.lib()
    B, H, _, W, C = split_in.shape
    dev = split_in.device
    gw = torch.empty((C, C, 3, 3), dtype=torch.float32, device=dev)
    ws_bytes = lib.msb_wgrad3x3_workspace_bytes(C, _cabi.ENGINES[engine])
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(lib.msb_wgrad3x3(_ptr(split_gout), _ptr(split_in), _ptr(gw), _cabi.ENGINES[engine], B, H, W, C,
                                     _ptr(ws), ws_bytes, _stream(dev)), "wgrad3x3")
    return gw
