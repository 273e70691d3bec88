"""ctypes binding of libmetasolver_b200.so (C ABI: include/metasolver_b200.h).

The library is built in-tree (build.py).  Loading fails loudly: the product has no other path.
"""
import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# MSB_LIB_PATH: a differently built copy of the library (the -DMSB_CONV_DEBUG instrumented build of scripts/)
LIB_PATH = os.environ.get("MSB_LIB_PATH") or os.path.join(HERE, "libmetasolver_b200.so")

MSB_MAX_STAGES = 4
TABLEAU_GRAD_DOUBLES = MSB_MAX_STAGES + MSB_MAX_STAGES * MSB_MAX_STAGES + MSB_MAX_STAGES   # [b | w | c]
ABI_VERSION = 6
RHS_PREACT_NF, RHS_POSTACT_NF, RHS_MNIST_GN_T, RHS_PREACT_GN, RHS_POSTACT_GN = 0, 1, 2, 3, 4
ACT_NONE, ACT_GELU_ERF, ACT_RELU = 0, 1, 2
(ATTACK_UNNORMALIZE, ATTACK_NORMALIZE, ATTACK_FGSM_STEP, ATTACK_PGD_STEP, ATTACK_FGSMR_INIT,
 ATTACK_FGSMR_STEP) = range(6)
ATTACK_MAX_CHANNELS = 4
PEER_MAX_RANKS, PEER_HANDLE_BYTES, PEER_HEADER_BYTES = 16, 64, 1024
ENGINE_AUTO, ENGINE_TCGEN05, ENGINE_SIMT = 0, 1, 2
ENGINES = {"auto": ENGINE_AUTO, "tcgen05": ENGINE_TCGEN05, "simt": ENGINE_SIMT}

EXPORTS = [
    "msb_abi_version", "msb_sizeof", "msb_last_error", "msb_device_supports_tcgen05", "msb_shape_supports_tcgen05",
    "msb_odeblock_workspace_bytes", "msb_odeblock_tape_bytes", "msb_odeblock_bwd_workspace_bytes",
    "msb_odeblock_forward", "msb_odeblock_backward", "msb_odeblock_backward_mnist", "msb_act_split", "msb_conv3x3",
    "msb_stem_forward", "msb_stem_backward_workspace_bytes", "msb_stem_backward",
    "msb_downblock_workspace_bytes", "msb_downblock_tape_bytes", "msb_downblock_bwd_workspace_bytes",
    "msb_downblock_forward", "msb_downblock_backward",
    "msb_conv3x3_workspace_bytes", "msb_wgrad3x3", "msb_wgrad3x3_workspace_bytes", "msb_launch_count", "msb_profile_enable", "msb_profile_read", "msb_profile_read_executed",
    "msb_set_option", "msb_get_option", "msb_attack_step", "msb_sgd_step",
    "msb_odeblock_bwd_workspace_bytes_tableau", "msb_odeblock_backward_tableau", "msb_odeblock_backward_mnist_tableau",
    "msb_pool_fc_forward", "msb_pool_fc_backward", "msb_cross_entropy_forward", "msb_cross_entropy_backward",
    "msb_augment_batch",
    "msb_peer_alloc", "msb_peer_open", "msb_peer_close", "msb_peer_free", "msb_peer_status", "msb_peer_allreduce_sgd",
]


MSB_MAX_SOLVERS = 8


class MsbTableau(ctypes.Structure):
    _fields_ = [("c", ctypes.c_float * MSB_MAX_STAGES), ("b", ctypes.c_float * MSB_MAX_STAGES),
                ("w", ctypes.c_float * (MSB_MAX_STAGES * MSB_MAX_STAGES))]


class MsbOdeDesc(ctypes.Structure):
    _fields_ = [
        ("rhs_kind", ctypes.c_int32), ("act", ctypes.c_int32), ("engine", ctypes.c_int32),
        ("batch", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("channels", ctypes.c_int32), ("n_steps", ctypes.c_int32), ("stages", ctypes.c_int32),
        ("c", ctypes.c_float * MSB_MAX_STAGES), ("b", ctypes.c_float * MSB_MAX_STAGES),
        ("w", ctypes.c_float * (MSB_MAX_STAGES * MSB_MAX_STAGES)),
        ("time_grid", ctypes.POINTER(ctypes.c_float)),
        ("save_tape", ctypes.c_int32), ("n_solvers", ctypes.c_int32),
        ("solver_tableaus", ctypes.POINTER(MsbTableau)),
    ]


class MsbDownDesc(ctypes.Structure):
    _fields_ = [("act", ctypes.c_int32), ("engine", ctypes.c_int32), ("batch", ctypes.c_int32),
                ("height", ctypes.c_int32), ("width", ctypes.c_int32), ("in_channels", ctypes.c_int32),
                ("out_channels", ctypes.c_int32), ("save_tape", ctypes.c_int32)]


class MsbMnistParams(ctypes.Structure):
    _fields_ = [
        ("norm_w", ctypes.c_void_p * 3), ("norm_b", ctypes.c_void_p * 3),
        ("conv_w", ctypes.c_void_p * 2), ("conv_b", ctypes.c_void_p * 2),
        ("groups", ctypes.c_int32), ("eps", ctypes.c_float),
    ]


class MsbMnistGrads(ctypes.Structure):
    _fields_ = [
        ("norm_w", ctypes.c_void_p * 3), ("norm_b", ctypes.c_void_p * 3),
        ("conv_w", ctypes.c_void_p * 2), ("conv_b", ctypes.c_void_p * 2),
    ]


_lib = None
_lock = threading.Lock()


def _declare(lib):
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    dp = ctypes.POINTER(MsbOdeDesc)
    lib.msb_abi_version.restype = i32
    lib.msb_sizeof.argtypes = [i32]
    lib.msb_sizeof.restype = sz
    lib.msb_last_error.restype = ctypes.c_char_p
    lib.msb_launch_count.restype = ctypes.c_uint64
    lib.msb_device_supports_tcgen05.argtypes = [i32]
    lib.msb_shape_supports_tcgen05.argtypes = [i32, i32, i32]
    for f in (lib.msb_odeblock_workspace_bytes, lib.msb_odeblock_tape_bytes, lib.msb_odeblock_bwd_workspace_bytes):
        f.argtypes = [dp]
        f.restype = sz
    lib.msb_odeblock_forward.argtypes = [dp, vp, vp, vp, ctypes.POINTER(MsbMnistParams), vp, vp, sz, vp, sz, vp]
    lib.msb_odeblock_backward.argtypes = [dp, vp, vp, vp, vp, sz, vp, vp, vp, vp, sz, vp]
    lib.msb_odeblock_bwd_workspace_bytes_tableau.argtypes = [dp]
    lib.msb_odeblock_bwd_workspace_bytes_tableau.restype = sz
    lib.msb_odeblock_backward_tableau.argtypes = [dp, vp, vp, vp, vp, sz, vp, vp, vp, vp, vp, sz, vp]
    lib.msb_odeblock_backward_mnist_tableau.argtypes = [dp, vp, ctypes.POINTER(MsbMnistParams), vp, sz, vp,
                                                        ctypes.POINTER(MsbMnistGrads), vp, vp, sz, vp]
    lib.msb_odeblock_backward_mnist.argtypes = [dp, vp, ctypes.POINTER(MsbMnistParams), vp, sz, vp,
                                                ctypes.POINTER(MsbMnistGrads), vp, sz, vp]
    lib.msb_stem_forward.argtypes = [vp, vp, i32, vp, vp, i32, i32, i32, i32, vp]
    lib.msb_stem_backward_workspace_bytes.argtypes = [i32]
    lib.msb_stem_backward_workspace_bytes.restype = sz
    lib.msb_stem_backward.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, sz, vp]
    ddp = ctypes.POINTER(MsbDownDesc)
    for f in (lib.msb_downblock_workspace_bytes, lib.msb_downblock_tape_bytes, lib.msb_downblock_bwd_workspace_bytes):
        f.argtypes = [ddp]
        f.restype = sz
    lib.msb_downblock_forward.argtypes = [ddp, vp, vp, vp, vp, vp, vp, sz, vp, sz, vp]
    lib.msb_downblock_backward.argtypes = [ddp, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp, vp, sz, vp]
    lib.msb_act_split.argtypes = [vp, i32, vp, vp, i32, i32, i32, i32, vp]
    lib.msb_conv3x3.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, sz, vp]
    lib.msb_conv3x3_workspace_bytes.argtypes = [i32]
    lib.msb_conv3x3_workspace_bytes.restype = sz
    lib.msb_wgrad3x3.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, sz, vp]
    lib.msb_wgrad3x3_workspace_bytes.argtypes = [i32, i32]
    lib.msb_wgrad3x3_workspace_bytes.restype = sz
    lib.msb_set_option.argtypes = [ctypes.c_char_p, i32]
    lib.msb_get_option.argtypes = [ctypes.c_char_p, ctypes.POINTER(i32)]
    f32, i64 = ctypes.c_float, ctypes.c_int64
    lib.msb_attack_step.argtypes = [i32, vp, vp, vp, vp, i64, i32, i32, i32, f32, f32, i32, ctypes.POINTER(f32), vp]
    lib.msb_sgd_step.argtypes = [vp, vp, vp, i64, f32, f32, f32, f32, i32, vp]
    lib.msb_augment_batch.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, ctypes.POINTER(f32), ctypes.POINTER(f32), vp, vp]
    lib.msb_pool_fc_forward.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.msb_pool_fc_backward.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.msb_cross_entropy_forward.argtypes = [vp, vp, vp, vp, i32, i32, vp]
    lib.msb_cross_entropy_backward.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp]
    u32 = ctypes.c_uint
    lib.msb_peer_alloc.argtypes = [sz, ctypes.POINTER(vp), ctypes.c_char_p]
    lib.msb_peer_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
    lib.msb_peer_close.argtypes = [vp]
    lib.msb_peer_free.argtypes = [vp]
    lib.msb_peer_status.argtypes = [vp, ctypes.POINTER(u32), ctypes.POINTER(u32)]
    lib.msb_peer_allreduce_sgd.argtypes = [ctypes.POINTER(vp), i32, i32, i64, i64, i64, i32, vp, vp, f32, f32, f32, f32, i32, u32, vp]
    lib.msb_profile_enable.argtypes = [i32]
    lib.msb_profile_read_executed.argtypes = [i32, ctypes.POINTER(ctypes.c_double)]
    lib.msb_profile_read.argtypes = [i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                     ctypes.POINTER(ctypes.c_int64)]


def lib():
    """Load (once) and return the shared library; raise RuntimeError if it is not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "metasolver_b200: %s is missing. Build it with `python neural-ode-metasolver_b200/build.py` "
                        "(or __graft_entry__.build()). There is no CPU / cuDNN path to fall back to." % LIB_PATH)
                l = ctypes.CDLL(LIB_PATH)
                _declare(l)
                if l.msb_abi_version() != ABI_VERSION:
                    raise RuntimeError("metasolver_b200: ABI version mismatch")
                for which, cls in enumerate((MsbOdeDesc, MsbTableau, MsbMnistParams, MsbMnistGrads, MsbDownDesc)):
                    if l.msb_sizeof(which) != ctypes.sizeof(cls):
                        raise RuntimeError("metasolver_b200: struct %s is %d bytes in the binding but %d in the library"
                                           % (cls.__name__, ctypes.sizeof(cls), l.msb_sizeof(which)))
                _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("metasolver_b200 %s failed: %s" % (what, lib().msb_last_error().decode()))
