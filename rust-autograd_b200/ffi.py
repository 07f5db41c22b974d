"""ctypes binding of the kernel-level C ABI (``include/agb200.h``).

This mirrors, in Python, what the Rust side would write in ``src/tensor_ops/cuda_ffi.rs`` (the sibling of the
reference's ``src/tensor_ops/blas_ffi.rs:17-158``): plain ``extern "C"`` declarations, no logic.  The product
path has no CPU fallback: if ``libagb200.so`` is missing or no B200 is present, loading / ``Device()`` raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGB200_LIB") or os.path.join(_HERE, "lib", "libagb200.so")      # AGB200_LIB: another build of the same library (A/B timing of two builds on one box)
MAX_RANK = 8

# status codes (include/agb200.h)
OK, ERR_NDARRAY, ERR_INCOMPATIBLE_SHAPE, ERR_TYPE_UNSUPPORTED, ERR_INVALID_DIMS, ERR_OUT_OF_BOUNDS = range(6)
ERR_CUDA, ERR_NCCL, ERR_UNSUPPORTED = 100, 101, 102
MATH_3XTF32, MATH_TF32, MATH_FP32 = 0, 1, 2

U_OPS = ["copy", "abs", "neg", "square", "inv", "invsqrt", "sign", "floor", "ceil", "sqrt", "pow", "ln", "log2", "log10",
         "exp", "exp2", "exp10", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh",
         "atanh", "sigmoid", "relu", "softplus", "elu", "clip", "scale", "add_scalar", "rsub_scalar", "rdiv_scalar", "lgamma", "digamma"]
B_OPS = ["add", "sub", "mul", "div", "eq", "ne", "gt", "lt", "ge", "le", "max", "min", "elu_grad", "clip_grad",
         "sigmoid_xent", "relu_grad"]
R_OPS = ["sum", "mean", "prod", "min", "max"]
RAND = {n: i for i, n in enumerate(["uniform", "normal", "bernoulli", "exp", "log_normal", "gamma"])}
U = {n: i for i, n in enumerate(U_OPS)}
B = {n: i for i, n in enumerate(B_OPS)}
R = {n: i for i, n in enumerate(R_OPS)}


class AgbTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("rank", C.c_int32), ("shape", C.c_int64 * MAX_RANK), ("stride", C.c_int64 * MAX_RANK)]


class AgbFuseLeaf(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("pitch", C.c_int64), ("cstride", C.c_int64), ("reg", C.c_int32)]


class AgbFuseInstr(C.Structure):
    _fields_ = [("kind", C.c_int32), ("op", C.c_int32), ("dst", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("p0", C.c_float)]


class AgbFuseOut(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("pitch", C.c_int64), ("reg", C.c_int32)]


F_UNARY, F_BINARY, F_BINARY_IMM_B, F_BINARY_IMM_A = 0, 1, 2, 3


class OpError(RuntimeError):
    """Mirrors ``OpError`` (reference ``src/op.rs:67-73``); ``code`` is the C status."""
    NAMES = {1: "NdArrayError", 2: "IncompatibleShape", 3: "TypeUnsupported", 4: "InvalidDims", 5: "OutOfBounds",
             100: "CudaError", 101: "NcclError", 102: "Unsupported"}

    def __init__(self, code, msg):
        self.code = code
        self.kind = self.NAMES.get(code, "Error%d" % code)
        super().__init__("%s: %s" % (self.kind, msg))


_P = C.c_void_p
_T = C.POINTER(AgbTensor)
_f, _i, _i64, _u64, _sz = C.c_float, C.c_int, C.c_int64, C.c_uint64, C.c_size_t

# name -> argtypes; every function returns int status except agb_last_error
SIGNATURES = {
    "agb_init": [_i, C.POINTER(_P)], "agb_destroy": [_P], "agb_device_count": [C.POINTER(_i)],
    "agb_sm_count": [_P, C.POINTER(_i)], "agb_set_math_mode": [_P, _i], "agb_get_math_mode": [_P, C.POINTER(_i)],
    "agb_set_deterministic": [_P, _i], "agb_get_deterministic": [_P, C.POINTER(_i)],
    "agb_launch_count": [_P, C.POINTER(_i64)],
    "agb_prof_enable": [_P, _i], "agb_prof_reset": [_P], "agb_prof_collect": [_P, _i, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(C.c_double)],
    "agb_alloc": [_P, _sz, C.POINTER(_P)], "agb_free": [_P, _P], "agb_trim": [_P],
    "agb_mem_stats": [_P, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)],
    "agb_host_alloc": [_sz, C.POINTER(_P)], "agb_host_free": [_P],
    "agb_h2d": [_P, _P, _P, _sz], "agb_d2h": [_P, _P, _P, _sz], "agb_d2d": [_P, _P, _P, _sz], "agb_memset0": [_P, _P, _sz],
    "agb_sync": [_P], "agb_flush_l2": [_P], "agb_arena_pin": [_P, _i], "agb_stage_mark": [_P], "agb_stage_h2d": [_P, _P, _P, _sz], "agb_stage_wait": [_P], "agb_stage_d2h": [_P, _P, _P, _sz, C.POINTER(_P)], "agb_event_sync": [_P],
    "agb_event_create": [C.POINTER(_P)], "agb_event_destroy": [_P], "agb_event_record": [_P, _P],
    "agb_event_elapsed_ms": [_P, _P, C.POINTER(_f)],
    "agb_graph_begin": [_P], "agb_graph_end": [_P, C.POINTER(_P)], "agb_graph_launch": [_P, _P], "agb_graph_destroy": [_P],
    "agb_gemm_f32": [_P, _i, _i, _T, _T, _T, _f],
    "agb_conv2d_fprop_f32": [_P, _T, _T, _T, _i, _i, _i], "agb_conv2d_fprop_fused_f32": [_P, _T, _T, _P, _i, _T, _i, _i, _i], "agb_conv2d_fprop_pool_f32": [_P, _T, _T, _P, _i, _T, _P, _i, _i, _i], "agb_conv2d_dgrad_f32": [_P, _T, _T, _T, _i, _i, _i],
    "agb_conv2d_dgrad_fused_f32": [_P, _T, _T, _T, _P, _T, _i, _i, _i],
    "agb_conv2d_fprop_fused_bits_f32": [_P, _T, _T, _P, _i, _T, _P, C.POINTER(_i), _i, _i, _i], "agb_conv2d_dgrad_fused_bits_f32": [_P, _T, _T, _T, _P, _P, _T, _i, _i, _i], "agb_maxpool2d_bwd_fused": [_P, _T, _P, _P, _P, _P, _T, _i, _i],
    "agb_conv2d_wgrad_f32": [_P, _T, _T, _T, _i, _i, _i], "agb_conv_prefers_channels_last": [_i, _i, _i, _i, _i, _i], "agb_im2col_f32": [_P, _T, _T, _i, _i, _i, _i, _i],
    "agb_maxpool2d_fwd": [_P, _T, _T, _P, _P, _i, _i, _i], "agb_maxpool2d_bwd": [_P, _T, _P, _P, _T],
    "agb_maxpool2d_gradgrad": [_P, _T, _P, _P, _T],
    "agb_convert_i32_f32": [_P, _P, _P, _i64],
    "agb_unary": [_P, _i, _f, _f, _T, _T], "agb_binary": [_P, _i, _f, _f, _T, _T, _T],
    "agb_add_n": [_P, _i, C.POINTER(_T), _T], "agb_fill": [_P, _T, _f],
    "agb_fused_ewise": [_P, _i64, _i64, _i, _P, _i, _P, _i, _P], "agb_copy_strided": [_P, _T, _T],
    "agb_concat_rows": [_P, _i, _P, _P, _i64, _i64, _P],
    "agb_dropout": [_P, _T, _T, _T, _f, _u64, _u64], "agb_random": [_P, _i, _f, _f, _u64, _u64, _T],
    "agb_stream_cell_bytes": [], "agb_dropout_stream": [_P, _T, _T, _T, _f, _u64, _u64, _P], "agb_random_stream": [_P, _i, _f, _f, _u64, _u64, _P, _T],
    "agb_reduce": [_P, _i, _P, _P, _i64, _i64, _i64], "agb_argreduce": [_P, _i, _P, _P, _i64, _i64, _i64],
    "agb_softmax": [_P, _P, _P, _i64, _i64, _i64], "agb_log_softmax": [_P, _P, _P, _i64, _i64, _i64],
    "agb_logsumexp": [_P, _P, _P, _i64, _i64, _i64],
    "agb_sparse_xent_fwd": [_P, _P, _P, _P, _P, _i64, _i64], "agb_sparse_xent_bwd": [_P, _P, _P, _P, _i64, _P, _i64, _i64],
    "agb_softmax_xent_fwd": [_P, _P, _P, _P, _P, _i64, _i64],
    "agb_gather": [_P, _P, _P, _P, _i64, _i64, _i64, _i64, _i], "agb_gather_grad": [_P, _P, _P, _P, _i64, _i64, _i64, _i64], "agb_scatter_add": [_P, _P, _P, _P, _i64, _i64, _i64, _i64],
    "agb_multi_tensor_adam": [_P, _i, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                              C.POINTER(_i64), _f, _f, _f, _f, _f],
    "agb_multi_tensor_sgd": [_P, _i, C.POINTER(_P), C.POINTER(_P), C.POINTER(_i64), _f, _f],
    "agb_multi_tensor_momentum": [_P, _i, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_i64), _f, _f, _f],
    "agb_multi_tensor_adagrad": [_P, _i, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_i64), _f, _f],
    "agb_nccl_unique_id": [_P], "agb_nccl_init": [_P, _i, _i, _P], "agb_allreduce_sum": [_P, _P, _i64],
    "agb_nccl_destroy": [_P], "agb_allreduce_sum_async": [_P, _P, _i64], "agb_allreduce_wait": [_P],
}

_lib = None


def load_library():
    """dlopen libagb200.so and attach prototypes.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, args in SIGNATURES.items():
        if not hasattr(lib, name) and os.environ.get("AGB200_LIB"):
            continue                       # an older build under comparison may lack the newest entry points
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.agb_last_error.argtypes = []
    lib.agb_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise OpError(status, load_library().agb_last_error().decode("utf-8", "replace"))


def make_tensor(ptr, shape, strides=None):
    t = AgbTensor()
    t.ptr = ptr
    t.rank = len(shape)
    assert t.rank <= MAX_RANK
    if strides is None:
        strides = []
        s = 1
        for d in reversed(shape):
            strides.append(s)
            s *= d
        strides = list(reversed(strides))
    for i, (d, s) in enumerate(zip(shape, strides)):
        t.shape[i] = int(d)
        t.stride[i] = int(s)
    return t


def np_ptr(a):
    assert isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
