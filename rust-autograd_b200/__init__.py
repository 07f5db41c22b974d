"""rust-autograd_b200 — B200-native (sm_100a) execution backend for rust-autograd's op-evaluation hot path.

Layout:  csrc/ (CUDA kernels + C ABI + C++ host engine)  ->  lib/libagb200.so  ->  ffi.py (ctypes prototypes)
         device.py (kernel-level handle)  +  autograd.py (host-side mirror of the reference's Graph/Evaluator API).
Import name: ``rust_autograd_b200`` (the hyphenated directory cannot be imported directly).
"""
from . import ffi  # noqa: F401
from .ffi import OpError, MATH_3XTF32, MATH_TF32, MATH_FP32  # noqa: F401
from .device import Device, DArray  # noqa: F401

__all__ = ["ffi", "Device", "DArray", "OpError", "MATH_3XTF32", "MATH_TF32", "MATH_FP32"]
