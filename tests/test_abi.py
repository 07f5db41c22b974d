"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ag[bx]_[a-z0-9_]+)\s*\(", src)))


def _headers():
    return sorted(f for f in os.listdir(os.path.join(ROOT, "include")) if f.endswith(".h"))


def test_library_exports_every_declared_symbol():
    from rust_autograd_b200 import ffi
    lib = ctypes.CDLL(ffi.LIB_PATH)
    missing = []
    n = 0
    for h in _headers():
        for name in _declared(h):
            n += 1
            try:
                getattr(lib, name)
            except AttributeError:
                missing.append("%s:%s" % (h, name))
    assert n > 50
    assert not missing, "declared in include/*.h but not exported: %s" % missing


def test_ctypes_prototypes_cover_the_header():
    from rust_autograd_b200 import ffi
    declared = set(_declared("agb200.h")) - {"agb_last_error"}
    assert declared == set(ffi.SIGNATURES), (declared ^ set(ffi.SIGNATURES))


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly (agb_init -> AGB_ERR_CUDA), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rust_autograd_b200 as agb
    with pytest.raises(agb.OpError) as e:
        agb.Device(0)
    assert e.value.code == agb.ffi.ERR_CUDA


def test_error_codes_mirror_op_error():
    """src/op.rs:67-73: five OpError variants -> codes 1..5"""
    from rust_autograd_b200 import ffi
    assert [ffi.OpError.NAMES[i] for i in range(1, 6)] == ["NdArrayError", "IncompatibleShape", "TypeUnsupported", "InvalidDims", "OutOfBounds"]


def test_fused_program_compiler_on_the_host():
    """engine/fuse.cc: random expression DAGs (unary / binary / immediate forms, shared sub-expressions, multi-consumer nodes, two roots) are
    compiled into agb_fused_ewise programs and interpreted on the host: every stored register must hold its node's value (register
    allocation never clobbers a live value, leaves are deduplicated, roots come first, oversized DAGs are refused).  No device needed."""
    import ctypes as C
    from rust_autograd_b200 import autograd as ag
    n = C.c_int()
    total = 0
    for seed in (1, 7, 2026, 99991):
        assert ag.lib().agx_fuse_selftest(3000, seed, C.byref(n)) == 0, seed
        assert 0 < n.value <= 3000
        total += n.value
    assert total < 4 * 3000          # some DAGs exceeded the leaf / register / output bounds and were refused
