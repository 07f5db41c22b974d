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
    allocation never clobbers a live value, leaves are deduplicated, roots come first, oversized DAGs are refused); the second half checks the
    leaf addressing ptr + r * pitch + c * cstride against the strided view it stands for on random sliced / broadcast views.  No device needed."""
    import ctypes as C
    from rust_autograd_b200 import autograd as ag
    n = C.c_int()
    total = 0
    for seed in (1, 7, 2026, 99991):
        assert ag.lib().agx_fuse_selftest(3000, seed, C.byref(n)) == 0, seed
        assert 0 < n.value <= 3000
        total += n.value
    assert total < 4 * 3000          # some DAGs exceeded the leaf / register / output bounds and were refused


REFERENCE_TENSOR_OPS = """_hessian_vector_product abs acos acosh add add_n argmax argmin asin asinh assign atan atanh batch_matmul batch_matmul_t batch_norm
bernoulli bernoulli_rng ceil clip concat control_dependencies conv2d conv2d_transpose convert_to_tensor cos cosh digamma_f32 digamma_f64 dilated_conv2d
dilated_conv2d_transpose div dropout dropout_rng elu equal exp exp10 exp2 expand_dims flatten floor gather gather_common grad grad_with_default greater
greater_equal identity inv inv_sqrt jacobians leaky_relu lesser lesser_equal lgamma_f32 lgamma_f64 ln log10 log2 log_normal log_normal_rng log_softmax map
matmul max_pool2d maximum mean_all mean_squared_error minimum mul neg normalize not_equal nth_tensor ones pow random_exp random_exp_rng random_gamma
random_gamma_rng random_normal random_normal_rng random_uniform random_uniform_rng rank reduce_logsumexp reduce_max reduce_mean reduce_min reduce_prod
reduce_sum reduce_variance relu reshape scalar setdiff1d shape sigmoid sigmoid_cross_entropy sign sin sinh size slice softmax softmax_cross_entropy softplus
sparse_softmax_cross_entropy split sqrt square squeeze standard_normal standard_normal_rng standard_uniform standard_uniform_rng stop_gradient sub sum_all tan
tanh tensordot tile transpose zeros""".split()


def test_python_view_exposes_every_tensor_ops_constructor():
    """Every `pub fn` of the reference's src/tensor_ops/mod.rs (126 names, listed above so that the test does not need the reference tree) has a
    same-named constructor in the Python view of the graph-level ABI, plus the Tensor methods of src/tensor.rs (eval, show, show_shape, print,
    raw_hook, map, access_elem) and the optimizers of src/optimizers."""
    from rust_autograd_b200 import autograd as ag
    assert len(REFERENCE_TENSOR_OPS) == 126
    missing = [n for n in REFERENCE_TENSOR_OPS if not callable(getattr(ag, n, None))]
    assert not missing, missing
    for m in ("eval", "show", "show_shape", "print", "raw_hook", "map", "access_elem"):
        assert callable(getattr(ag.Tensor, m, None)), m
    for o in ("Adam", "SGD", "MomentumSGD", "AdaGrad"):
        assert hasattr(ag.optimizers, o), o


def test_rust_bindings_are_in_sync_with_the_header():
    """rust/src/tensor_ops/cuda_ffi.rs + rust/build.rs (the crate-side half of the boundary, generated by scripts/gen_rust_ffi.py): up to
    date, one `pub fn` per header function with the same number of arguments as the ctypes prototype the tests call through."""
    import subprocess
    import sys
    from rust_autograd_b200 import ffi
    assert subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_rust_ffi.py"), "--check"]).returncode == 0, "run scripts/gen_rust_ffi.py"
    rs = open(os.path.join(ROOT, "rust", "src", "tensor_ops", "cuda_ffi.rs")).read()
    decl = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (agb_[a-z0-9_]+)\((.*?)\) -> ", rs)}
    assert set(decl) == set(_declared("agb200.h"))
    for name, args in decl.items():
        if name == "agb_last_error":
            continue
        n_rust = 0 if not args.strip() else len([a for a in args.split(", ") if ":" in a])
        assert n_rust == len(ffi.SIGNATURES[name]), name
    build = open(os.path.join(ROOT, "rust", "build.rs")).read()
    for f in os.listdir(os.path.join(ROOT, "rust-autograd_b200", "csrc")):
        if f.endswith(".cu"):
            assert '"%s"' % f[:-3] in build, f
    assert "arch=compute_100a,code=sm_100a" in build
