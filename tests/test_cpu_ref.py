"""CPU test: the f32 C restatement that serves as the timing baseline (oracle/cpu_ref.c, bench.py --impl reference) computes what the
numpy oracle computes — per-op and over two Adam steps of a small VGG-style stack — so the baseline it times is the real algorithm
(im2col + per-sample sgemm, sequential filter gradient, scalar pooling loops, five-pass Adam), not a shortcut."""
import numpy as np
import pytest

from oracle import cpu_ref as CR, ref_graph as OG, ref_ops as R
from rust_autograd_b200 import workloads as W

LAYERS = [(3, 32), (32, 32), "pool", (32, 64), "pool"]


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("builtin", [False, True], ids=["openblas", "builtin_sgemm"])
def test_ops_match_numpy_oracle(builtin):
    lib = CR.load()
    ops = CR.Ops()
    saved = None
    if builtin:                                   # the packed AVX2 kernel that is used when numpy's OpenBLAS cannot be found
        import ctypes as C
        saved = getattr(lib, "_openblas", None)
        lib.cr_set_sgemm(None)
    try:
        rng = np.random.default_rng(0)
        x = rng.standard_normal((3, 5, 9, 11)).astype(np.float32)
        w = rng.standard_normal((7, 5, 3, 3)).astype(np.float32)
        y, cols = ops.conv2d(x, w, 1)
        assert rel(y, R.conv2d(x, w, 1, 1, 1)) <= 1e-5
        gy = rng.standard_normal(y.shape).astype(np.float32)
        assert rel(ops.conv2d_filter_grad(cols, gy, w.shape), R.conv2d_filter_grad(x, gy, w.shape, 1, 1, 1)) <= 1e-5
        assert rel(ops.conv2d_transpose(gy, w, 1, 9, 11), R.conv2d_transpose(gy, w, 1, 1, 1)) <= 1e-5
        a, b = rng.standard_normal((37, 53)).astype(np.float32), rng.standard_normal((53, 29)).astype(np.float32)
        assert rel(ops.matmul(a, b), R.matmul(a, b)) <= 1e-5
        assert rel(ops.matmul(a, rng.standard_normal((37, 9)).astype(np.float32), ta=True).shape, (53, 9)) == 0
        xc = np.ascontiguousarray(x[:, :, :8, :10])
        p, idx = ops.max_pool2d(xc, 2, 2)
        pr, ir, _ = R.max_pool2d(xc, 2, 0, 2)
        assert np.array_equal(p, pr) and np.array_equal(idx, ir)
    finally:
        if builtin and saved is not None:
            import ctypes as C
            lib.cr_set_sgemm(C.cast(saved.scipy_cblas_sgemm64_, C.c_void_p))


def test_training_steps_match_numpy_oracle():
    rng0 = np.random.default_rng(3)
    x = rng0.standard_normal((4, 3, 32, 32)).astype(np.float32)
    y = rng0.integers(0, 10, (4, 1)).astype(np.float32)
    tr = CR.VggTrainer(CR.vgg_params(np.random.default_rng(0), 32, LAYERS), LAYERS, 32)
    env = OG.VariableEnvironment()
    W.vgg_init(env, np.random.default_rng(0), size=32, layers=LAYERS)
    adam = OG.optimizers.Adam.default("adam", env.default_namespace().current_var_ids(), env)
    for _ in range(2):
        loss, _ = tr.step(x, y)

        def step(g):
            l, _ = W.vgg_loss(OG, g, size=32, layers=LAYERS)
            params, grads = OG.optimizers.grad_helper([l], g.default_namespace())
            return float(np.asarray(g.evaluator().push(l).push(adam.get_update_op(params, grads, g)).feed("x", x).feed("y", y).run()[0].unwrap()).ravel()[0])
        assert abs(loss - env.run(step)) <= 1e-5 * abs(loss)
    for i, n in enumerate(["conv0_w", "conv0_b", "conv1_w", "conv1_b", "conv2_w", "conv2_b", "fc_w", "fc_b"]):
        assert rel(tr.p[n], env.get_array_by_id(i)) <= 1e-4, n      # Adam's normalised step amplifies f32-vs-f64 differences on tiny gradients
