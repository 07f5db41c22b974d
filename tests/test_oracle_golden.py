"""Pins the oracle (oracle/ref_ops.py) against every known-answer vector the reference's own tests hold for the hot path
(tests/golden/reference_kats.json; each entry cites the reference file:line), and cross-checks the unpinned
restatements (conv family, xent, optimizers) by finite differences / algebraic identities."""
import json
import os

import numpy as np
import pytest

from oracle import ref_ops as R

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


def test_im2col_batch_kat():
    k = KATS["im2col_batch"]
    x = np.tile(np.arange(k["xch"] * k["xh"] * k["xw"], dtype=np.float32).reshape(1, k["xch"], k["xh"], k["xw"]), (k["batch"], 1, 1, 1))
    cols = R.im2col(x, k["kh"], k["kw"], k["pad"], k["stride"], k["dilation"])
    assert cols.ravel().tolist() == [float(v) for v in k["expected"]]


def test_max_pool_kat():
    k = KATS["max_pool"]
    y, idx, idx_i = R.max_pool2d(np.array(k["x"], np.float32).reshape(1, 1, k["h"], k["w"]), k["size"], k["pad"], k["stride"])
    assert y.ravel().tolist() == k["output"] and idx.ravel().tolist() == k["argmax"]


def test_deconv_kat():
    k = KATS["deconv"]
    out = R.conv2d_transpose(np.ones((k["batch"], k["ych"], k["yh"], k["yw"]), np.float32), np.ones((k["ych"], k["xch"], k["kh"], k["kw"]), np.float32), k["pad"], k["stride"])
    assert out.shape == (2, 3, 3, 3)
    for b in range(2):
        for c in range(3):
            assert out[b, c].ravel().tolist() == [float(v) for v in k["expected_per_channel"]]


@pytest.mark.parametrize("k", KATS["argmax"])
def test_argmax_kats(k):
    assert np.array_equal(R.arg_reduce(np.array(k["x"], np.float32), k["axis"], False, True), np.array(k["expected"], np.float32))


@pytest.mark.parametrize("k", KATS["argmin"])
def test_argmin_kats(k):
    assert np.array_equal(R.arg_reduce(np.array(k["x"], np.float32), k["axis"], False, False), np.array(k["expected"], np.float32))


@pytest.mark.parametrize("k", KATS["matmul"])
def test_matmul_kats(k):
    assert np.array_equal(R.matmul(np.array(k["a"], np.float32), np.array(k["b"], np.float32), k["ta"], k["tb"]), np.array(k["expected"], np.float32))


@pytest.mark.parametrize("k", KATS["batch_matmul"])
def test_batch_matmul_kats(k):
    assert np.array_equal(R.batch_matmul(np.array(k["a"], np.float32), np.array(k["b"], np.float32), k["ta"], k["tb"]), np.array(k["expected"], np.float32))


@pytest.mark.parametrize("k", KATS["compare"])
def test_compare_kats(k):
    assert np.array_equal(R.compare(k["op"], k["a"], k["b"]), np.array(k["expected"], np.float32))


@pytest.mark.parametrize("k", KATS["reduce"])
def test_reduce_kats(k):
    assert np.array_equal(R.reduce(k["op"], np.array(k["x"], np.float32), k["axes"], False), np.array(k["expected"], np.float32))


def test_misc_kats():
    k = KATS["sum_all"]
    assert R.sum_all(np.array(k["x"], np.float32)) == k["expected"]
    k = KATS["add_n"]
    assert np.array_equal(R.add_n([np.ones(k["shape"], np.float32)] * k["n"]), np.array(k["expected"], np.float32))
    k = KATS["clip"]
    assert R.unary("clip", k["x"], k["min"], k["max"]).tolist() == k["expected"]
    k = KATS["sign"]
    assert R.unary("sign", k["x"]).tolist() == k["expected"]


# ---- unpinned restatements: self-consistency ----
def _naive_conv(x, w, pad, stride, dil):
    B, C, H, W = x.shape
    O, _, kh, kw = w.shape
    yh, yw = R.conv_out_size(H, kh, pad, stride, dil), R.conv_out_size(W, kw, pad, stride, dil)
    y = np.zeros((B, O, yh, yw))
    for b in range(B):
        for o in range(O):
            for i in range(yh):
                for j in range(yw):
                    s = 0.0
                    for c in range(C):
                        for p in range(kh):
                            for q in range(kw):
                                yy, xx = i * stride - pad + p * dil, j * stride - pad + q * dil
                                if 0 <= yy < H and 0 <= xx < W:
                                    s += float(x[b, c, yy, xx]) * float(w[o, c, p, q])
                    y[b, o, i, j] = s
    return y


@pytest.mark.parametrize("pad,stride,dil", [(0, 1, 1), (1, 1, 1), (1, 2, 1), (2, 1, 2), (0, 2, 2)])
def test_conv2d_matches_direct_loops(pad, stride, dil):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 7, 6)).astype(np.float32)
    w = rng.standard_normal((4, 3, 3, 2)).astype(np.float32)
    np.testing.assert_allclose(R.conv2d(x, w, pad, stride, dil), _naive_conv(x, w, pad, stride, dil), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("pad,stride,dil", [(0, 1, 1), (1, 1, 1), (1, 2, 1), (2, 1, 2)])
def test_conv_grads_are_adjoints(pad, stride, dil):
    """<conv(x,w), gy> == <x, conv_transpose(gy,w)> == <w, filter_grad(x,gy)> (the reference checks the same thing by FD,
    tests/test_tensor_ops_grad.rs:1358-1493)."""
    rng = np.random.default_rng(1)
    H = 7 if stride == 1 else 2 * 3 + (dil * 2 + 1) - 2 * pad - 2 + 0   # sizes where transpose exactly inverts the shape
    x = rng.standard_normal((2, 3, 9, 9)).astype(np.float32)
    w = rng.standard_normal((4, 3, 3, 3)).astype(np.float32)
    y = R.conv2d(x, w, pad, stride, dil)
    gy = rng.standard_normal(y.shape).astype(np.float32)
    gx = R.conv2d_transpose(gy, w, pad, stride, dil)
    gw = R.conv2d_filter_grad(x, gy, w.shape, pad, stride, dil)
    lhs = float((y.astype(np.float64) * gy).sum())
    if gx.shape == x.shape:
        assert abs(lhs - float((x.astype(np.float64) * gx).sum())) < 1e-3 * max(1, abs(lhs))
    assert abs(lhs - float((w.astype(np.float64) * gw).sum())) < 1e-3 * max(1, abs(lhs))


def test_xent_and_softmax_identities():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((5, 7)).astype(np.float32) * 3
    t = rng.integers(0, 7, 5).astype(np.float32)
    loss, log_x = R.sparse_softmax_cross_entropy(x, t)
    assert loss.shape == (5, 1) and log_x.shape == (5, 7)
    np.testing.assert_allclose(np.exp(log_x), R.softmax(x, 1), rtol=1e-5)
    np.testing.assert_allclose(R.log_softmax(x, 1), log_x, rtol=1e-6, atol=1e-6)
    onehot = np.eye(7, dtype=np.float32)[t.astype(int)]
    l2, _ = R.softmax_cross_entropy(x, onehot)
    np.testing.assert_allclose(l2, loss[:, 0], rtol=1e-6)
    # finite-difference check of the fused backward
    g = R.sparse_softmax_cross_entropy_grad(log_x, t, np.ones((5, 1), np.float32))
    eps = 1e-3
    for (i, j) in [(0, 0), (2, 3), (4, 6)]:
        xp, xm = x.copy(), x.copy()
        xp[i, j] += eps
        xm[i, j] -= eps
        fd = (R.sparse_softmax_cross_entropy(xp, t)[0].astype(np.float64).sum() - R.sparse_softmax_cross_entropy(xm, t)[0].astype(np.float64).sum()) / (2 * eps)
        assert abs(fd - g[i, j]) < 2e-3


def test_adam_first_step_is_alpha_sized():
    """t starts at 1 (optimizers/adam.rs:97): after the first update |dp| ~= alpha for any gradient scale."""
    p, g = np.ones(4, np.float32), np.array([1e-3, 1.0, -5.0, 100.0], np.float32)
    p2, m, v, t = R.adam_update(p, g, np.zeros(4, np.float32), np.zeros(4, np.float32), np.float32(1.0))
    np.testing.assert_allclose(p - p2, 1e-3 * np.sign(g), rtol=1e-4)
    assert t == 2.0


def test_pool_grad_roundtrip():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3, 6, 6)).astype(np.float32)
    y, idx, _ = R.max_pool2d(x, 2, 0, 2)
    assert np.array_equal(x.ravel()[idx.astype(np.int64)], y)
    gx = R.max_pool2d_grad(np.ones_like(y), idx, 2, 0, 2)
    assert gx.sum() == y.size and gx.shape == x.shape
    assert np.array_equal(R.max_pool2d_grad_grad(x, idx, 2, 0, 2), y)
