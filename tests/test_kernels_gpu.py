"""-m gpu parity tests, kernel level: every entry point of include/agb200.h (called through the C ABI via ctypes) against
the oracle (oracle/ref_ops.py) on the same seeded inputs, plus the reference's own known-answer vectors replayed on the GPU.

Tolerances (BASELINE.json north_star): bit-exact for index/integer-valued outputs; f32 within 1e-5 relative for
3xTF32 / fp32 GEMM-conv and for elementwise / reduction kernels; 1e-2 relative in TF32 mode.  For contractions
"relative" is taken against the largest output magnitude (forward-error convention for sums with cancellation)."""
import json
import os

import numpy as np
import pytest

from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
TOL = {0: 1e-5, 1: 1e-2, 2: 1e-5}      # math mode -> relative tolerance
MODES = [0, 1, 2]                        # 3xTF32, TF32, FP32


def rel_err(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def close(got, ref, tol=1e-5):
    np.testing.assert_allclose(np.asarray(got, np.float64), np.asarray(ref, np.float64), rtol=tol, atol=tol * 1e-1)


# ------------------------------------------------------------------------------------------------ runtime
def test_runtime_roundtrip_and_arena(dev):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((37, 129)).astype(np.float32)
    d = dev.upload(a)
    assert np.array_equal(d.numpy(), a)
    assert np.array_equal(d.transpose().numpy(), a.T)
    assert np.array_equal(d.slice(1, 5, 77).numpy(), a[:, 5:77])
    assert dev.sm_count() == 148
    assert dev.launch_count() > 0


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("k", KATS["matmul"])
@pytest.mark.parametrize("mode", MODES)
def test_matmul_kats(dev, k, mode):
    dev.set_math_mode(mode)
    c = dev.gemm(dev.upload(np.array(k["a"], np.float32)), dev.upload(np.array(k["b"], np.float32)), k["ta"], k["tb"])
    assert np.array_equal(c.numpy(), np.array(k["expected"], np.float32))


@pytest.mark.parametrize("k", KATS["batch_matmul"])
def test_batch_matmul_kats(dev, k):
    dev.set_math_mode(0)
    c = dev.gemm(dev.upload(np.array(k["a"], np.float32)), dev.upload(np.array(k["b"], np.float32)), k["ta"], k["tb"])
    assert np.array_equal(c.numpy(), np.array(k["expected"], np.float32))


GEMM_SHAPES = [(200, 10, 784), (784, 10, 200), (200, 784, 10), (128, 128, 128), (256, 512, 64), (130, 70, 33), (257, 129, 100),
               (512, 1024, 2048), (1, 1, 1), (5, 3, 7), (128, 4096, 1024), (1000, 1000, 1000), (64, 64, 32), (16, 64, 40),
               (2048, 640, 96), (2304, 512, 64),        # M >= 2048, N >= 512, dense: both operands pre-split in global memory in 3xTF32 mode (no splitter warps)
               (256, 10, 65536), (64, 10, 5000),        # skinny classifier shapes: split-K path
               (65536, 10, 256), (256, 65536, 10), (300, 12, 40000), (32, 50000, 6)]     # the classifier's gradients: thin N / thin K with a 40-byte pitch (padded copies -> tcgen05)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_matmul_vs_oracle(dev, m, n, k, ta, tb, mode):
    dev.set_math_mode(mode)
    rng = np.random.default_rng(m * 31 + n * 7 + k)
    a = rng.standard_normal((k, m) if ta else (m, k)).astype(np.float32)
    b = rng.standard_normal((n, k) if tb else (k, n)).astype(np.float32)
    c = dev.gemm(dev.upload(a), dev.upload(b), ta, tb).numpy()
    assert rel_err(c, R.matmul(a, b, ta, tb)) <= TOL[mode]


@pytest.mark.parametrize("mode", MODES)
def test_matmul_transposed_views_and_beta(dev, mode):
    """Transpose is a stride permutation (math_ops.rs:448): the kernel must consume such views without copies;
    beta=1 accumulates (used by the conv filter-grad batch loop, conv2d.rs:703-722)."""
    dev.set_math_mode(mode)
    rng = np.random.default_rng(5)
    a, b = rng.standard_normal((96, 160)).astype(np.float32), rng.standard_normal((96, 192)).astype(np.float32)
    da, db = dev.upload(a), dev.upload(b)
    c = dev.gemm(da.transpose(), db)                       # [160,96] x [96,192]
    assert rel_err(c.numpy(), R.matmul(a.T, b)) <= TOL[mode]
    c2 = dev.gemm(da.transpose(), db, out=c, beta=1.0)
    assert rel_err(c2.numpy(), 2 * R.matmul(a.T, b).astype(np.float64)) <= TOL[mode]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("batch,m,n,k", [((3,), 64, 128, 96), ((2, 3), 33, 17, 9), ((4,), 256, 256, 256)])
def test_batch_matmul_vs_oracle(dev, batch, m, n, k, mode):
    dev.set_math_mode(mode)
    rng = np.random.default_rng(11)
    a, b = rng.standard_normal(batch + (m, k)).astype(np.float32), rng.standard_normal(batch + (n, k)).astype(np.float32)
    c = dev.gemm(dev.upload(a), dev.upload(b), False, True).numpy()
    assert rel_err(c, R.batch_matmul(a, b, False, True)) <= TOL[mode]


def test_matmul_errors(dev):
    import rust_autograd_b200 as agb
    with pytest.raises(agb.OpError) as e:
        dev.gemm(dev.upload(np.zeros((3, 4), np.float32)), dev.upload(np.zeros((5, 6), np.float32)))
    assert e.value.kind == "IncompatibleShape"          # dot_ops.rs:580-584


def test_gemm_linearity_full_size(dev):
    """Size-independent property at BASELINE's largest microbench size (8192^3): C(a, b1+b2) == C(a,b1) + C(a,b2)."""
    dev.set_math_mode(0)
    n = 8192
    rng = np.random.default_rng(3)
    a = dev.upload(rng.standard_normal((n, n)).astype(np.float32))
    b1h, b2h = rng.standard_normal((n, n)).astype(np.float32), rng.standard_normal((n, n)).astype(np.float32)
    b1, b2 = dev.upload(b1h), dev.upload(b2h)
    c12 = dev.gemm(a, dev.binary("add", b1, b2))
    c1 = dev.gemm(a, b1)
    c = dev.gemm(a, b2, out=c1, beta=1.0)
    s = dev.reduce("max", dev.unary("abs", dev.binary("sub", c12, c)).reshape((n * n,)), 0).numpy()
    scale = dev.reduce("max", dev.unary("abs", c12).reshape((n * n,)), 0).numpy()
    assert float(s) <= 2e-5 * float(scale)


# ------------------------------------------------------------------------------------------------ conv family
def test_im2col_kat(dev):
    k = KATS["im2col_batch"]
    x = np.tile(np.arange(k["xch"] * k["xh"] * k["xw"], dtype=np.float32).reshape(1, k["xch"], k["xh"], k["xw"]), (k["batch"], 1, 1, 1))
    cols = dev.im2col(dev.upload(x), k["kh"], k["kw"], k["pad"], k["stride"], k["dilation"]).numpy()
    assert cols.ravel().tolist() == [float(v) for v in k["expected"]]


@pytest.mark.parametrize("mode", MODES)
def test_deconv_kat(dev, mode):
    dev.set_math_mode(mode)
    k = KATS["deconv"]
    out = dev.conv2d_transpose(dev.upload(np.ones((k["batch"], k["ych"], k["yh"], k["yw"]), np.float32)),
                               dev.upload(np.ones((k["ych"], k["xch"], k["kh"], k["kw"]), np.float32)), k["pad"], k["stride"]).numpy()
    assert out.shape == (2, 3, 3, 3)
    assert np.array_equal(out, np.tile(np.array(k["expected_per_channel"], np.float32).reshape(1, 1, 3, 3), (2, 3, 1, 1)))


CONV_CASES = [  # B, C, H, W, O, kh, kw, pad, stride, dil
    (2, 1, 28, 28, 32, 3, 3, 1, 1, 1),      # cnn_mnist conv1
    (3, 32, 14, 14, 64, 3, 3, 1, 1, 1),     # cnn_mnist conv2
    (2, 3, 9, 7, 4, 3, 2, 0, 1, 1),
    (2, 3, 9, 9, 5, 3, 3, 1, 2, 1),
    (2, 4, 10, 10, 6, 3, 3, 2, 1, 2),
    (1, 2, 5, 5, 3, 1, 1, 0, 1, 1),
    (2, 64, 32, 32, 64, 3, 3, 1, 1, 1),     # VGG-like tile-sized layer
    (2, 64, 16, 16, 128, 3, 3, 1, 1, 1),
    (1, 128, 8, 8, 256, 3, 3, 1, 1, 1),
    (2, 3, 32, 32, 64, 3, 3, 1, 1, 1),      # VGG L0 (C=3)
    (2, 16, 12, 12, 24, 5, 5, 2, 1, 1),
    (2, 8, 15, 15, 8, 3, 3, 1, 2, 1),
    # shapes that take the tcgen05 implicit-GEMM path (stride 1, W >= 32, C, O >= 32)
    (2, 32, 32, 32, 32, 3, 3, 1, 1, 1),
    (1, 64, 64, 64, 128, 3, 3, 1, 1, 1),
    (2, 128, 32, 32, 256, 3, 3, 1, 1, 1),
    (1, 256, 32, 32, 256, 3, 3, 1, 1, 1),
    (2, 64, 36, 40, 64, 3, 3, 1, 1, 1),     # partial pixel tiles
    (2, 32, 40, 40, 32, 3, 3, 2, 1, 2),     # dilation 2
    (1, 32, 32, 32, 64, 5, 5, 2, 1, 1),     # 5x5
    (1, 32, 34, 34, 32, 3, 3, 0, 1, 1),     # no padding
    (3, 48, 32, 64, 96, 3, 3, 1, 1, 1),     # channel counts that are not multiples of 32 / 64
    # wide feature maps (output width >= 128): persistent halo-reusing kernel in TF32 mode
    (1, 64, 6, 128, 64, 3, 3, 1, 1, 1),
    (2, 32, 5, 160, 48, 3, 3, 1, 1, 1),     # odd row count, partial second 128-pixel strip, 48 channels
    (1, 64, 4, 256, 128, 3, 3, 1, 1, 1),    # TN = 128
    (1, 32, 8, 130, 32, 3, 3, 0, 1, 1),     # no padding
    (1, 32, 9, 136, 32, 3, 3, 2, 1, 2),     # dilation 2
    (3, 96, 7, 128, 160, 3, 3, 1, 1, 1),    # three c-blocks, two o-tiles (the second partial)
    (1, 32, 6, 128, 32, 5, 5, 2, 1, 1),     # 5x5
    # enough 8-row patches to fill the machine: two M-tiles per CTA sharing the filter tiles (TF32 mode)
    (16, 128, 44, 64, 128, 3, 3, 1, 1, 1),  # TN = 128, last patch has 4 of 8 rows
    (40, 64, 28, 32, 256, 3, 3, 1, 1, 1),   # TN = 256
    # narrow maps (output width 16 / 32 / 64, Cout <= 128) with enough tiles: column-copies window kernel (TF32 mode)
    (40, 64, 30, 32, 96, 3, 3, 1, 1, 1),    # yw = 32: 8-row tiles, last tile 6 rows, 96 of 128 channels
    (80, 32, 24, 16, 64, 3, 3, 1, 1, 1),    # yw = 16: 16-row tiles, TN = 64
    (40, 32, 32, 32, 32, 3, 3, 2, 1, 2),    # dilation 2
    (40, 40, 32, 32, 64, 3, 3, 1, 1, 1),    # Cin = 40: three 16-channel blocks, the last one partial
    (38, 64, 16, 64, 128, 3, 3, 1, 1, 1),   # yw = 64
    # narrow maps on the per-tap kernel: whole rows / several images per 128-lane tile
    (5, 64, 7, 7, 64, 3, 3, 1, 1, 1),       # 7x7: two images per tile, odd batch
    (3, 32, 14, 14, 256, 3, 3, 1, 1, 1),    # 14x14: 9 + 5 rows
    (2, 512, 7, 7, 512, 3, 3, 1, 1, 1),     # ResNet-style deep layer
    (4, 32, 5, 6, 32, 3, 3, 1, 1, 1),       # 5x6
    # strided convolutions: tensor-core forward (TMA boxes with element stride), CUDA-core dgrad / wgrad
    (2, 32, 33, 33, 64, 3, 3, 1, 2, 1),
    (2, 64, 56, 56, 128, 3, 3, 1, 2, 1),    # ResNet stage transition
    (3, 32, 17, 20, 32, 3, 3, 0, 2, 1),
    (2, 32, 40, 40, 32, 3, 3, 1, 3, 1),     # stride 3
    (2, 64, 14, 14, 64, 1, 1, 0, 2, 1),     # 1x1 stride-2 projection
]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_family_vs_oracle(dev, case, mode):
    dev.set_math_mode(mode)
    B, C, H, W, O, kh, kw, pad, stride, dil = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((O, C, kh, kw)) * 0.1).astype(np.float32)
    dx, dw = dev.upload(x), dev.upload(w)
    y_ref = R.conv2d(x, w, pad, stride, dil)
    y = dev.conv2d(dx, dw, pad, stride, dil).numpy()
    assert rel_err(y, y_ref) <= TOL[mode], "fprop"
    gy = rng.standard_normal(y_ref.shape).astype(np.float32)
    dgy = dev.upload(gy)
    gx = dev.conv2d_transpose(dgy, dw, pad, stride, dil).numpy()
    assert rel_err(gx, R.conv2d_transpose(gy, w, pad, stride, dil)) <= TOL[mode], "dgrad"
    gw = dev.conv2d_filter_grad(dx, dgy, w.shape, pad, stride, dil).numpy()
    assert rel_err(gw, R.conv2d_filter_grad(x, gy, w.shape, pad, stride, dil)) <= TOL[mode], "wgrad"


FUSED_CASES = [  # B, C, H, W, O, kh, kw, pad, dil — stride 1 ("same"-style backward convs) + one SIMT-path shape
    (2, 64, 32, 32, 64, 3, 3, 1, 1), (1, 128, 32, 32, 256, 3, 3, 1, 1), (2, 48, 36, 40, 96, 3, 3, 1, 1), (2, 8, 12, 12, 16, 3, 3, 1, 1),
    (2, 64, 7, 128, 64, 3, 3, 1, 1), (1, 32, 4, 192, 80, 3, 3, 1, 1),      # wide maps: halo-reusing kernel
    (40, 64, 28, 32, 160, 3, 3, 1, 1),                                      # two M-tiles per CTA
    (80, 32, 16, 32, 96, 3, 3, 1, 1), (38, 64, 16, 64, 64, 3, 3, 1, 1),     # narrow maps: column-copies window kernel
    (5, 32, 7, 7, 64, 3, 3, 1, 1), (3, 64, 14, 14, 48, 3, 3, 1, 1),          # narrow maps on the per-tap kernel (partial tiles in lanes)
]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("cl", [False, True])
@pytest.mark.parametrize("case", FUSED_CASES)
def test_conv2d_fused_epilogues_vs_oracle(dev, case, cl, mode):
    """conv + bias + ReLU (fprop epilogue) and conv2d_transpose * (mask_src > 0) (dgrad epilogue), both memory orders"""
    dev.set_math_mode(mode)
    B, C, H, W, O, kh, kw, pad, dil = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((O, C, kh, kw)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(O).astype(np.float32)
    up = dev.upload_channels_last if cl else dev.upload
    dw = dev.upload(w)
    y_ref = np.maximum(R.conv2d(x, w, pad, 1, dil) + bias.reshape(1, O, 1, 1), 0)
    y = dev.conv2d(up(x), dw, pad, 1, dil, bias=dev.upload(bias), relu=True, channels_last=cl).numpy()
    assert rel_err(y, y_ref) <= TOL[mode], "fprop+bias+relu"
    gy = rng.standard_normal(y_ref.shape).astype(np.float32)
    mask_src = rng.standard_normal(x.shape).astype(np.float32)
    mask_src[0, 0, 0, :4] = 0.0                                    # exactly-zero entries are masked out (x > 0 is strict)
    gx_ref = R.conv2d_transpose(gy, w, pad, 1, dil) * (mask_src > 0)
    gx, cs = dev.conv2d_transpose(up(gy), dw, pad, 1, dil, mask_src=up(mask_src), channels_last=cl, chan_sum=True)
    gx = gx.numpy()
    assert rel_err(gx, gx_ref) <= TOL[mode], "dgrad*mask"
    assert np.all(gx[mask_src <= 0] == 0.0)
    assert rel_err(cs.numpy(), gx.astype(np.float64).sum(axis=(0, 2, 3))) <= 1e-5, "per-channel sums of the stored values (bias gradient)"
    # mask in the other memory order than the output: still correct (un-fused tail)
    gx2, cs2 = dev.conv2d_transpose(up(gy), dw, pad, 1, dil, mask_src=(dev.upload if cl else dev.upload_channels_last)(mask_src), channels_last=cl, chan_sum=True)
    assert rel_err(gx2.numpy(), gx_ref) <= TOL[mode]
    assert rel_err(cs2.numpy(), gx2.numpy().astype(np.float64).sum(axis=(0, 2, 3))) <= 1e-5


@pytest.mark.parametrize("cl", [False, True])
@pytest.mark.parametrize("case", [(2, 64, 28, 28, 32, 3, 3, 1, 2, 1), (3, 32, 15, 18, 96, 3, 3, 0, 2, 1), (2, 32, 20, 20, 32, 1, 1, 0, 2, 1), (2, 32, 26, 26, 160, 3, 3, 1, 3, 1)])
def test_strided_dgrad_phases_with_mask_and_channel_sums(dev, case, cl):
    """stride-s conv2d_transpose = s*s unit-stride phase convolutions on the tensor cores (TF32 mode), with the ReLU mask and the
    per-channel sums in the epilogue; a 1x1 stride-2 filter leaves three of four phases empty (zeros)."""
    dev.set_math_mode(1)
    B, C, H, W, O, kh, kw, pad, stride, dil = case
    rng = np.random.default_rng(sum(case))
    w = (rng.standard_normal((O, C, kh, kw)) * 0.1).astype(np.float32)
    yh, yw = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1, (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    gy = rng.standard_normal((B, O, yh, yw)).astype(np.float32)
    gx_plain = R.conv2d_transpose(gy, w, pad, stride, dil)
    mask_src = rng.standard_normal(gx_plain.shape).astype(np.float32)
    up = dev.upload_channels_last if cl else dev.upload
    gx, cs = dev.conv2d_transpose(up(gy), dev.upload(w), pad, stride, dil, mask_src=up(mask_src), channels_last=cl, chan_sum=True)
    gx = gx.numpy()
    assert gx.shape == gx_plain.shape
    assert rel_err(gx, gx_plain * (mask_src > 0)) <= TOL[1]
    assert rel_err(cs.numpy(), gx.astype(np.float64).sum(axis=(0, 2, 3))) <= 1e-5
    assert rel_err(dev.conv2d_transpose(up(gy), dev.upload(w), pad, stride, dil, channels_last=cl).numpy(), gx_plain) <= TOL[1]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", [(2, 1, 28, 28, 32, 3, 3, 1, 1, 1), (2, 3, 32, 32, 64, 3, 3, 1, 1, 1), (3, 3, 17, 13, 48, 3, 3, 1, 2, 1), (2, 2, 11, 11, 24, 3, 3, 2, 1, 2),
                                  (1, 4, 9, 9, 8, 2, 2, 0, 1, 1), (2, 3, 20, 20, 200, 3, 3, 1, 1, 1), (5, 1, 6, 6, 10, 5, 5, 2, 1, 1),
                                  # geometries of the tcgen05 first-layer kernels (W % 4 == 0, output width >= 32): partial second 128-pixel strip, no padding, 128 channels,
                                  # stride 2, dilation 2, 5x5 taps with 12 channels
                                  (3, 3, 40, 160, 64, 3, 3, 1, 1, 1), (2, 3, 33, 136, 32, 3, 3, 0, 1, 1), (2, 4, 20, 64, 128, 2, 2, 0, 1, 1), (4, 3, 24, 132, 48, 3, 3, 1, 2, 1),
                                  (2, 3, 24, 72, 96, 3, 3, 2, 1, 2), (3, 1, 36, 36, 12, 5, 5, 2, 1, 1), (9, 3, 64, 128, 64, 3, 3, 1, 1, 1)])
def test_small_channel_conv_channels_last(dev, case, mode):
    """first-layer kernels (C <= 4): NCHW input, channels-last output / output-gradient, ragged pixel groups, O not a multiple of 64"""
    dev.set_math_mode(mode)
    B, C, H, W, O, kh, kw, pad, stride, dil = case
    rng = np.random.default_rng(sum(case) + 1)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((O, C, kh, kw)) * 0.3).astype(np.float32)
    bias = rng.standard_normal(O).astype(np.float32)
    y_ref = np.maximum(R.conv2d(x, w, pad, stride, dil) + bias.reshape(1, O, 1, 1), 0)
    y = dev.conv2d(dev.upload(x), dev.upload(w), pad, stride, dil, bias=dev.upload(bias), relu=True, channels_last=True).numpy()
    assert rel_err(y, y_ref) <= TOL[mode]
    gy = rng.standard_normal(y_ref.shape).astype(np.float32)
    gw = dev.conv2d_filter_grad(dev.upload(x), dev.upload_channels_last(gy), w.shape, pad, stride, dil).numpy()
    assert rel_err(gw, R.conv2d_filter_grad(x, gy, w.shape, pad, stride, dil)) <= TOL[mode]


@pytest.mark.parametrize("case", [(2, 64, 8, 128, 64, 1), (1, 32, 7, 256, 96, 1), (2, 32, 6, 130, 48, 0)])
def test_conv_relu_pool_fused_matches_separate_kernels(dev, case):
    """conv+bias+ReLU+max_pool2d(2,0,2) in one epilogue must be bit-identical (values AND argmax) to the conv kernel followed by the
    pooling kernel, and within the TF32 tolerance of the oracle; odd heights / widths drop the last row / column like the reference."""
    dev.set_math_mode(1)
    B, C, H, W, O, pad = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((O, C, 3, 3)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(O).astype(np.float32)
    dx, dw, db = dev.upload_channels_last(x), dev.upload(w), dev.upload(bias)
    fused = dev.conv2d_pool(dx, dw, pad, 1, 1, bias=db, relu=True)
    assert fused is not None, "the wide-map kernel should take this layer"
    y = dev.conv2d(dx, dw, pad, 1, 1, bias=db, relu=True, channels_last=True)
    py, pidx = dev.max_pool2d(y, 2, 0, 2, int32_index=True)
    assert np.array_equal(fused[0].numpy(), py.numpy())
    assert np.array_equal(fused[1].numpy().view(np.int32), pidx.numpy().view(np.int32))
    y_ref = np.maximum(R.conv2d(x, w, pad, 1, 1) + bias.reshape(1, O, 1, 1), 0)
    assert rel_err(fused[0].numpy(), R.max_pool2d(y_ref, 2, 0, 2)[0]) <= TOL[1]
    dev.set_math_mode(0)
    assert dev.conv2d_pool(dx, dw, pad, 1, 1, bias=db, relu=True) is None      # 3xTF32 mode: not fused, the caller falls back


def test_conv_errors(dev):
    import rust_autograd_b200 as agb
    with pytest.raises(agb.OpError) as e:
        dev.conv2d(dev.upload(np.zeros((1, 3, 5, 5), np.float32)), dev.upload(np.zeros((2, 4, 3, 3), np.float32)))
    assert e.value.kind == "IncompatibleShape"


# ------------------------------------------------------------------------------------------------ pooling (bit-exact)
def test_max_pool_kat(dev):
    k = KATS["max_pool"]
    y, idx = dev.max_pool2d(dev.upload(np.array(k["x"], np.float32).reshape(1, 1, k["h"], k["w"])), k["size"], k["pad"], k["stride"])
    assert y.numpy().ravel().tolist() == k["output"] and idx.numpy().ravel().tolist() == k["argmax"]


@pytest.mark.parametrize("shape,size,stride", [((200, 32, 28, 28), 2, 2), ((4, 3, 7, 9), 3, 2), ((2, 5, 8, 8), 2, 1), ((3, 2, 9, 9), 3, 3)])
def test_max_pool_family_bit_exact(dev, shape, size, stride):
    rng = np.random.default_rng(1)
    x = rng.integers(-3, 4, shape).astype(np.float32)          # many ties -> exercises "first maximum wins"
    x[0, 0, :size, :size] = -np.inf                               # all -inf window keeps max_i = 0 (max_pool2d.rs:51-52)
    y_ref, idx_ref, _ = R.max_pool2d(x, size, 0, stride)
    y, idx = dev.max_pool2d(dev.upload(x), size, 0, stride)
    assert np.array_equal(y.numpy(), y_ref) and np.array_equal(idx.numpy(), idx_ref)
    gy = rng.standard_normal(y_ref.shape).astype(np.float32)
    gx = dev.max_pool2d_grad(dev.upload(gy), idx, size, 0, stride).numpy()
    close(gx, R.max_pool2d_grad(gy, idx_ref, size, 0, stride))
    gshape = gx.shape
    ggx = rng.standard_normal(gshape).astype(np.float32)
    ggy = dev.max_pool2d_grad_grad(dev.upload(ggx), idx, size, 0, stride).numpy() if gshape == x.shape else None
    if ggy is not None:
        assert np.array_equal(ggy, R.max_pool2d_grad_grad(ggx, idx_ref, size, 0, stride))


@pytest.mark.parametrize("cl", [False, True])
@pytest.mark.parametrize("shape,size,stride", [((6, 32, 28, 28), 2, 2), ((3, 8, 9, 9), 3, 3), ((2, 6, 8, 10), 2, 2), ((2, 5, 8, 8), 2, 1), ((2, 4, 7, 9), 3, 2)])
def test_max_pool_grad_fused_bit_exact(dev, shape, size, stride, cl):
    """gather-form backward for windows that tile the input, int32 indices, ReLU gate = (pooled output > 0); both memory orders"""
    rng = np.random.default_rng(2)
    x = np.maximum(rng.integers(-3, 4, shape), 0).astype(np.float32)          # a ReLU output: zeros and ties everywhere
    y_ref, idx_ref, _ = R.max_pool2d(x, size, 0, stride)
    up = dev.upload_channels_last if cl else dev.upload
    dx = up(x)
    y, idx = dev.max_pool2d(dx, size, 0, stride, int32_index=True)
    assert np.array_equal(y.numpy(), y_ref) and np.array_equal(idx.numpy().view(np.int32), idx_ref.astype(np.int32))
    gy = rng.standard_normal(y_ref.shape).astype(np.float32)
    gx_ref = R.max_pool2d_grad(gy, idx_ref, size, 0, stride)
    gx = dev.max_pool2d_grad(up(gy), idx, size, 0, stride, int32_index=True).numpy()
    close(gx, gx_ref)
    if gx_ref.shape == x.shape:
        gated, cs = dev.max_pool2d_grad(up(gy), idx, size, 0, stride, gate=y, int32_index=True, chan_sum=True)
        gated = gated.numpy()
        close(gated, gx_ref * (x > 0))                                       # == relu_grad(x, max_pool2d_grad(gy))
        assert rel_err(cs.numpy(), gated.astype(np.float64).sum(axis=(0, 2, 3))) <= 1e-5
    _, idx_f = dev.max_pool2d(dx, size, 0, stride)
    close(dev.max_pool2d_grad(up(gy), idx_f, size, 0, stride).numpy(), gx_ref)
    close(dev.max_pool2d_grad(up(gy), idx_f, size, 0, stride, window_known=False).numpy(), gx_ref)


# ------------------------------------------------------------------------------------------------ elementwise
UNARY = {"abs": (-3, 3), "neg": (-3, 3), "square": (-3, 3), "inv": (0.5, 3), "invsqrt": (0.5, 3), "sign": (-3, 3), "floor": (-3, 3),
         "ceil": (-3, 3), "sqrt": (0.1, 9), "ln": (0.1, 9), "log2": (0.1, 9), "log10": (0.1, 9), "exp": (-3, 3), "exp2": (-3, 3),
         "exp10": (-2, 2), "sin": (-3, 3), "cos": (-3, 3), "tan": (-1, 1), "asin": (-0.9, 0.9), "acos": (-0.9, 0.9), "atan": (-3, 3),
         "sinh": (-3, 3), "cosh": (-3, 3), "tanh": (-3, 3), "asinh": (-3, 3), "acosh": (1.1, 5), "atanh": (-0.9, 0.9),
         "sigmoid": (-6, 6), "relu": (-3, 3), "softplus": (-6, 6), "lgamma": (0.1, 30), "digamma": (0.1, 30)}


@pytest.mark.parametrize("op", sorted(UNARY))
def test_unary_vs_oracle(dev, op):
    lo, hi = UNARY[op]
    rng = np.random.default_rng(len(op))
    x = rng.uniform(lo, hi, (13, 1031)).astype(np.float32)
    x[0, :3] = [0.0 if lo < 0 < hi else lo, lo, hi]
    y = dev.unary(op, dev.upload(x)).numpy()
    close(y, R.unary(op, x), 1e-5)
    yt = dev.unary(op, dev.upload(x).transpose()).numpy()        # strided input view
    close(yt, R.unary(op, x.T), 1e-5)


def test_gamma_functions_negative_arguments(dev):
    """Lgamma / Digamma (math_ops.rs:1021-1060) left of zero, away from the poles: ln|Gamma(x)| and the reflected digamma."""
    rng = np.random.default_rng(12)
    x = (-rng.integers(0, 5, 4000) - rng.uniform(0.15, 0.85, 4000)).astype(np.float32)
    d = dev.upload(x)
    close(dev.unary("lgamma", d).numpy(), R.unary("lgamma", x), 1e-4)
    close(dev.unary("digamma", d).numpy(), R.unary("digamma", x), 1e-5)


def test_unary_param_ops(dev):
    rng = np.random.default_rng(7)
    x = rng.uniform(0.2, 3, (1000,)).astype(np.float32)
    d = dev.upload(x)
    close(dev.unary("pow", d, 2.5).numpy(), R.unary("pow", x, 2.5))
    close(dev.unary("elu", dev.upload(x - 1.5), 0.7).numpy(), R.unary("elu", x - 1.5, 0.7))
    assert np.array_equal(dev.unary("clip", d, 1.0, 2.0).numpy(), R.unary("clip", x, 1.0, 2.0))
    k = KATS["clip"]
    assert dev.unary("clip", dev.upload(np.array(k["x"], np.float32)), k["min"], k["max"]).numpy().tolist() == k["expected"]
    k = KATS["sign"]
    assert dev.unary("sign", dev.upload(np.array(k["x"], np.float32))).numpy().tolist() == k["expected"]
    gy = rng.standard_normal(1000).astype(np.float32)
    close(dev.binary("elu_grad", dev.upload(x - 1.5), dev.upload(gy), 0.7).numpy(), R.elu_grad(x - 1.5, gy, 0.7))
    assert np.array_equal(dev.binary("clip_grad", d, dev.upload(gy), 1.0, 2.0).numpy(), R.clip_grad(x, gy, 1.0, 2.0))


BIN = {"add": "add", "sub": "sub", "mul": "mul", "div": "div"}
CMP = {"eq": "equal", "ne": "not_equal", "gt": "greater", "lt": "lesser", "ge": "greater_equal", "le": "lesser_equal", "max": "maximum", "min": "minimum"}


@pytest.mark.parametrize("sa,sb", [((200, 32, 28, 28), (1, 32, 28, 28)), ((128, 64), (1, 64)), ((7, 5, 3), (7, 1, 3)), ((1000,), (1000,)),
                                   ((4, 1, 6), (1, 5, 6)), ((33, 17), (33, 1)), ((3, 4), ())])
def test_binary_broadcast_vs_oracle(dev, sa, sb):
    rng = np.random.default_rng(len(sa) * 10 + len(sb))
    a = rng.integers(-4, 5, sa).astype(np.float32) + rng.integers(0, 2, sa).astype(np.float32) * 0.5
    b = rng.integers(1, 5, sb).astype(np.float32)
    da, db = dev.upload(a), dev.upload(b)
    for op, ref in BIN.items():
        close(dev.binary(op, da, db).numpy(), R.binary_arith(ref, a, b), 1e-6)
        if op != "div":
            close(dev.binary(op, db, da).numpy(), R.binary_arith(ref, b, a), 1e-6)
    for op, ref in CMP.items():
        assert np.array_equal(dev.binary(op, da, db).numpy(), R.compare(ref, a, b)), op


@pytest.mark.parametrize("k", KATS["compare"])
def test_compare_kats(dev, k):
    op = {v: kk for kk, v in CMP.items()}[k["op"]]
    assert dev.binary(op, dev.upload(np.array(k["a"], np.float32)), dev.upload(np.array(k["b"], np.float32))).numpy().tolist() == k["expected"]


def test_add_n_fill_copy_dropout(dev):
    rng = np.random.default_rng(9)
    xs = [rng.standard_normal((17, 33)).astype(np.float32) for _ in range(11)]
    assert rel_err(dev.add_n([dev.upload(x) for x in xs]).numpy(), R.add_n(xs)) <= 1e-6
    k = KATS["add_n"]
    assert np.array_equal(dev.add_n([dev.fill(k["shape"], 1.0) for _ in range(k["n"])]).numpy(), np.array(k["expected"], np.float32))
    assert np.array_equal(dev.fill((5, 7), 2.5).numpy(), np.full((5, 7), 2.5, np.float32))
    assert np.array_equal(dev.fill((5, 7), 0.0).numpy(), np.zeros((5, 7), np.float32))
    x = rng.standard_normal((6, 10, 14)).astype(np.float32)
    d = dev.upload(x)
    assert np.array_equal(dev.copy(d.transpose((2, 0, 1))).numpy(), x.transpose(2, 0, 1))
    assert np.array_equal(dev.copy(d.slice(2, 3, 11)).numpy(), x[:, :, 3:11])
    mask = (rng.uniform(size=x.shape) < 0.75).astype(np.float32)
    y, _ = dev.dropout(d, 0.25, mask=dev.upload(mask))
    assert np.array_equal(y.numpy(), R.dropout(x, mask, 0.25))
    big = dev.upload(np.ones((1 << 20,), np.float32))
    y, m = dev.dropout(big, 0.25, seed=1234)
    m = m.numpy()
    assert set(np.unique(m)) == {0.0, 1.0} and abs(m.mean() - 0.75) < 5e-3 and np.array_equal(y.numpy(), m)


def test_concat_rows(dev):
    """agb_concat_rows: 70 blocks (two launches of the 64-entry pointer table), contiguous and sliced (pitch != cols) sources."""
    rng = np.random.default_rng(5)
    big = rng.standard_normal((24, 96)).astype(np.float32)
    dbig = dev.upload(big)
    xs = [rng.standard_normal((24, 32)).astype(np.float32) for _ in range(69)]
    ds = [dev.upload(x) for x in xs] + [dbig.slice(1, 32, 64)]
    got = dev.concat_rows(ds).numpy()
    assert np.array_equal(got, np.concatenate(xs + [big[:, 32:64]], axis=0))


@pytest.mark.parametrize("B,D", [(37, 48), (301, 128)])      # one element per thread / four elements per thread (>= 2^15 elements, ragged tail)
def test_fused_ewise_lstm_cell_matches_single_op_kernels(dev, B, D):
    """agb_fused_ewise on the LSTM cell (examples/lstm_lm.rs:36-45): gates are sliced views of one [B, 4D] buffer (read in place, pitch 4D),
    the bias is a row broadcast, outputs i, f, o, g, c', tanh(c'), h from ONE launch: bit-identical to the chain of agb_unary / agb_binary
    launches it replaces, and within 1e-5 of the oracle."""
    from rust_autograd_b200 import ffi
    rng = np.random.default_rng(21)
    xh = rng.standard_normal((B, 4 * D)).astype(np.float32)
    bias = rng.standard_normal((1, 4 * D)).astype(np.float32)
    c0 = rng.standard_normal((B, D)).astype(np.float32)
    dxh, db, dc = dev.upload(xh), dev.upload(bias), dev.upload(c0)
    # registers: 0..3 gate slices of xh, 4..7 bias slices, 8 c
    leaves = [(dxh.slice(1, k * D, (k + 1) * D), k) for k in range(4)] + [(db.slice(1, k * D, (k + 1) * D), 4 + k) for k in range(4)] + [(dc, 8)]
    U, Bn = ffi.F_UNARY, ffi.F_BINARY
    prog = [(Bn, "add", 9, 0, 4, 0.0), (Bn, "add", 10, 1, 5, 0.0), (Bn, "add", 11, 2, 6, 0.0), (Bn, "add", 12, 3, 7, 0.0),
            (U, "sigmoid", 13, 9, 0, 0.0), (U, "sigmoid", 14, 10, 0, 0.0), (U, "tanh", 15, 11, 0, 0.0), (U, "sigmoid", 16, 12, 0, 0.0),
            (Bn, "mul", 17, 14, 8, 0.0), (Bn, "mul", 18, 13, 15, 0.0), (Bn, "add", 19, 17, 18, 0.0),
            (U, "tanh", 20, 19, 0, 0.0), (Bn, "mul", 21, 16, 20, 0.0),
            (ffi.F_BINARY_IMM_A, "sub", 22, 0, 20, 1.0), (ffi.F_BINARY_IMM_B, "gt", 23, 21, 0, 0.0), (U, "scale", 24, 21, 0, 0.5)]
    outs = dev.fused_ewise(B, D, leaves, prog, [13, 14, 16, 15, 19, 20, 21, 22])
    pre = dev.binary("add", dxh, db)
    i_, f_, g_, o_ = (dev.copy(pre.slice(1, k * D, (k + 1) * D)) for k in range(4))
    i, f, g, o = dev.unary("sigmoid", i_), dev.unary("sigmoid", f_), dev.unary("tanh", g_), dev.unary("sigmoid", o_)
    c1 = dev.binary("add", dev.binary("mul", f, dc), dev.binary("mul", i, g))
    tc = dev.unary("tanh", c1)
    h = dev.binary("mul", o, tc)
    one_minus = dev.unary("rsub_scalar", tc, 1.0)
    for got, want in zip(outs, [i, f, o, g, c1, tc, h, one_minus]):
        assert np.array_equal(got.numpy(), want.numpy())
    sig = lambda v: R.unary("sigmoid", v)
    p = (xh.astype(np.float64) + bias).astype(np.float32)
    c_ref = sig(p[:, D:2 * D]).astype(np.float64) * c0 + sig(p[:, :D]).astype(np.float64) * np.tanh(p[:, 2 * D:3 * D].astype(np.float64))
    close(outs[4].numpy(), c_ref, 1e-5)            # the north_star tolerance for elementwise kernels (sums with cancellation near 0)
    close(outs[6].numpy(), sig(p[:, 3 * D:]).astype(np.float64) * np.tanh(c_ref), 1e-5)
    # immediates on either side, compare ops, column broadcast ([B,1] leaf), single row
    col = dev.upload(rng.standard_normal((B, 1)).astype(np.float32))
    y, = dev.fused_ewise(B, D, [(dc, 0), (col, 1)], [(Bn, "mul", 2, 0, 1, 0.0), (ffi.F_BINARY_IMM_B, "gt", 3, 2, 0, 0.0), (Bn, "mul", 4, 3, 0, 0.0)], [4])
    cc = col.numpy()
    assert np.array_equal(y.numpy(), ((c0 * cc) > 0).astype(np.float32) * c0)
    y, = dev.fused_ewise(1, B * D, [(dc.reshape((1, B * D)), 5)], [(U, "square", 6, 5, 0, 0.0)], [6])
    assert np.array_equal(y.numpy().reshape(B, D), c0 * c0)
    # malformed programs are rejected on the host, before any launch
    with pytest.raises(ffi.OpError):
        dev.fused_ewise(B, D, [(dc, 0)], [(Bn, "add", 1, 0, 7, 0.0)], [1])          # reads a register nothing wrote
    with pytest.raises(ffi.OpError):
        dev.fused_ewise(B, D, [(dc, 0)], [(U, "clip", 1, 0, 0, 0.0)], [1])          # two-parameter op


def test_random_generators_moments_and_reproducibility(dev):
    """agb_random (random_ops.rs:6-214, ndarray_ext.rs:276-388).  The reference's stream is parity-unpinned and its tests only check range
    and a != b (tests/test_array_gen.rs:4-40): here range, first two moments on 2^20 samples (4 sigma of the estimator), reproducibility from
    (seed, offset), and independence of offsets / seeds."""
    n = 1 << 20
    cases = [("uniform", -2.0, 3.0, 0.5, 25.0 / 12), ("normal", 1.5, 2.0, 1.5, 4.0), ("bernoulli", 0.3, 0.0, 0.3, 0.21), ("exp", 2.0, 0.0, 0.5, 0.25),
             ("log_normal", 0.1, 0.5, float(np.exp(0.1 + 0.125)), float((np.exp(0.25) - 1) * np.exp(0.2 + 0.25))),
             ("gamma", 2.5, 1.5, 3.75, 5.625), ("gamma", 0.4, 2.0, 0.8, 1.6)]
    for kind, p0, p1, mean, var in cases:
        a = dev.random(kind, (n,), p0, p1, seed=7, offset=3).numpy().astype(np.float64)
        assert np.isfinite(a).all(), kind
        assert abs(a.mean() - mean) <= 4 * np.sqrt(var / n) + 1e-6, (kind, a.mean(), mean)
        kurt_bound = 0.05 if kind in ("log_normal", "gamma") else 0.02       # heavier tails: looser variance estimate
        assert abs(a.var() - var) <= kurt_bound * var, (kind, a.var(), var)
        assert np.array_equal(a, dev.random(kind, (n,), p0, p1, seed=7, offset=3).numpy()), kind          # reproducible
        b = dev.random(kind, (n,), p0, p1, seed=7, offset=4).numpy()
        c_ = dev.random(kind, (n,), p0, p1, seed=8, offset=3).numpy()
        assert (a != b).mean() > (0.3 if kind == "bernoulli" else 0.99) and (a != c_).mean() > (0.3 if kind == "bernoulli" else 0.99), kind
    u = dev.random("uniform", (n,), 0.0, 1.0, seed=1).numpy()
    assert u.min() >= 0.0 and u.max() < 1.0
    assert set(np.unique(dev.random("bernoulli", (4096,), 0.5, 0.0, seed=1).numpy())) == {0.0, 1.0}
    assert (dev.random("exp", (n,), 1.0, 0.0, seed=2).numpy() >= 0).all() and (dev.random("gamma", (n,), 0.5, 1.0, seed=2).numpy() > 0).all()
    # odd sizes / shapes, launch-geometry independence: a prefix of a longer draw equals the shorter draw
    s = dev.random("normal", (1001,), 0.0, 1.0, seed=5).numpy()
    assert np.array_equal(s, dev.random("normal", (4003,), 0.0, 1.0, seed=5).numpy()[:1001])
    from rust_autograd_b200 import ffi
    with pytest.raises(ffi.OpError):
        dev.random("uniform", (8,), 1.0, 1.0)          # Uniform::new(low, high) requires low < high
    with pytest.raises(ffi.OpError):
        dev.random("gamma", (8,), -1.0, 1.0)


# ------------------------------------------------------------------------------------------------ reductions
@pytest.mark.parametrize("k", KATS["reduce"])
def test_reduce_kats(dev, k):
    x = np.array(k["x"], np.float32)
    assert np.array_equal(dev.reduce(k["op"], dev.upload(x), k["axes"][0]).numpy(), np.array(k["expected"], np.float32))


@pytest.mark.parametrize("op", ["sum", "mean", "prod", "min", "max"])
@pytest.mark.parametrize("shape,axis", [((200, 1), 0), ((200, 32, 28, 28), 0), ((128, 4096), 1), ((128, 4096), 0), ((7, 13, 5), 1), ((1 << 20,), 0),
                                        ((3, 100000), 1), ((100000, 3), 0)])
def test_reduce_vs_oracle(dev, op, shape, axis):
    rng = np.random.default_rng(axis + len(shape))
    x = (rng.uniform(0.97, 1.03, shape) if op == "prod" else rng.standard_normal(shape)).astype(np.float32)
    if op == "prod" and shape[axis] > 5000:
        x = np.ones(shape, np.float32)
    y = dev.reduce(op, dev.upload(x), axis).numpy()
    ref = R.reduce(op, x, [axis])
    if op in ("min", "max"):
        assert np.array_equal(y, ref)
    else:
        assert rel_err(y, ref) <= 1e-5


@pytest.mark.parametrize("k", KATS["argmax"] + KATS["argmin"])
def test_arg_kats(dev, k):
    is_max = k in KATS["argmax"]
    x = np.array(k["x"], np.float32)
    got = dev.argreduce(is_max, dev.upload(x), k["axis"]).numpy()
    assert np.array_equal(got, np.array(k["expected"], np.float32))


@pytest.mark.parametrize("shape,axis", [((1000, 10), 1), ((10, 1000), 0), ((4, 7, 9), 1), ((3, 100000), 1)])
def test_argreduce_bit_exact_with_ties(dev, shape, axis):
    rng = np.random.default_rng(4)
    x = rng.integers(0, 5, shape).astype(np.float32)
    for is_max in (True, False):
        assert np.array_equal(dev.argreduce(is_max, dev.upload(x), axis).numpy(), R.arg_reduce(x, axis, False, is_max))


# ------------------------------------------------------------------------------------------------ softmax family
@pytest.mark.parametrize("shape,axis", [((200, 10), 1), ((128, 8192), 1), ((7, 13, 5), 1), ((64, 100), 0), ((4, 40000), 1),
                                         ((37, 256), 1), ((19, 132), 1), ((50, 512), 1), ((9, 1024), 1), ((11, 1020), 1), ((5, 1023), 1), ((6, 12288), 1), ((3, 16384), 1), ((2, 16388), 1), ((3, 20000), 1), ((2, 100001), 1), ((2, 131072), 1), ((1, 140000), 1), ((5, 2048), 1), ((7, 128), 1)])
def test_softmax_family_vs_oracle(dev, shape, axis):
    rng = np.random.default_rng(8)
    x = (rng.standard_normal(shape) * 4).astype(np.float32)
    d = dev.upload(x)
    close(dev.softmax_like("softmax", d, axis).numpy(), R.softmax(x, axis), 1e-5)
    close(dev.softmax_like("log_softmax", d, axis).numpy(), R.log_softmax(x, axis), 1e-5)
    close(dev.softmax_like("logsumexp", d, axis).numpy(), R.logsumexp(x, axis, True), 1e-5)


@pytest.mark.parametrize("b,c", [(200, 10), (128, 8192), (1, 3), (1000, 1000), (77, 256), (31, 1024), (5, 16384), (3, 50000), (9, 2048)])
def test_sparse_xent_vs_oracle(dev, b, c):
    rng = np.random.default_rng(b + c)
    x = (rng.standard_normal((b, c)) * 3).astype(np.float32)
    t = rng.integers(0, c, (b,)).astype(np.float32)
    loss_ref, logx_ref = R.sparse_softmax_cross_entropy(x, t)
    loss, log_x = dev.sparse_xent_fwd(dev.upload(x), dev.upload(t))
    assert loss.shape == (b, 1)
    close(loss.numpy(), loss_ref, 1e-5)
    close(log_x.numpy(), logx_ref, 1e-5)
    gy = rng.standard_normal((b, 1)).astype(np.float32)
    gx = dev.sparse_xent_bwd(log_x, dev.upload(t), dev.upload(gy)).numpy()
    close(gx, R.sparse_softmax_cross_entropy_grad(logx_ref, t, gy), 1e-5)
    onehot = np.eye(c, dtype=np.float32)[t.astype(int)]
    l2, lx2 = dev.softmax_xent_fwd(dev.upload(x), dev.upload(onehot))
    assert l2.shape == (b,)
    close(l2.numpy(), R.softmax_cross_entropy(x, onehot)[0], 1e-5)
    close(dev.binary("sigmoid_xent", dev.upload(x), dev.upload(onehot)).numpy(), R.sigmoid_cross_entropy(x, onehot), 1e-5)


def test_sparse_xent_bad_label_is_out_of_bounds(dev):
    import rust_autograd_b200 as agb
    x = np.zeros((4, 3), np.float32)
    t = np.array([0, 1, 7, 2], np.float32)        # "Wrong label value" panic in the reference (xent_ops.rs:104)
    dev.sparse_xent_fwd(dev.upload(x), dev.upload(t))
    with pytest.raises(agb.OpError) as e:
        dev.sync()
    assert e.value.kind == "OutOfBounds"


# ------------------------------------------------------------------------------------------------ gather
@pytest.mark.parametrize("pshape,ishape,axis", [((8192, 64), (128, 1), 0), ((5, 7, 3), (4,), 1), ((6, 4), (2, 3), -1)])
def test_gather_family_bit_exact(dev, pshape, ishape, axis):
    rng = np.random.default_rng(6)
    p = rng.standard_normal(pshape).astype(np.float32)
    ax = axis % len(pshape)
    idx = rng.integers(-pshape[ax], pshape[ax], ishape).astype(np.float32)
    idx.ravel()[:2] = idx.ravel()[0]          # duplicates accumulate in the grad
    out = dev.gather(dev.upload(p), dev.upload(idx), axis).numpy()
    ref = R.gather(p, idx, axis)
    assert np.array_equal(out, ref)
    gy = rng.integers(-3, 4, ref.shape).astype(np.float32)
    gx = dev.gather_grad(dev.upload(gy), dev.upload(idx), pshape, axis).numpy()
    assert np.array_equal(gx, R.gather_grad(idx, pshape, gy, axis))
    # two scatter-adds into one zero table == the sum of two GatherGrads (integer-valued data: exact in any order)
    acc = dev.fill(pshape, 0.0)
    dev.scatter_add(acc, dev.upload(gy), dev.upload(idx), axis)
    dev.scatter_add(acc, dev.upload(2 * gy), dev.upload(idx), axis)
    assert np.array_equal(acc.numpy(), 3 * R.gather_grad(idx, pshape, gy, axis))


# ------------------------------------------------------------------------------------------------ optimizers
def test_multi_tensor_optimizers_vs_oracle(dev):
    rng = np.random.default_rng(10)
    shapes = [(32, 1, 3, 3), (1, 32, 28, 28), (64, 32, 3, 3), (3136, 10), (1, 10), (7,)]
    ps = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    gs = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    ms = [rng.standard_normal(s).astype(np.float32) * 0.1 for s in shapes]
    vs = [np.abs(rng.standard_normal(s)).astype(np.float32) * 0.1 for s in shapes]
    ts = [np.array([float(i + 1)], np.float32) for i in range(len(shapes))]
    up = lambda l: [dev.upload(a) for a in l]
    dp, dg, dm, dv, dt = up(ps), up(gs), up(ms), up(vs), up(ts)
    dev.adam(dp, dg, dm, dv, dt)
    for i in range(len(shapes)):
        p2, m2, v2, t2 = R.adam_update(ps[i], gs[i], ms[i], vs[i], ts[i])
        close(dp[i].numpy(), p2, 1e-5), close(dm[i].numpy(), m2, 1e-5), close(dv[i].numpy(), v2, 1e-5)
        assert dt[i].numpy()[0] == t2[0]
    dp = up(ps)
    dev.adam(dp, dg, up(ms), up(vs), up(ts), grad_scale=0.25)      # 1/world after the NCCL sum
    close(dp[3].numpy(), R.adam_update(ps[3], gs[3] * np.float32(0.25), ms[3], vs[3], ts[3])[0], 1e-5)
    dp = up(ps)
    dev.sgd(dp, dg, 0.1)
    for i in range(len(shapes)):
        close(dp[i].numpy(), R.sgd_update(ps[i], gs[i], 0.1), 1e-6)
    dp, dv2 = up(ps), up(ms)
    dev.momentum(dp, dg, dv2, 0.01, 0.9)
    for i in range(len(shapes)):
        p2, v2 = R.momentum_sgd_update(ps[i], gs[i], ms[i], 0.01, 0.9)
        close(dp[i].numpy(), p2, 1e-6), close(dv2[i].numpy(), v2, 1e-6)
    dp, dh = up(ps), up(vs)
    dev.adagrad(dp, dg, dh, 0.01)
    for i in range(len(shapes)):
        p2, h2 = R.adagrad_update(ps[i], gs[i], vs[i], 0.01)
        close(dp[i].numpy(), p2, 1e-5), close(dh[i].numpy(), h2, 1e-6)


# ------------------------------------------------------------------------------------------------ full-size properties
def test_reduce_sum_full_size_checksum(dev):
    """2^28 f32 (1 GiB) reduce_sum of an exactly representable pattern: the sum is known in closed form."""
    n = 1 << 28
    x = dev.fill((n,), 0.5)
    assert float(dev.reduce("sum", x, 0).numpy()) == n * 0.5
    x2 = x.reshape((1 << 14, 1 << 14))
    assert np.array_equal(dev.reduce("sum", x2, 1).numpy(), np.full((1 << 14,), (1 << 14) * 0.5, np.float32))
    sm = dev.softmax_like("softmax", x2, 1)
    assert float(dev.reduce("sum", sm.reshape((n,)), 0).numpy()) == float(1 << 14)


# ------------------------------------------------------------------------------------------------ deterministic reductions
DET_CASES = [  # name, B, C, H, W, O, math mode
    ("wgrad_taps_64_64", 40, 64, 32, 64, 64, 1), ("wgrad_tile_128_256", 16, 128, 32, 32, 256, 1), ("wgrad_tile_256_256_3x", 8, 256, 16, 32, 256, 0),
    ("wgrad_pair_64_128", 24, 64, 32, 32, 128, 1), ("first_layer_tc", 9, 3, 64, 128, 64, 1), ("first_layer_mma_3x", 6, 3, 40, 44, 64, 0),
]


@pytest.mark.parametrize("case", DET_CASES, ids=[c[0] for c in DET_CASES])
def test_filter_gradient_is_bit_reproducible(dev, case):
    """The reference's filter gradient is a sequential sum (conv2d.rs:631-734): run to run it gives the same bits.  With deterministic
    reductions on (the default) so do the split-K kernels here; the atomic form agrees with it to fp32 reassociation."""
    _, B, C, H, W, O, mode = case
    dev.set_math_mode(mode)
    rng = np.random.default_rng(B + C + O)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    gy = rng.standard_normal((B, O, H, W)).astype(np.float32)
    up = dev.upload if C <= 4 else dev.upload_channels_last
    dx, dgy = up(x), dev.upload_channels_last(gy)
    runs = [dev.conv2d_filter_grad(dx, dgy, (O, C, 3, 3), 1, 1, 1).numpy() for _ in range(3)]
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    dev.set_deterministic(False)
    try:
        atomic = dev.conv2d_filter_grad(dx, dgy, (O, C, 3, 3), 1, 1, 1).numpy()
    finally:
        dev.set_deterministic(True)
    assert rel_err(atomic, runs[0]) <= 2e-6
    assert rel_err(runs[0], R.conv2d_filter_grad(x, gy, (O, C, 3, 3), 1, 1, 1)) <= TOL[mode]


@pytest.mark.parametrize("case", [(2, 64, 8, 128, 64), (40, 64, 32, 32, 128), (38, 128, 16, 64, 64), (16, 256, 32, 32, 256), (3, 64, 14, 14, 48)],
                         ids=["rows", "cols128", "cols64", "pair256", "per_tap"])
def test_bias_gradient_side_sums_are_bit_reproducible(dev, case):
    """per-channel sums of the masked dgrad epilogue (= the bias gradient of the layer below): same bits every run, and the atomic form
    agrees to reassociation"""
    dev.set_math_mode(1)
    B, C, H, W, O = case
    rng = np.random.default_rng(sum(case))
    gy = rng.standard_normal((B, O, H, W)).astype(np.float32)
    w = (rng.standard_normal((O, C, 3, 3)) * 0.1).astype(np.float32)
    mask_src = rng.standard_normal((B, C, H, W)).astype(np.float32)
    dgy, dw, dm = dev.upload_channels_last(gy), dev.upload(w), dev.upload_channels_last(mask_src)
    runs = []
    for _ in range(3):
        gx, cs = dev.conv2d_transpose(dgy, dw, 1, 1, 1, mask_src=dm, channels_last=True, chan_sum=True)
        runs.append((gx.numpy(), cs.numpy()))
    assert all(np.array_equal(runs[0][0], r[0]) and np.array_equal(runs[0][1], r[1]) for r in runs[1:])
    assert rel_err(runs[0][1], runs[0][0].astype(np.float64).sum(axis=(0, 2, 3))) <= 1e-5
    dev.set_deterministic(False)
    try:
        _, cs_atomic = dev.conv2d_transpose(dgy, dw, 1, 1, 1, mask_src=dm, channels_last=True, chan_sum=True)
        assert rel_err(cs_atomic.numpy(), runs[0][1]) <= 1e-5
    finally:
        dev.set_deterministic(True)


def test_split_k_gemm_and_pool_sums_are_bit_reproducible(dev):
    dev.set_math_mode(1)
    rng = np.random.default_rng(3)
    a = rng.standard_normal((64, 8192)).astype(np.float32)           # [64, 32] output, long K: split-K with per-split copies of C (<= 1 MB of partials)
    b = rng.standard_normal((8192, 32)).astype(np.float32)
    da, db = dev.upload(a), dev.upload(b)
    runs = [dev.gemm(da, db).numpy() for _ in range(3)]
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    assert rel_err(runs[0], R.matmul(a, b)) <= TOL[1]
    at = rng.standard_normal((4096, 64)).astype(np.float32)           # A^T B with a long K: the weight-gradient shape
    bt = rng.standard_normal((4096, 32)).astype(np.float32)
    dat, dbt = dev.upload(at), dev.upload(bt)
    runs = [dev.gemm(dat, dbt, trans_a=True).numpy() for _ in range(3)]
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    assert rel_err(runs[0], R.matmul(at, bt, True, False)) <= TOL[1]
    # gated pool backward with per-channel sums
    x = rng.standard_normal((6, 64, 32, 32)).astype(np.float32)
    y, idx = dev.max_pool2d(dev.upload_channels_last(np.maximum(x, 0)), 2, 0, 2, int32_index=True)
    gy = dev.upload_channels_last(rng.standard_normal((6, 64, 16, 16)).astype(np.float32))
    runs = []
    for _ in range(3):
        gx, cs = dev.max_pool2d_grad(gy, idx, 2, 0, 2, gate=y, int32_index=True, chan_sum=True)
        runs.append((gx.numpy(), cs.numpy()))
    assert all(np.array_equal(runs[0][1], r[1]) for r in runs[1:])
    assert rel_err(runs[0][1], runs[0][0].astype(np.float64).sum(axis=(0, 2, 3))) <= 1e-5


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", [("fwd", 64, 8192, 10), ("fwd", 70, 5000, 3), ("fwd", 33, 16384, 16), ("wgrad", 300, 8192, 10), ("wgrad", 64, 5001, 7),
                                  ("dgrad", 70, 8192, 10), ("dgrad", 256, 4100, 16)], ids=lambda c: "%s-%d-%d-%d" % c)
def test_skinny_gemm_vs_oracle(dev, case, mode):
    """the classifier GEMMs of a CNN (one extent <= 16 next to a large operand): dedicated streaming kernels in exact fp32 FMA, in every math mode"""
    dev.set_math_mode(mode)
    kind, b, feat, cls = case
    rng = np.random.default_rng(b + feat + cls)
    x = rng.standard_normal((b, feat)).astype(np.float32)
    w = (rng.standard_normal((feat, cls)) * 0.05).astype(np.float32)
    g = rng.standard_normal((b, cls)).astype(np.float32)
    if kind == "fwd":
        got, ref = dev.gemm(dev.upload(x), dev.upload(w)).numpy(), R.matmul(x, w)
        prev = rng.standard_normal((b, cls)).astype(np.float32)
        acc = dev.upload(prev)
        dev.gemm(dev.upload(x), dev.upload(w), out=acc, beta=1.0)
        assert rel_err(acc.numpy(), prev + ref) <= 1e-5
    elif kind == "wgrad":
        got, ref = dev.gemm(dev.upload(x), dev.upload(g), trans_a=True).numpy(), R.matmul(x, g, True, False)
    else:
        got, ref = dev.gemm(dev.upload(g), dev.upload(w), trans_b=True).numpy(), R.matmul(g, w, False, True)
    assert got.shape == ref.shape
    assert rel_err(got, ref) <= 1e-5


@pytest.mark.parametrize("case", [(4, 3, 64, 128, 64), (16, 256, 32, 32, 256), (3, 64, 14, 14, 128), (40, 128, 32, 32, 256)],
                         ids=["first_layer->rows", "pair->pair", "per_tap->per_tap", "pair->pair(128->256)"])
def test_relu_sign_bits_side_channel(dev, case):
    """The fused conv + bias + ReLU kernels leave the SIGN BITS of their output next to it (1 bit per element, channels-last order), and the masked
    dgrad of the next layer reads them instead of the activation (activation_ops.rs:161-166 needs only x > 0): the bits must be exactly (y > 0)
    and the gradient exactly the one computed from the activation itself."""
    dev.set_math_mode(1)
    B, C, H, W, O = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((O, C, 3, 3)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(O).astype(np.float32)
    dx = dev.upload(x) if C <= 4 else dev.upload_channels_last(x)
    y, bits = dev.conv2d_relu_bits(dx, dev.upload(w), 1, 1, 1, bias=dev.upload(bias))
    assert bits is not None, "this layer runs on a kernel that writes the bits"
    yv = y.numpy()                                                  # logical [B, O, H, W]
    words = bits.numpy().view(np.uint32)
    expect = np.packbits((yv.transpose(0, 2, 3, 1) > 0).reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).ravel()
    assert np.array_equal(words[:expect.size], expect)
    # the next layer's masked dgrad: O -> O2 channels, gradient w.r.t. y
    O2 = O
    gy = rng.standard_normal((B, O2, H, W)).astype(np.float32)
    w2 = (rng.standard_normal((O2, O, 3, 3)) * 0.1).astype(np.float32)
    dgy, dw2 = dev.upload_channels_last(gy), dev.upload(w2)
    gx_a, cs_a = dev.conv2d_transpose(dgy, dw2, 1, 1, 1, mask_src=y, channels_last=True, chan_sum=True)
    gx_b, cs_b = dev.conv2d_transpose(dgy, dw2, 1, 1, 1, mask_src=y, channels_last=True, chan_sum=True, mask_bits=bits)
    assert np.array_equal(gx_a.numpy(), gx_b.numpy()) and np.array_equal(cs_a.numpy(), cs_b.numpy())
    assert rel_err(gx_b.numpy(), R.conv2d_transpose(gy, w2, 1, 1, 1) * (yv > 0)) <= TOL[1]
