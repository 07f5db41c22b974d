"""ORACLE (test infrastructure, NOT product code) — numpy restatement of rust-autograd's `Op::compute` kernels.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this.

The reference (`/root/reference`, crate `autograd` 2.0.0-rc3) is Rust and cannot be compiled in this image (no
cargo/rustc; its arithmetic lives in un-vendored crates: ndarray 0.16.1, matrixmultiply 0.3.2, rand 0.8, Cargo.toml:15-35).
Each function below restates one reference routine and cites the `file:line` it follows.  Floating-point work is done
in float64 and rounded to float32 once (the reference accumulates in f32 in an order fixed by ndarray/matrixmultiply
internals that are not part of the repository; parity is asserted at 1e-5 relative, BASELINE.json north_star).
Integer/index-valued outputs (argmax, pool indices, gather) are exact.

Pinning status: GEMM, im2col, max-pool (+indices), argmax (+ties), reductions, compare/select, clip,
tile, add_n are pinned by the reference's own known-answer tests (tests/test_oracle_golden.py holds the vectors and
cites them).  conv2d / conv2d_transpose / filter-grad values beyond the im2col KAT, softmax/xent values and the
optimizers have NO golden values in the reference (its tests are finite-difference self-checks or assert nothing):
for those this file is "parity unpinned" — the restatement itself is the oracle, cross-checked by finite differences.
"""
import numpy as np

F32 = np.float32
F32_MIN = np.finfo(np.float32).min     # T::min_value() for floats = most negative finite
F32_MAX = np.finfo(np.float32).max


OUT_DTYPE = np.float32        # the reference tests run their finite-difference checks with F = f64: tests may switch this


def set_out_dtype(dt):
    """Results are rounded once to this dtype (float32 = the product's element type; float64 for FD self-checks)."""
    global OUT_DTYPE
    OUT_DTYPE = dt


def _f32(a):
    return np.asarray(a, dtype=np.float64).astype(OUT_DTYPE)


def _f64(a):
    return np.asarray(a, dtype=np.float64)


# ----------------------------------------------------------------------------------------------------------------
# dense contractions — src/tensor_ops/dot_ops.rs
# ----------------------------------------------------------------------------------------------------------------
class OpError(Exception):
    """src/op.rs:67-73"""

    def __init__(self, kind, msg):
        self.kind = kind
        super().__init__("%s: %s" % (kind, msg))


def matmul(a, b, transpose_a=False, transpose_b=False):
    """MatMul::compute, dot_ops.rs:565-606: 2-D only; transposes are stride swaps (:574-579)."""
    a, b = _f64(a), _f64(b)
    if a.ndim != 2 or b.ndim != 2:
        raise OpError("IncompatibleShape", "matmul: lhs/rhs input's ndim must be 2")   # dot_ops.rs:568-573
    if transpose_a:
        a = a.T
    if transpose_b:
        b = b.T
    if a.shape[1] != b.shape[0]:
        raise OpError("IncompatibleShape", "shapes %s and %s not aligned" % (a.shape, b.shape))  # :580-584
    return _f32(a @ b)


def batch_matmul(a, b, transpose_a=False, transpose_b=False):
    """BatchMatMul::compute, dot_ops.rs:632-695: rank >= 2, identical leading dims (no broadcast :661),
    transposes on the last two axes (:640-647)."""
    a, b = _f64(a), _f64(b)
    if a.ndim < 2 or b.ndim < 2:
        raise OpError("IncompatibleShape", "BatchMatMul: Left-hand-side input's ndim must be >= 2")
    if transpose_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b:
        b = np.swapaxes(b, -1, -2)
    if a.ndim != b.ndim or a.shape[:-2] != b.shape[:-2]:
        raise OpError("IncompatibleShape", "Input shapes mismatch: %s vs %s" % (a.shape, b.shape))   # :661-666
    if a.shape[-1] != b.shape[-2]:
        raise OpError("IncompatibleShape", "Input shapes mismatch: %s vs %s" % (a.shape, b.shape))
    return _f32(np.matmul(a, b))


# ----------------------------------------------------------------------------------------------------------------
# convolution family — src/tensor_ops/conv_ops/{mod,conv2d,conv2d_transpose}.rs
# ----------------------------------------------------------------------------------------------------------------
def conv_out_size(x, k, pad, stride, dil):
    """conv2d.rs:132-133 / conv_ops/mod.rs:88-89"""
    return (x + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


def im2col(x, kh, kw, pad, stride, dil, dtype=np.float64):
    """im2col / im2col_batch, conv_ops/mod.rs:73-176.  x [B,C,H,W] -> cols [B,C,kh,kw,yh,yw]
    (loop order c, kh, kw, yh, yw at :93-121; out-of-image taps are zero :104-114).
    Quirk mod.rs:98 (x start uses `ph`) is unobservable: the public API only passes square padding."""
    x = np.asarray(x, dtype=dtype)
    B, C, H, W = x.shape
    yh, yw = conv_out_size(H, kh, pad, stride, dil), conv_out_size(W, kw, pad, stride, dil)
    xp = np.zeros((B, C, H + 2 * pad + stride, W + 2 * pad + stride), dtype=dtype)
    xp[:, :, pad:pad + H, pad:pad + W] = x
    cols = np.empty((B, C, kh, kw, yh, yw), dtype=dtype)
    for i in range(kh):
        for j in range(kw):
            ys, xs = i * dil, j * dil
            cols[:, :, i, j] = xp[:, :, ys:ys + stride * yh:stride, xs:xs + stride * yw:stride][:, :, :yh, :yw]
    return cols


def col2im(cols, H, W, pad, stride, dil):
    """col2im, conv_ops/mod.rs:178-223: scatter-add of cols [B,C,kh,kw,yh,yw] into [B,C,H,W]."""
    cols = _f64(cols)
    B, C, kh, kw, yh, yw = cols.shape
    xp = np.zeros((B, C, H + 2 * pad + stride * yh, W + 2 * pad + stride * yw), dtype=np.float64)
    for i in range(kh):
        for j in range(kw):
            ys, xs = i * dil, j * dil
            xp[:, :, ys:ys + stride * yh:stride, xs:xs + stride * yw:stride][:, :, :yh, :yw] += cols[:, :, i, j]
    return xp[:, :, pad:pad + H, pad:pad + W]


def conv2d(x, w, pad=0, stride=1, dil=1, return_cols=False):
    """Conv2D::compute, conv2d.rs:532-554 -> conv2d_impl :407-487 -> slow_im2col_gemm_fused_kernel :115-211:
    y[b] = W[O, C*kh*kw] . im2col(x[b])[C*kh*kw, yh*yw].  Outputs (y [B,O,yh,yw], cols [B,C,kh,kw,yh,yw])."""
    x, w = np.asarray(x), np.asarray(w)
    if x.ndim != 4:
        raise OpError("IncompatibleShape", "conv2d: lhs input must be 4D (got %s)" % (x.shape,))     # conv2d.rs:346-404
    if w.ndim != 4:
        raise OpError("IncompatibleShape", "conv2d: filter must be 4D (got %s)" % (w.shape,))
    if x.shape[1] != w.shape[1]:
        raise OpError("IncompatibleShape", "conv2d: input channel dim (%d) must match filter's second dim (%d)" % (x.shape[1], w.shape[1]))
    O, C, kh, kw = w.shape
    cols = im2col(x, kh, kw, pad, stride, dil)
    B, _, _, _, yh, yw = cols.shape
    y = np.einsum("ok,bkp->bop", _f64(w).reshape(O, C * kh * kw), cols.reshape(B, C * kh * kw, yh * yw), optimize=True)
    y = _f32(y.reshape(B, O, yh, yw))
    return (y, _f32(cols)) if return_cols else y


def conv2d_transpose(gy, w, pad=0, stride=1, dil=1):
    """Conv2DTranspose::compute, conv2d_transpose.rs:250-272 -> conv2d_transpose_impl :89-247:
    cols[b] = W^T[C*kh*kw, O] . gy[b][O, yh*yw]; gx = col2im(cols).  w is [O(=gy channels), C, kh, kw].
    Output size follows the CODE (:55-56): xh = s(yh-1) - 2p + (d(kh-1)+1) (the doc comment differs)."""
    gy, w = np.asarray(gy), np.asarray(w)
    if gy.ndim != 4:
        raise OpError("IncompatibleShape", "conv2d_transpose: Input must be 4D (got %s)" % (gy.shape,))
    if w.ndim != 4:
        raise OpError("IncompatibleShape", "conv2d_transpose: Filter must be 4D (got %s)" % (w.shape,))
    if gy.shape[1] != w.shape[0]:
        raise OpError("IncompatibleShape", "conv2d_transpose: Number of input channels (%d) must match second filter dim (%d)" % (gy.shape[1], w.shape[0]))
    B, O, yh, yw = gy.shape
    _, C, kh, kw = w.shape
    xh = stride * (yh - 1) - 2 * pad + (dil * (kh - 1) + 1)
    xw = stride * (yw - 1) - 2 * pad + (dil * (kw - 1) + 1)
    cols = np.einsum("ok,bop->bkp", _f64(w).reshape(O, C * kh * kw), _f64(gy).reshape(B, O, yh * yw), optimize=True)
    return _f32(col2im(cols.reshape(B, C, kh, kw, yh, yw), xh, xw, pad, stride, dil))


def conv2d_filter_grad(x, gy, w_shape, pad=0, stride=1, dil=1):
    """Conv2DFilterGrad::compute, conv2d.rs:737-744 -> conv2d_filter_grad_impl :631-734:
    gw[O, C*kh*kw] = sum_b gy[b][O, yh*yw] . cols[b]^T (beta=1 accumulation over the batch :703-722).
    The reference consumes the materialised cols; here they are recomputed from x (same values)."""
    O, C, kh, kw = w_shape
    cols = im2col(x, kh, kw, pad, stride, dil)
    B, _, _, _, yh, yw = cols.shape
    gy = _f64(gy)
    if gy.shape != (B, O, yh, yw):
        raise OpError("IncompatibleShape", "conv2d_filter_grad: gy shape %s != %s" % (gy.shape, (B, O, yh, yw)))
    gw = np.einsum("bop,bkp->ok", gy.reshape(B, O, yh * yw), cols.reshape(B, C * kh * kw, yh * yw), optimize=True)
    return _f32(gw.reshape(O, C, kh, kw))


def conv2d_transpose_filter_grad(gy_img, x, w_shape, pad=0, stride=1, dil=1):
    """Conv2DTransposeFilterGrad::compute, conv2d_transpose.rs:433-451 -> :303-431: for y = conv2d_transpose(x, w),
    gw[xc, yc*kh*kw] = sum_b x[b][xc, xh*xw] . im2col(gy[b])^T : the roles of image and gradient are swapped.
    (Quirk :324 `yh*yh`: only square spatial sizes are correct in the reference; same arithmetic otherwise.)"""
    return conv2d_filter_grad(gy_img, x, w_shape, pad, stride, dil)


# ----------------------------------------------------------------------------------------------------------------
# pooling — src/tensor_ops/conv_ops/max_pool2d.rs
# ----------------------------------------------------------------------------------------------------------------
def max_pool2d(x, size, pad=0, stride=1):
    """MaxPool2D::compute :166-227 / impl_max_pool! :21-88.  Strict `>` scan starting from T::min_value(), first maximum
    in row-major window order wins (:62-69); index = flat offset into the WHOLE input buffer incl. batch and channel
    (:61 `index = w + xw*(h + xh*(c + b*ch))`), returned as float (:74-75).  Only pad == 0 is meaningful (usize wrap :43,53)."""
    x = np.asarray(x, dtype=OUT_DTYPE)
    assert pad == 0, "reference underflows usize for pad > 0 (max_pool2d.rs:43,53)"
    B, C, H, W = x.shape
    yh, yw = (H + 2 * pad - size) // stride + 1, (W + 2 * pad - size) // stride + 1
    best = np.full((B, C, yh, yw), F32_MIN, dtype=OUT_DTYPE)
    besti = np.zeros((B, C, yh, yw), dtype=np.int64)
    base = (np.arange(B * C, dtype=np.int64) * (H * W)).reshape(B, C, 1, 1)
    oy = (np.arange(yh) * stride).reshape(1, 1, yh, 1)
    ox = (np.arange(yw) * stride).reshape(1, 1, 1, yw)
    for dh in range(size):          # row-major window scan: h outer, w inner (:57-60)
        for dw in range(size):
            hh, ww = oy + dh, ox + dw
            valid = (hh < H) & (ww < W)           # h_end/w_end clipping (:44-48,54-58)
            hc, wc = np.minimum(hh, H - 1), np.minimum(ww, W - 1)
            v = x[:, :, hc.reshape(yh, 1), wc.reshape(1, yw)]
            take = valid & (v > best)             # NaN > x is False, like the reference
            best = np.where(take, v, best)
            besti = np.where(take, base + hc * W + wc, besti)
    return best, besti.astype(OUT_DTYPE), besti


def max_pool2d_grad(gy, idx, size, pad=0, stride=1):
    """MaxPool2DGrad::compute :245-279 / impl_max_pool_grad! :111-135: gx = zeros [B,C,xh,xw], xh = s(yh-1)-2p+size (:263-264);
    gx[idx[i]] += gy[i] sequentially (duplicates accumulate)."""
    gy = _f64(gy)
    B, C, yh, yw = gy.shape
    xh, xw = stride * (yh - 1) - 2 * pad + size, stride * (yw - 1) - 2 * pad + size
    gx = np.zeros(B * C * xh * xw, dtype=np.float64)
    np.add.at(gx, np.asarray(idx).astype(np.int64).ravel(), gy.ravel())
    return _f32(gx.reshape(B, C, xh, xw))


def max_pool2d_grad_grad(ggx, idx, size, pad=0, stride=1):
    """MaxPool2DGradGrad::compute :297-331 / impl_max_pool_grad_grad! :137-159: ggy[i] = ggx[idx[i]]."""
    ggx = np.asarray(ggx, dtype=OUT_DTYPE)
    return ggx.ravel()[np.asarray(idx).astype(np.int64)].reshape(np.asarray(idx).shape)


# ----------------------------------------------------------------------------------------------------------------
# elementwise — binary_ops.rs, math_ops.rs, activation_ops.rs, array_ops.rs
# ----------------------------------------------------------------------------------------------------------------
def _is_scalar_shape(shape):
    """ndarray_ext.rs:120-122: rank 0 or [0]"""
    return len(shape) == 0 or tuple(shape) == (0,)


def binary_arith(op, a, b):
    """AddOp/SubOp/MulOp/DivOp::compute, binary_ops.rs:147-290 + macro :304-347: scalar fast paths (rank-0 or shape [0];
    Div by scalar = multiply by reciprocal :251-255), otherwise ndarray broadcasting arithmetic (equal rank)."""
    a, b = np.asarray(a, dtype=OUT_DTYPE), np.asarray(b, dtype=OUT_DTYPE)
    a64, b64 = _f64(a), _f64(b)
    if op == "add":
        r = a64 + b64
    elif op == "sub":
        r = a64 - b64
    elif op == "mul":
        r = a64 * b64
    elif op == "div":
        if b.size == 1 and (_is_scalar_shape(b.shape) or b.shape == (1,)):     # binary_ops.rs:245-255
            r = a64 * _f64(F32(1.0) / b.reshape(()))
        else:
            r = a64 / b64
    else:
        raise ValueError(op)
    return _f32(r)


def compare(op, a, b):
    """impl_cmp_op!, math_ops.rs:86-184: 0/1-valued floats; Maximum/Minimum select."""
    a, b = np.asarray(a, dtype=OUT_DTYPE), np.asarray(b, dtype=OUT_DTYPE)
    if op == "equal":
        return (a == b).astype(OUT_DTYPE)
    if op == "not_equal":
        return (a != b).astype(OUT_DTYPE)
    if op == "greater":
        return (a > b).astype(OUT_DTYPE)
    if op == "lesser":
        return (a < b).astype(OUT_DTYPE)
    if op == "greater_equal":
        return (a >= b).astype(OUT_DTYPE)
    if op == "lesser_equal":
        return (a <= b).astype(OUT_DTYPE)
    if op == "maximum":
        return np.where(a > b, a, b).astype(OUT_DTYPE)      # math_ops.rs:160-167 `if a > b {a} else {b}`
    if op == "minimum":
        return np.where(a < b, a, b).astype(OUT_DTYPE)
    raise ValueError(op)


def unary(op, x, p0=0.0, p1=0.0):
    """math_ops.rs:277-1019 (x.map(f)), activation_ops.rs:113-226, array_ops.rs:537-574 (Clip)."""
    x32 = np.asarray(x, dtype=OUT_DTYPE)
    x = _f64(x32)
    with np.errstate(all="ignore"):
        if op == "abs":
            r = np.abs(x)
        elif op == "neg":
            r = -x
        elif op == "square":
            r = x * x
        elif op == "inv":
            r = 1.0 / x
        elif op == "invsqrt":
            r = 1.0 / np.sqrt(x)
        elif op == "sign":                           # math_ops.rs:370-381: 0 -> 0 else signum
            r = np.where(x == 0, 0.0, np.sign(x))
        elif op == "floor":
            r = np.floor(x)
        elif op == "ceil":
            r = np.ceil(x)
        elif op == "sqrt":
            r = np.sqrt(x)
        elif op == "pow":
            r = np.power(x, p0)
        elif op == "ln":
            r = np.log(x)
        elif op == "log2":
            r = np.log2(x)
        elif op == "log10":
            r = np.log10(x)
        elif op == "exp":
            r = np.exp(x)
        elif op == "exp2":
            r = np.exp2(x)
        elif op == "exp10":
            r = np.power(10.0, x)
        elif op in ("sin", "cos", "tan", "sinh", "cosh", "tanh"):
            r = getattr(np, op)(x)
        elif op in ("asin", "acos", "atan", "asinh", "acosh", "atanh"):
            r = getattr(np, "arc" + op[1:])(x)
        elif op == "sigmoid":                        # activation_ops.rs:138-141
            r = np.tanh(x * 0.5) * 0.5 + 0.5
        elif op == "relu":                           # activation_ops.rs:156  x.max(0): NaN -> 0
            r = np.where(np.isnan(x), 0.0, np.maximum(x, 0.0))
        elif op == "softplus":                       # activation_ops.rs:115 (unguarded; evaluated in f32 range)
            r = np.log(np.exp(x32.astype(np.float32) if OUT_DTYPE == np.float32 else x32).astype(np.float64) + 1.0)
        elif op == "elu":                            # activation_ops.rs:188-198
            r = np.where(x > 0, x, p0 * (np.exp(x) - 1.0))
        elif op == "clip":                           # array_ops.rs:540-545  a.min(max).max(min)
            r = np.maximum(np.minimum(x, p1), p0)
        elif op == "scale":
            r = x * _f64(F32(p0))
        elif op in ("lgamma", "digamma"):            # math_ops.rs:1021-1060: `special` 0.10 Gamma::{ln_gamma().0, digamma} (crate absent here);
            from scipy import special as _sp         # restated with the same functions from scipy.special (ln|Gamma(x)|, psi(x))
            r = _sp.gammaln(x) if op == "lgamma" else _sp.digamma(x)
        else:
            raise ValueError(op)
    return _f32(r)


def elu_grad(x, gy, alpha):
    """ELUGrad::compute, activation_ops.rs:204-226"""
    x, gy = _f64(x), _f64(gy)
    return _f32(np.where(x > 0, 1.0, alpha * (np.exp(x) - 1.0) + alpha) * gy)


def clip_grad(x, gy, lo, hi):
    """ClipGrad::compute, array_ops.rs:556-574"""
    x = np.asarray(x, dtype=OUT_DTYPE)
    return _f32(((x > F32(lo)) & (x < F32(hi))).astype(np.float64) * _f64(gy))


def add_n(xs):
    """AddN::compute, array_ops.rs:503-528: left fold"""
    acc = _f64(xs[0]).copy()
    for x in xs[1:]:
        acc = acc + _f64(x)
    return _f32(acc)


def dropout(x, mask, ratio, train=True):
    """Dropout::compute, random_ops.rs:218-237: train: y = x*mask, NOT rescaled; eval: y = x*(1-ratio)."""
    if train:
        return _f32(_f64(x) * _f64(mask))
    return _f32(_f64(x) * _f64(F32(1.0) - F32(ratio)))


# ----------------------------------------------------------------------------------------------------------------
# reductions — src/tensor_ops/reduction_ops.rs
# ----------------------------------------------------------------------------------------------------------------
def _norm_axes(axes, ndim):
    return sorted(set(int(a) + ndim if int(a) < 0 else int(a) for a in np.asarray(axes).ravel()))


def reduce(op, x, axes, keep_dims=False):
    """impl_reduce_forward!, reduction_ops.rs:54-108 (sorted axes folded highest first; empty axes / rank-0 -> view of x :63-70);
    ReduceMean :187-215: sum then multiply by 1/len with len accumulated as f32 (:198-209)."""
    x32 = np.asarray(x, dtype=OUT_DTYPE)
    if x32.ndim == 0 or np.asarray(axes).size == 0:
        return x32
    ax = tuple(_norm_axes(axes, x32.ndim))
    x = _f64(x32)
    if op == "sum":
        r = x.sum(axis=ax, keepdims=keep_dims)
    elif op == "mean":
        ln = F32(1.0)
        for a in ax:
            ln = F32(ln * F32(x32.shape[a]))
        r = x.sum(axis=ax, keepdims=keep_dims) * _f64(F32(1.0) / ln)
    elif op == "prod":
        r = x.prod(axis=ax, keepdims=keep_dims)
    elif op == "min":   # fold from T::max_value() with Float::min (NaN ignored)  :288-330
        r = np.minimum(np.min(np.where(np.isnan(x), np.inf, x), axis=ax, keepdims=keep_dims), F32_MAX)
    elif op == "max":
        r = np.maximum(np.max(np.where(np.isnan(x), -np.inf, x), axis=ax, keepdims=keep_dims), F32_MIN)
    else:
        raise ValueError(op)
    return _f32(r)


def sum_all(x):
    """ReduceSumToScalar::compute, reduction_ops.rs:123-128: 0-d result."""
    return _f32(_f64(x).sum())


def arg_reduce(x, axis, keep_dim=False, is_max=True):
    """ArgMax/ArgMin via argx_helper, reduction_ops.rs:365-429: FIRST occurrence of the extreme along `axis`, as float."""
    x = np.asarray(x, dtype=OUT_DTYPE)
    axis = axis + x.ndim if axis < 0 else axis
    r = np.argmax(x, axis=axis) if is_max else np.argmin(x, axis=axis)     # numpy returns the first occurrence
    r = r.astype(OUT_DTYPE)
    return np.expand_dims(r, axis) if keep_dim else r


def broadcast_to(x, shape):
    """ReduceGradCommon / MaybeBroadcast, reduction_ops.rs:459-496, binary_ops.rs:108-137"""
    return np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=OUT_DTYPE), shape))


def reduce_grad_common(gy, x_shape, axes, keep_dims=False):
    """ReduceGradCommon::compute, reduction_ops.rs:459-496: re-insert reduced axes (unless keep_dims) and broadcast to x_shape."""
    gy = np.asarray(gy, dtype=OUT_DTYPE)
    if len(x_shape) == 0 or tuple(gy.shape) == tuple(x_shape):
        return gy.reshape(x_shape)
    if not keep_dims:
        for a in _norm_axes(axes, len(x_shape)):
            gy = np.expand_dims(gy, a)
    return broadcast_to(gy, x_shape)


def maybe_reduce_sum(gy, target_shape):
    """MaybeReduceSum::compute, binary_ops.rs:39-94: identity when shapes match; scalar target -> full sum reshaped;
    else sum over every axis where target == 1 < gy (keeping the axis)."""
    gy32 = np.asarray(gy, dtype=OUT_DTYPE)
    target_shape = tuple(int(s) for s in target_shape)
    if tuple(gy32.shape) == target_shape:
        return gy32
    if _is_scalar_shape(target_shape):
        return _f32(_f64(gy32).sum()).reshape(())
    r = _f64(gy32)
    for i, (g, t) in enumerate(zip(gy32.shape, target_shape)):
        if t == 1 and g > 1:
            r = r.sum(axis=i, keepdims=True)
    return _f32(r)


# ----------------------------------------------------------------------------------------------------------------
# softmax family — activation_ops.rs:61-96, math_ops.rs:540-593, xent_ops.rs
# ----------------------------------------------------------------------------------------------------------------
def logsumexp(x, axis, keep_dims=True):
    """logsumexp_forward, math_ops.rs:540-593: max (fold from T::min_value()) -> exp(x-max) -> sum -> ln -> + max"""
    x = _f64(np.asarray(x, dtype=OUT_DTYPE))
    m = np.maximum(x.max(axis=axis, keepdims=True), F32_MIN)
    r = np.log(np.exp(x - m).sum(axis=axis, keepdims=True)) + m
    return _f32(r if keep_dims else np.squeeze(r, axis))


def softmax(x, axis):
    """softmax_impl, activation_ops.rs:61-96"""
    x = _f64(np.asarray(x, dtype=OUT_DTYPE))
    m = np.maximum(x.max(axis=axis, keepdims=True), F32_MIN)
    e = np.exp(x - m)
    return _f32(e / e.sum(axis=axis, keepdims=True))


def log_softmax(x, axis):
    """LogSoftmax::compute, xent_ops.rs:17-22: x - logsumexp(x, axis, keep)"""
    x = _f64(np.asarray(x, dtype=OUT_DTYPE))
    m = np.maximum(x.max(axis=axis, keepdims=True), F32_MIN)
    return _f32(x - (np.log(np.exp(x - m).sum(axis=axis, keepdims=True)) + m))


def sparse_softmax_cross_entropy(x, t):
    """SparseSoftmaxCrossEntropy::compute, xent_ops.rs:63-113: axis 1, 2-D logits, labels [B] or [B,1] as floats;
    outputs (loss [B,1], log_x [B,C])."""
    x = np.asarray(x, dtype=OUT_DTYPE)
    t = np.asarray(t)
    if x.ndim != 2:
        raise OpError("IncompatibleShape", "SparseSoftmaxCrossEntropy: given first argument's ndim is not 2: shape=%s" % (x.shape,))
    if not (t.ndim == 1 or (t.ndim == 2 and t.shape[1] == 1)):
        raise OpError("IncompatibleShape", "SparseSoftmaxCrossEntropy: second argument's shape must be (batch_size, 1) or (batch_size,). given shape=%s" % (t.shape,))
    x64 = _f64(x)
    m = np.maximum(x64.max(axis=1, keepdims=True), F32_MIN)
    log_x = x64 - (np.log(np.exp(x64 - m).sum(axis=1, keepdims=True)) + m)
    idx = t.astype(np.int64).ravel()
    loss = -log_x[np.arange(x.shape[0]), idx].reshape(x.shape[0], 1)
    return _f32(loss), _f32(log_x)


def sparse_softmax_cross_entropy_grad(log_x, t, gy):
    """SparseSoftmaxCrossEntropyGrad::compute, xent_ops.rs:139-152: (exp(log_x) - onehot(t)) * gy"""
    x = np.exp(_f64(log_x))
    idx = np.asarray(t).astype(np.int64).ravel()
    x[np.arange(x.shape[0]), idx] -= 1.0
    return _f32(x * _f64(gy))


def softmax_cross_entropy(x, t):
    """SoftmaxCrossEntropy::compute, xent_ops.rs:160-177: outputs (loss (B,), log_x (B,C))"""
    x64 = _f64(np.asarray(x, dtype=OUT_DTYPE))
    m = np.maximum(x64.max(axis=1, keepdims=True), F32_MIN)
    log_x = x64 - (np.log(np.exp(x64 - m).sum(axis=1, keepdims=True)) + m)
    return _f32(-(_f64(t) * log_x).sum(axis=1)), _f32(log_x)


def sigmoid_cross_entropy(x, t):
    """SigmoidCrossEntropy::compute, xent_ops.rs:33-46"""
    x, t = _f64(np.asarray(x, dtype=OUT_DTYPE)), _f64(t)
    return _f32(np.log(np.exp(-np.abs(x)) + 1.0) + np.maximum(0.0, x) - t * x)


# ----------------------------------------------------------------------------------------------------------------
# gather / scatter — array_ops.rs:353-474
# ----------------------------------------------------------------------------------------------------------------
def gather(param, indices, axis):
    """Gather::compute, array_ops.rs:353-384: out shape = param[..axis] + indices.shape + param[axis+1..]; negative ids wrap."""
    param = np.asarray(param, dtype=OUT_DTYPE)
    idx = np.asarray(indices).astype(np.int64)
    axis = axis + param.ndim if axis < 0 else axis
    idx = np.where(idx < 0, idx + param.shape[axis], idx)
    return np.take(param, idx, axis=axis)


def gather_grad(indices, param_shape, gy, axis):
    """GatherGrad::compute, array_ops.rs:401-466: gx = zeros(param); sequential row add (duplicates accumulate)."""
    idx = np.asarray(indices).astype(np.int64)
    axis = axis + len(param_shape) if axis < 0 else axis
    idx = np.where(idx < 0, idx + param_shape[axis], idx).ravel()
    pre = int(np.prod(param_shape[:axis], dtype=np.int64))
    post = int(np.prod(param_shape[axis + 1:], dtype=np.int64))
    gy = _f64(gy).reshape(pre, idx.size, post)
    gx = np.zeros((pre, param_shape[axis], post), dtype=np.float64)
    for p in range(pre):
        np.add.at(gx[p], idx, gy[p])
    return _f32(gx.reshape(param_shape))


# ----------------------------------------------------------------------------------------------------------------
# optimizers — src/tensor_ops/gradient_descent_ops/*.rs (in place on f32 arrays, evaluated in f32 like the reference)
# ----------------------------------------------------------------------------------------------------------------
def adam_update(p, g, m, v, t, alpha=1e-3, eps=1e-8, b1=0.9, b2=0.999):
    """AdamOp::compute, gradient_descent_ops/adam.rs:11-58.  t is the per-variable counter starting at 1 (optimizers/adam.rs:97)."""
    p64, g64, m64, v64 = _f64(p), _f64(g), _f64(m), _f64(v)
    b1_, b2_, al, ep = _f64(F32(b1)), _f64(F32(b2)), _f64(F32(alpha)), _f64(F32(eps))
    m_new = m64 * b1_ + _f64(F32(1.0) - F32(b1)) * g64
    v_new = v64 * b2_ + _f64(F32(1.0) - F32(b2)) * g64 * g64
    tv = float(np.asarray(t, dtype=np.float32).reshape(-1)[0])
    rv = _f64(F32(1.0) / (F32(1.0) - np.power(F32(b2), F32(tv), dtype=np.float32)))
    rm = _f64(F32(1.0) / (F32(1.0) - np.power(F32(b1), F32(tv), dtype=np.float32)))
    m_hat = (m_new * rm) / (np.sqrt(v_new * rv) + ep)
    return _f32(p64 - al * m_hat), _f32(m_new), _f32(v_new), _f32(np.asarray(t, dtype=np.float64) + 1.0)


def sgd_update(p, g, alpha):
    """SGDOp::compute, sgd.rs:14-26: p -= alpha * g  (via scaled_add(-alpha, g))"""
    return _f32(_f64(p) - _f64(F32(alpha)) * _f64(g))


def momentum_sgd_update(p, g, v, lr=0.01, momentum=0.9):
    """MomentumSGDOp::compute, sgd.rs:28-40: v = momentum*v - lr*g; p += v"""
    v_new = _f64(v) * _f64(F32(momentum)) - _f64(F32(lr)) * _f64(g)
    return _f32(_f64(p) + v_new), _f32(v_new)


def adagrad_update(p, g, h, lr):
    """AdaGradOp::compute, adagrad.rs:8-21: h += g*g; p -= lr * g / (sqrt(h) + 1e-7)"""
    h_new = _f64(h) + _f64(g) * _f64(g)
    return _f32(_f64(p) - _f64(F32(lr)) * _f64(g) / (np.sqrt(h_new) + _f64(F32(1e-7)))), _f32(h_new)
