"""Driver of oracle/cpu_ref.c: the f32 CPU restatement of the reference's default-build hot path, evaluated op by op the way the
reference's Evaluator walks the VGG training graph (workloads.vgg_loss + grad_helper + Adam), every intermediate a fresh array.

TEST / BASELINE INFRASTRUCTURE ONLY (see the header of cpu_ref.c): imported by tests/ and by bench.py's `--impl reference` /
`cpu_baseline` legs, never by the package.  `kind` is "port": the Rust crate itself cannot be built in this image (no cargo/rustc)."""
import ctypes as C
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_f, _i, _i64, _P = C.c_float, C.c_int, C.c_int64, C.c_void_p
_lib = None
SGEMM_BACKEND = None


def load():
    """dlopen oracle/lib/libcpuref.so (built by `make -C oracle`) and hand it a SINGLE-THREADED cblas_sgemm from the OpenBLAS that numpy
    ships, standing in for matrixmultiply::sgemm (the crate's default build; Cargo.toml:21 has no `threading` feature)."""
    global _lib, SGEMM_BACKEND
    if _lib is not None:
        return _lib
    path = os.path.join(HERE, "lib", "libcpuref.so")
    if not os.path.exists(path):
        raise RuntimeError("oracle/lib/libcpuref.so is missing: run `make -C oracle` (or __graft_entry__.build())")
    lib = C.CDLL(path)
    lib.cr_mean.restype = _f
    SGEMM_BACKEND = "built-in packed AVX2 6x16 kernel"
    cands = glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so"))
    if cands and os.environ.get("CPUREF_BUILTIN_SGEMM") != "1":
        try:
            ob = C.CDLL(cands[0])
            lib.cr_set_sgemm(C.cast(ob.scipy_cblas_sgemm64_, _P))
            lib._openblas = ob
            SGEMM_BACKEND = "OpenBLAS cblas_sgemm (scipy-openblas shipped with numpy), 1 thread per call"
        except (OSError, AttributeError):
            pass
    _lib = lib
    return lib


class single_threaded_blas:
    """One OpenBLAS thread per sgemm call while the port runs (parallelism only over samples, like rayon in the reference); the previous
    setting is restored on exit because numpy — and with it the f64 oracle of the parity tests — shares this OpenBLAS."""

    def __enter__(self):
        ob = getattr(load(), "_openblas", None)
        self.prev = ob.scipy_openblas_get_num_threads64_() if ob is not None else None
        if ob is not None:
            ob.scipy_openblas_set_num_threads64_(1)

    def __exit__(self, *exc):
        if self.prev is not None:
            load()._openblas.scipy_openblas_set_num_threads64_(self.prev)


def _p(a):
    return a.ctypes.data_as(_P)


def _new(*shape):
    return np.empty(shape, np.float32)


class Ops:
    """One method per reference op; every call allocates its outputs like Op::compute does."""

    def __init__(self):
        self.lib = load()

    def conv2d(self, x, w, pad):
        B, Cc, H, W_ = x.shape
        O, _, kh, kw = w.shape
        yh, yw = H + 2 * pad - kh + 1, W_ + 2 * pad - kw + 1
        y, cols = _new(B, O, yh, yw), _new(B, Cc * kh * kw, yh * yw)
        self.lib.cr_conv2d(_p(x), _p(w), _p(y), _p(cols), B, Cc, H, W_, O, kh, kw, pad, 1, 1)
        return y, cols

    def conv2d_transpose(self, gy, w, pad, H, W_):
        B, O = gy.shape[:2]
        _, Cc, kh, kw = w.shape
        gx = _new(B, Cc, H, W_)
        self.lib.cr_conv2d_transpose(_p(gy), _p(w), _p(gx), B, Cc, H, W_, O, kh, kw, pad, 1, 1)
        return gx

    def conv2d_filter_grad(self, cols, gy, wshape):
        O, Cc, kh, kw = wshape
        B, _, yh, yw = gy.shape
        gw = _new(*wshape)
        self.lib.cr_conv2d_filter_grad(_p(cols), _p(gy), _p(gw), B, Cc, O, kh, kw, yh, yw)
        return gw

    def max_pool2d(self, x, size, stride):
        B, Cc, H, W_ = x.shape
        yh, yw = (H - size) // stride + 1, (W_ - size) // stride + 1
        y, idx = _new(B, Cc, yh, yw), _new(B, Cc, yh, yw)
        self.lib.cr_max_pool2d(_p(x), _p(y), _p(idx), B, Cc, H, W_, size, stride)
        return y, idx

    def max_pool2d_grad(self, gy, idx, xshape):
        gx = _new(*xshape)
        self.lib.cr_max_pool2d_grad(_p(gy), _p(idx), _p(gx), _i64(gy.size), _i64(gx.size))
        return gx

    def add_bias_nchw(self, x, b):
        y = _new(*x.shape)
        self.lib.cr_add_bias_nchw(_p(x), _p(b), _p(y), x.shape[0], x.shape[1], _i64(x.shape[2] * x.shape[3]))
        return y

    def add_rowvec(self, x, b):
        y = _new(*x.shape)
        self.lib.cr_add_rowvec(_p(x), _p(b), _p(y), _i64(x.shape[0]), _i64(x.shape[1]))
        return y

    def relu(self, x):
        y = _new(*x.shape)
        self.lib.cr_relu(_p(x), _p(y), _i64(x.size))
        return y

    def relu_grad(self, x, gy):          # gy * greater(x, 0): two ops in the reference graph
        m, g = _new(*x.shape), _new(*x.shape)
        self.lib.cr_greater0(_p(x), _p(m), _i64(x.size))
        self.lib.cr_mul(_p(gy), _p(m), _p(g), _i64(x.size))
        return g

    def sum_to_channels(self, g):
        out = _new(1, g.shape[1], 1, 1)
        self.lib.cr_sum_to_channels(_p(g), _p(out), g.shape[0], g.shape[1], _i64(g.shape[2] * g.shape[3]))
        return out

    def sum_rows(self, g):
        out = _new(1, g.shape[1])
        self.lib.cr_sum_rows(_p(g), _p(out), _i64(g.shape[0]), _i64(g.shape[1]))
        return out

    def matmul(self, a, b, ta=False, tb=False):
        m = a.shape[1] if ta else a.shape[0]
        k = a.shape[0] if ta else a.shape[1]
        n = b.shape[0] if tb else b.shape[1]
        c = _new(m, n)
        self.lib.cr_matmul(_p(a), _p(b), _p(c), _i64(m), _i64(n), _i64(k), int(ta), int(tb))
        return c

    def sparse_xent(self, x, t):
        loss, log_x = _new(x.shape[0], 1), _new(*x.shape)
        self.lib.cr_sparse_xent(_p(x), _p(t), _p(loss), _p(log_x), _i64(x.shape[0]), _i64(x.shape[1]))
        return loss, log_x

    def sparse_xent_grad(self, log_x, t, gy):
        gx = _new(*log_x.shape)
        self.lib.cr_sparse_xent_grad(_p(log_x), _p(t), _p(gy), _p(gx), _i64(log_x.shape[0]), _i64(log_x.shape[1]))
        return gx

    def adam(self, p, g, m, v, t, alpha=1e-3, eps=1e-8, b1=0.9, b2=0.999):
        tmp = _new(p.size)
        self.lib.cr_adam(_p(p), _p(g), _p(m), _p(v), _p(t), _i64(p.size), _f(alpha), _f(eps), _f(b1), _f(b2), _p(tmp))


class VggTrainer:
    """The training step of workloads.vgg_loss (conv + bias + relu blocks, 2x2 max-pool, FC, sparse softmax xent, reduce_mean) with
    grad_helper's gradients and Adam on every variable — the op sequence the reference's Evaluator would run, in f32."""

    def __init__(self, params, layers, size):
        self.ops = Ops()
        self.layers, self.size = layers, size
        self.p = {k: np.ascontiguousarray(v, np.float32).copy() for k, v in params.items()}
        self.m = {k: np.zeros_like(v) for k, v in self.p.items()}
        self.v = {k: np.zeros_like(v) for k, v in self.p.items()}
        self.t = {k: np.ones(1, np.float32) for k in self.p}

    def step(self, x, y, update=True):
        with single_threaded_blas():
            return self._step(x, y, update)

    def _step(self, x, y, update):
        o, p = self.ops, self.p
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        tape, t, i = [], x, 0
        for l in self.layers:
            if l == "pool":
                pooled, idx = o.max_pool2d(t, 2, 2)
                tape.append(("pool", t.shape, idx))
                t = pooled
                continue
            z, cols = o.conv2d(t, p["conv%d_w" % i], 1)
            zb = o.add_bias_nchw(z, p["conv%d_b" % i])
            a = o.relu(zb)
            tape.append(("conv", i, t.shape, cols, zb))
            t, i = a, i + 1
        B = x.shape[0]
        flat = t.reshape(B, -1)
        logits = o.add_rowvec(o.matmul(flat, p["fc_w"]), p["fc_b"])
        loss_rows, log_x = o.sparse_xent(logits, y)
        loss = float(o.lib.cr_mean(_p(loss_rows), _i64(B)))
        # backward (the graph grad_helper builds: d mean -> xent grad -> FC -> reshape -> blocks in reverse)
        gy = np.full((B, 1), 1.0 / B, np.float32)
        g_logits = o.sparse_xent_grad(log_x, y, gy)
        grads = {"fc_b": o.sum_rows(g_logits), "fc_w": o.matmul(flat, g_logits, ta=True)}
        g = o.matmul(g_logits, p["fc_w"], tb=True).reshape(t.shape)
        for rec in reversed(tape):
            if rec[0] == "pool":
                g = o.max_pool2d_grad(g, rec[2], rec[1])
                continue
            _, i, xshape, cols, zb = rec
            gz = o.relu_grad(zb, g)
            grads["conv%d_b" % i] = o.sum_to_channels(gz)
            grads["conv%d_w" % i] = o.conv2d_filter_grad(cols, gz, p["conv%d_w" % i].shape)
            if i > 0:
                g = o.conv2d_transpose(gz, p["conv%d_w" % i], 1, xshape[2], xshape[3])
        if update:
            for k in p:
                o.adam(p[k], np.ascontiguousarray(grads[k].reshape(p[k].shape)), self.m[k], self.v[k], self.t[k])
        return loss, grads


def vgg_params(rng, size=128, layers=None, classes=10):
    """Same draws, in the same order, as workloads.vgg_init"""
    from rust_autograd_b200 import workloads as W

    class _Env:
        def __init__(self):
            self.vars, self._n = {}, None

        def default_namespace_mut(self):
            return self

        def slot(self):
            return self

        def name(self, n):
            self._n = n
            return self

        def set(self, v):
            self.vars[self._n] = np.asarray(v, np.float32)
    e = _Env()
    W.vgg_init(e, rng, size=size, classes=classes, layers=layers or W.VGG_LAYERS)
    return e.vars
