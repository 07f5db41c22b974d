"""Kernel-level Python handle on the C ABI: a ``Device`` (context + stream + HBM arena) and ``DArray``
(device-resident f32 array with shape/strides, the device analogue of ``NdArray``/``NdArrayView``).

Only plumbing lives here (allocation, H2D/D2H, view arithmetic, output-shape computation for the kernel
entry points); every number is produced by ``libagb200.so``.
"""
import ctypes as C

import numpy as np

from . import ffi


class DArray:
    """f32 array in HBM.  ``owner`` keeps the backing block alive for views."""

    def __init__(self, dev, ptr, shape, strides=None, owner=None, owns=False):
        self.dev, self.ptr, self.shape = dev, int(ptr) if ptr else 0, tuple(int(d) for d in shape)
        if strides is None:
            strides, s = [], 1
            for d in reversed(self.shape):
                strides.append(s)
                s *= d
            strides = tuple(reversed(strides))
        self.strides = tuple(int(s) for s in strides)
        self.owner, self.owns = owner, owns

    @property
    def size(self):
        n = 1
        for d in self.shape:
            n *= d
        return n

    @property
    def ndim(self):
        return len(self.shape)

    def desc(self):
        return ffi.make_tensor(self.ptr, self.shape, self.strides)

    def is_contiguous(self):
        s = 1
        for d, st in zip(reversed(self.shape), reversed(self.strides)):
            if d != 1 and st != s:
                return False
            s *= d
        return True

    # zero-copy views (reference: OpOutput::View, Transpose = stride permutation math_ops.rs:448)
    def reshape(self, shape):
        assert self.is_contiguous()
        return DArray(self.dev, self.ptr, shape, owner=self)

    def transpose(self, perm=None):
        perm = perm or tuple(reversed(range(self.ndim)))
        return DArray(self.dev, self.ptr, [self.shape[p] for p in perm], [self.strides[p] for p in perm], owner=self)

    def broadcast_to(self, shape):
        shape = tuple(shape)
        pad = len(shape) - self.ndim
        sh = (1,) * pad + self.shape
        st = (0,) * pad + self.strides
        strides = []
        for d, t, s in zip(sh, shape, st):
            if d == t:
                strides.append(s)
            elif d == 1:
                strides.append(0)
            else:
                raise ffi.OpError(ffi.ERR_INCOMPATIBLE_SHAPE, "cannot broadcast %s to %s" % (self.shape, shape))
        return DArray(self.dev, self.ptr, shape, strides, owner=self)

    def slice(self, axis, start, stop):
        shape = list(self.shape)
        shape[axis] = stop - start
        return DArray(self.dev, self.ptr + 4 * start * self.strides[axis], shape, self.strides, owner=self)

    def numpy(self):
        return self.dev.download(self)

    def free(self):
        if self.owns and self.ptr and self.dev.ctx:
            ffi.check(self.dev.lib.agb_free(self.dev.ctx, self.ptr))
            self.ptr, self.owns = 0, False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Device:
    def __init__(self, index=0):
        self.lib = ffi.load_library()
        ctx = C.c_void_p()
        ffi.check(self.lib.agb_init(index, C.byref(ctx)))
        self.ctx = ctx
        self.index = index

    def close(self):
        if self.ctx:
            ctx, self.ctx = self.ctx, None      # DArrays that outlive the context must not call agb_free on it
            self.lib.agb_destroy(ctx)

    # ---- memory ----
    def empty(self, shape):
        n = 1
        for d in shape:
            n *= int(d)
        p = C.c_void_p()
        ffi.check(self.lib.agb_alloc(self.ctx, max(n, 1) * 4, C.byref(p)))
        return DArray(self, p.value, shape, owns=True)

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        t = self.empty(a.shape)
        ffi.check(self.lib.agb_h2d(self.ctx, t.ptr, ffi.np_ptr(a), a.nbytes))
        self.sync()          # the numpy temporary may die right after this call
        return t

    def download(self, t):
        if not t.is_contiguous():
            c = self.empty(t.shape)
            ffi.check(self.lib.agb_copy_strided(self.ctx, t.desc(), c.desc()))
            t = c
        out = np.empty(t.shape, dtype=np.float32)
        ffi.check(self.lib.agb_d2h(self.ctx, ffi.np_ptr(out), t.ptr, out.nbytes))
        ffi.check(self.lib.agb_sync(self.ctx))
        return out

    def sync(self):
        ffi.check(self.lib.agb_sync(self.ctx))

    def set_math_mode(self, mode):
        ffi.check(self.lib.agb_set_math_mode(self.ctx, mode))

    def set_deterministic(self, on):
        ffi.check(self.lib.agb_set_deterministic(self.ctx, 1 if on else 0))

    def launch_count(self):
        v = C.c_int64()
        ffi.check(self.lib.agb_launch_count(self.ctx, C.byref(v)))
        return v.value

    def sm_count(self):
        v = C.c_int()
        ffi.check(self.lib.agb_sm_count(self.ctx, C.byref(v)))
        return v.value

    # ---- timing on the context's stream ----
    def event(self):
        e = C.c_void_p()
        ffi.check(self.lib.agb_event_create(C.byref(e)))
        return e

    def record(self, e):
        ffi.check(self.lib.agb_event_record(self.ctx, e))

    def elapsed_ms(self, a, b):
        ms = C.c_float()
        ffi.check(self.lib.agb_event_elapsed_ms(a, b, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        ffi.check(self.lib.agb_flush_l2(self.ctx))

    # ---- kernels (thin: compute the output shape, call C) ----
    def unary(self, op, x, p0=0.0, p1=0.0):
        y = self.empty(x.shape)
        ffi.check(self.lib.agb_unary(self.ctx, ffi.U[op], p0, p1, x.desc(), y.desc()))
        return y

    def binary(self, op, a, b, p0=0.0, p1=0.0):
        shape = np.broadcast_shapes(a.shape, b.shape)
        y = self.empty(shape)
        ffi.check(self.lib.agb_binary(self.ctx, ffi.B[op], p0, p1, a.broadcast_to(shape).desc(), b.broadcast_to(shape).desc(), y.desc()))
        return y

    def add_n(self, xs):
        y = self.empty(xs[0].shape)
        descs = [x.desc() for x in xs]
        arr = (C.POINTER(ffi.AgbTensor) * len(xs))(*[C.pointer(d) for d in descs])
        ffi.check(self.lib.agb_add_n(self.ctx, len(xs), arr, y.desc()))
        return y

    def fused_ewise(self, rows, cols, leaves, program, out_regs):
        """agb_fused_ewise: `leaves` = [(2-D DArray view broadcastable to [rows, cols], reg)], `program` = [(kind, op name, dst, a, b, p0)],
        `out_regs` = registers to store; returns one contiguous [rows, cols] array per output register."""
        lv = (ffi.AgbFuseLeaf * max(len(leaves), 1))()
        for i, (x, reg) in enumerate(leaves):
            assert x.ndim == 2
            pitch = 0 if x.shape[0] == 1 and rows != 1 else x.strides[0]
            cs = 0 if x.shape[1] == 1 and cols != 1 else x.strides[1]
            lv[i] = ffi.AgbFuseLeaf(x.ptr, pitch, cs, reg)
        ins = (ffi.AgbFuseInstr * len(program))()
        for i, (kind, op, dst, a, b, p0) in enumerate(program):
            ins[i] = ffi.AgbFuseInstr(kind, (ffi.U if kind == ffi.F_UNARY else ffi.B)[op], dst, a, b, p0)
        ys = [self.empty((rows, cols)) for _ in out_regs]
        outs = (ffi.AgbFuseOut * len(out_regs))()
        for i, (y, reg) in enumerate(zip(ys, out_regs)):
            outs[i] = ffi.AgbFuseOut(y.ptr, cols, reg)
        ffi.check(self.lib.agb_fused_ewise(self.ctx, rows, cols, len(leaves), lv, len(program), ins, len(out_regs), outs))
        return ys

    def concat_rows(self, xs):
        """agb_concat_rows: equally-shaped 2-D blocks (unit column stride) stacked along axis 0 in one launch."""
        rows, cols = xs[0].shape
        y = self.empty((len(xs) * rows, cols))
        ptrs = (C.c_void_p * len(xs))(*[x.ptr for x in xs])
        pitch = (C.c_int64 * len(xs))(*[x.strides[0] for x in xs])
        ffi.check(self.lib.agb_concat_rows(self.ctx, len(xs), ptrs, pitch, rows, cols, y.ptr))
        return y

    def fill(self, shape, v):
        y = self.empty(shape)
        ffi.check(self.lib.agb_fill(self.ctx, y.desc(), v))
        return y

    def copy(self, x):
        y = self.empty(x.shape)
        ffi.check(self.lib.agb_copy_strided(self.ctx, x.desc(), y.desc()))
        return y

    def _axis_view(self, shape, axis):
        outer = 1
        for d in shape[:axis]:
            outer *= d
        inner = 1
        for d in shape[axis + 1:]:
            inner *= d
        return outer, shape[axis], inner

    def reduce(self, op, x, axis, keep_dims=False):
        assert x.is_contiguous()
        axis %= x.ndim
        o, r, i = self._axis_view(x.shape, axis)
        shape = list(x.shape)
        if keep_dims:
            shape[axis] = 1
        else:
            shape.pop(axis)
        y = self.empty(shape)
        ffi.check(self.lib.agb_reduce(self.ctx, ffi.R[op], x.ptr, y.ptr, o, r, i))
        return y

    def argreduce(self, is_max, x, axis, keep_dims=False):
        assert x.is_contiguous()
        axis %= x.ndim
        o, r, i = self._axis_view(x.shape, axis)
        shape = list(x.shape)
        if keep_dims:
            shape[axis] = 1
        else:
            shape.pop(axis)
        y = self.empty(shape)
        ffi.check(self.lib.agb_argreduce(self.ctx, int(is_max), x.ptr, y.ptr, o, r, i))
        return y

    def softmax_like(self, kind, x, axis):
        assert x.is_contiguous()
        axis %= x.ndim
        o, r, i = self._axis_view(x.shape, axis)
        if kind == "logsumexp":
            shape = list(x.shape)
            shape[axis] = 1
            y = self.empty(shape)
        else:
            y = self.empty(x.shape)
        fn = {"softmax": self.lib.agb_softmax, "log_softmax": self.lib.agb_log_softmax, "logsumexp": self.lib.agb_logsumexp}[kind]
        ffi.check(fn(self.ctx, x.ptr, y.ptr, o, r, i))
        return y

    def sparse_xent_fwd(self, logits, labels):
        b, c = logits.shape
        loss, log_x = self.empty((b, 1)), self.empty((b, c))
        ffi.check(self.lib.agb_sparse_xent_fwd(self.ctx, logits.ptr, labels.ptr, loss.ptr, log_x.ptr, b, c))
        return loss, log_x

    def sparse_xent_bwd(self, log_x, labels, gy):
        b, c = log_x.shape
        gx = self.empty((b, c))
        ffi.check(self.lib.agb_sparse_xent_bwd(self.ctx, log_x.ptr, labels.ptr, gy.ptr, gy.size, gx.ptr, b, c))
        return gx

    def softmax_xent_fwd(self, logits, t):
        b, c = logits.shape
        loss, log_x = self.empty((b,)), self.empty((b, c))
        ffi.check(self.lib.agb_softmax_xent_fwd(self.ctx, logits.ptr, t.ptr, loss.ptr, log_x.ptr, b, c))
        return loss, log_x

    def gemm(self, a, b, trans_a=False, trans_b=False, out=None, beta=0.0):
        ash, bsh = list(a.shape), list(b.shape)
        m, k = (ash[-1], ash[-2]) if trans_a else (ash[-2], ash[-1])
        n = bsh[-2] if trans_b else bsh[-1]
        c = out if out is not None else self.empty(ash[:-2] + [m, n])
        ffi.check(self.lib.agb_gemm_f32(self.ctx, int(trans_a), int(trans_b), a.desc(), b.desc(), c.desc(), beta))
        return c

    @staticmethod
    def conv_out(x, k, pad, stride, dil):
        return (x + 2 * pad - (dil * (k - 1) + 1)) // stride + 1

    def empty_channels_last(self, shape):
        """logical [B,C,H,W] array whose memory order is N,H,W,C (strides {HWC, 1, WC, C})"""
        b, c, h, w = shape
        return self.empty((b, h, w, c)).transpose((0, 3, 1, 2))

    def upload_channels_last(self, a):
        return self.upload(np.ascontiguousarray(np.transpose(a, (0, 2, 3, 1)))).transpose((0, 3, 1, 2))

    def conv2d(self, x, w, pad=0, stride=1, dil=1, bias=None, relu=False, channels_last=False):
        b, _, h, wd = x.shape
        o, _, kh, kw = w.shape
        shape = (b, o, self.conv_out(h, kh, pad, stride, dil), self.conv_out(wd, kw, pad, stride, dil))
        y = self.empty_channels_last(shape) if channels_last else self.empty(shape)
        if bias is None and not relu:
            ffi.check(self.lib.agb_conv2d_fprop_f32(self.ctx, x.desc(), w.desc(), y.desc(), pad, stride, dil))
        else:
            ffi.check(self.lib.agb_conv2d_fprop_fused_f32(self.ctx, x.desc(), w.desc(), bias.ptr if bias is not None else None, int(relu), y.desc(), pad, stride, dil))
        return y

    def conv2d_relu_bits(self, x, w, pad=0, stride=1, dil=1, bias=None):
        """relu(conv2d(x, w) [+ bias]) channels-last plus the sign bits of the result (numel / 32 words, None when the kernel that ran did not
        write them): the 1/32-size mask the fused dgrad of the next layer reads instead of the activation"""
        b, _, h, wd = x.shape
        o, _, kh, kw = w.shape
        shape = (b, o, self.conv_out(h, kh, pad, stride, dil), self.conv_out(wd, kw, pad, stride, dil))
        y = self.empty_channels_last(shape)
        bits = self.empty(((int(np.prod(shape)) + 31) // 32,))
        written = C.c_int(0)
        ffi.check(self.lib.agb_conv2d_fprop_fused_bits_f32(self.ctx, x.desc(), w.desc(), bias.ptr if bias is not None else None, 1, y.desc(), bits.ptr,
                                                           C.byref(written), pad, stride, dil))
        return y, (bits if written.value else None)

    def conv2d_pool(self, x, w, pad=0, stride=1, dil=1, bias=None, relu=True):
        """max_pool2d([relu](conv2d(x, w) [+ bias]), 2, 0, 2) in one kernel; returns (y_pooled, idx_int32) channels-last, or None when the
        fused kernel does not take the layer (AGB_ERR_UNSUPPORTED)"""
        b, _, h, wd = x.shape
        o, _, kh, kw = w.shape
        yh, yw = self.conv_out(h, kh, pad, stride, dil), self.conv_out(wd, kw, pad, stride, dil)
        y, idx = self.empty_channels_last((b, o, yh // 2, yw // 2)), self.empty_channels_last((b, o, yh // 2, yw // 2))
        st = self.lib.agb_conv2d_fprop_pool_f32(self.ctx, x.desc(), w.desc(), bias.ptr if bias is not None else None, int(relu), y.desc(), idx.ptr, pad, stride, dil)
        if st == ffi.ERR_UNSUPPORTED:
            return None
        ffi.check(st)
        return y, idx

    def conv2d_transpose(self, gy, w, pad=0, stride=1, dil=1, mask_src=None, channels_last=False, chan_sum=False, mask_bits=None):
        """gx = conv2d_transpose(gy, w) [* (mask_src > 0)]; with chan_sum also returns sum_{b,h,w} gx as a [C] array"""
        b, _, yh, yw = gy.shape
        _, c, kh, kw = w.shape
        xh = stride * (yh - 1) - 2 * pad + (dil * (kh - 1) + 1)
        xw = stride * (yw - 1) - 2 * pad + (dil * (kw - 1) + 1)
        gx = self.empty_channels_last((b, c, xh, xw)) if channels_last else self.empty((b, c, xh, xw))
        if mask_src is None and not chan_sum:
            ffi.check(self.lib.agb_conv2d_dgrad_f32(self.ctx, gy.desc(), w.desc(), gx.desc(), pad, stride, dil))
            return gx
        cs = self.empty((c,)) if chan_sum else None
        if mask_bits is not None:
            ffi.check(self.lib.agb_conv2d_dgrad_fused_bits_f32(self.ctx, gy.desc(), w.desc(), mask_src.desc(), mask_bits.ptr,
                                                               cs.ptr if chan_sum else None, gx.desc(), pad, stride, dil))
        else:
            ffi.check(self.lib.agb_conv2d_dgrad_fused_f32(self.ctx, gy.desc(), w.desc(), mask_src.desc() if mask_src is not None else None,
                                                          cs.ptr if chan_sum else None, gx.desc(), pad, stride, dil))
        return (gx, cs) if chan_sum else gx

    def conv2d_filter_grad(self, img, g, wshape, pad=0, stride=1, dil=1):
        gw = self.empty(wshape)
        ffi.check(self.lib.agb_conv2d_wgrad_f32(self.ctx, img.desc(), g.desc(), gw.desc(), pad, stride, dil))
        return gw

    def im2col(self, x, kh, kw, pad=0, stride=1, dil=1):
        b, c, h, w = x.shape
        cols = self.empty((b, c, kh, kw, self.conv_out(h, kh, pad, stride, dil), self.conv_out(w, kw, pad, stride, dil)))
        ffi.check(self.lib.agb_im2col_f32(self.ctx, x.desc(), cols.desc(), kh, kw, pad, stride, dil))
        return cols

    def max_pool2d(self, x, size, pad=0, stride=1, int32_index=False):
        """idx holds float-encoded offsets (the reference's format) or, with int32_index, raw int32 (DArray.numpy().view(np.int32));
        y and idx take the memory order of x (NCHW or channels-last)"""
        b, c, h, w = x.shape
        yh, yw = (h + 2 * pad - size) // stride + 1, (w + 2 * pad - size) // stride + 1
        mk = self.empty if x.is_contiguous() else self.empty_channels_last
        y, idx = mk((b, c, yh, yw)), mk((b, c, yh, yw))
        ffi.check(self.lib.agb_maxpool2d_fwd(self.ctx, x.desc(), y.desc(), None if int32_index else idx.ptr, idx.ptr if int32_index else None, size, pad, stride))
        return y, idx

    def max_pool2d_grad(self, gy, idx, size, pad=0, stride=1, gate=None, int32_index=False, window_known=True, chan_sum=False):
        """gx = scatter(gy [* (gate > 0)]); gx takes the memory order of gy / idx"""
        b, c, yh, yw = gy.shape
        shape = (b, c, stride * (yh - 1) - 2 * pad + size, stride * (yw - 1) - 2 * pad + size)
        gx = self.empty(shape) if gy.is_contiguous() else self.empty_channels_last(shape)
        fi, ii = (None, idx.ptr) if int32_index else (idx.ptr, None)
        if gate is None and not window_known and not chan_sum:
            ffi.check(self.lib.agb_maxpool2d_bwd(self.ctx, gy.desc(), fi, ii, gx.desc()))
            return gx
        cs = self.empty((c,)) if chan_sum else None
        ffi.check(self.lib.agb_maxpool2d_bwd_fused(self.ctx, gy.desc(), fi, ii, gate.ptr if gate is not None else None, cs.ptr if chan_sum else None,
                                                   gx.desc(), size if window_known else 0, stride if window_known else 0))
        return (gx, cs) if chan_sum else gx

    def max_pool2d_grad_grad(self, ggx, idx, size, pad=0, stride=1):
        b, c, h, w = ggx.shape
        ggy = self.empty((b, c, (h + 2 * pad - size) // stride + 1, (w + 2 * pad - size) // stride + 1))
        ffi.check(self.lib.agb_maxpool2d_gradgrad(self.ctx, ggx.desc(), idx.ptr, None, ggy.desc()))
        return ggy

    def gather(self, param, indices, axis, normalize_negative=True):
        axis %= param.ndim
        pre, al, post = self._axis_view(param.shape, axis)
        out = self.empty(tuple(param.shape[:axis]) + tuple(indices.shape) + tuple(param.shape[axis + 1:]))
        ffi.check(self.lib.agb_gather(self.ctx, param.ptr, indices.ptr, out.ptr, pre, al, post, indices.size, int(normalize_negative)))
        return out

    def gather_grad(self, gy, indices, param_shape, axis):
        axis %= len(param_shape)
        pre, al, post = self._axis_view(tuple(param_shape), axis)
        gx = self.empty(param_shape)
        ffi.check(self.lib.agb_gather_grad(self.ctx, gy.ptr, indices.ptr, gx.ptr, pre, al, post, indices.size))
        return gx

    def random(self, kind, shape, p0=0.0, p1=1.0, seed=1, offset=0):
        """agb_random: uniform [p0, p1) / normal(p0, p1) / bernoulli(p0) / exp(rate p0) / log_normal(p0, p1) / gamma(shape p0, scale p1)."""
        y = self.empty(shape)
        ffi.check(self.lib.agb_random(self.ctx, ffi.RAND[kind], p0, p1, seed, offset, y.desc()))
        return y

    def scatter_add(self, gx, gy, indices, axis):
        """agb_scatter_add: gx[:, idx[j], :] += gy[:, j, :] in place (GatherGrad without its zero fill)."""
        axis %= len(gx.shape)
        pre, al, post = self._axis_view(tuple(gx.shape), axis)
        ffi.check(self.lib.agb_scatter_add(self.ctx, gy.ptr, indices.ptr, gx.ptr, pre, al, post, indices.size))
        return gx

    def dropout(self, x, ratio, mask=None, seed=0, offset=0):
        y = self.empty(x.shape)
        m = mask if mask is not None else self.empty(x.shape)
        ffi.check(self.lib.agb_dropout(self.ctx, x.desc(), y.desc(), m.desc(), ratio, seed, offset))
        return y, m

    def _plist(self, arrs):
        return (C.c_void_p * len(arrs))(*[a.ptr for a in arrs])

    def adam(self, ps, gs, ms, vs, ts, alpha=1e-3, eps=1e-8, b1=0.9, b2=0.999, grad_scale=1.0):
        sizes = (C.c_int64 * len(ps))(*[p.size for p in ps])
        ffi.check(self.lib.agb_multi_tensor_adam(self.ctx, len(ps), self._plist(ps), self._plist(gs), self._plist(ms),
                                                 self._plist(vs), self._plist(ts), sizes, alpha, eps, b1, b2, grad_scale))

    def sgd(self, ps, gs, alpha, grad_scale=1.0):
        sizes = (C.c_int64 * len(ps))(*[p.size for p in ps])
        ffi.check(self.lib.agb_multi_tensor_sgd(self.ctx, len(ps), self._plist(ps), self._plist(gs), sizes, alpha, grad_scale))

    def momentum(self, ps, gs, vs, lr, momentum, grad_scale=1.0):
        sizes = (C.c_int64 * len(ps))(*[p.size for p in ps])
        ffi.check(self.lib.agb_multi_tensor_momentum(self.ctx, len(ps), self._plist(ps), self._plist(gs), self._plist(vs), sizes, lr, momentum, grad_scale))

    def adagrad(self, ps, gs, hs, lr, grad_scale=1.0):
        sizes = (C.c_int64 * len(ps))(*[p.size for p in ps])
        ffi.check(self.lib.agb_multi_tensor_adagrad(self.ctx, len(ps), self._plist(ps), self._plist(gs), self._plist(hs), sizes, lr, grad_scale))
