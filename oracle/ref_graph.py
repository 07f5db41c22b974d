"""ORACLE (test infrastructure, NOT product code) — CPU restatement of rust-autograd's graph layer: symbolic graph,
reverse-mode `grad` built from each op's `Op::grad` composition, evaluator, variables and optimizers.  Numerics come from
oracle/ref_ops.py.  The public names equal those of rust_autograd_b200.autograd so that one test body can be executed
against the CUDA engine and against this oracle ("compare tensors forward and after grad", BASELINE.json north_star).

Reference files followed (under /root/reference/src): gradient.rs:20-248 (compute_gradients), evaluation.rs:252-362
(evaluation order, memoisation, variables/placeholders), tensor_ops/mod.rs (constructors), every `fn grad` listed in
SURVEY.md §11 (cited inline), optimizers/*.rs.  Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
import heapq
import math

import numpy as np

from . import ref_ops as R


class Panic(RuntimeError):
    pass


class EvalError(Exception):
    def __init__(self, kind, msg):
        self.kind = kind
        super().__init__("%s: %s" % (kind, msg))


class Tensor:
    __array_priority__ = 1000

    def __init__(self, graph, op, inputs=(), attrs=None, differentiable=True, backprop_inputs=None, selectors=None):
        self.graph, self.op, self.inputs, self.attrs = graph, op, list(inputs), attrs or {}
        self.selectors = list(selectors) if selectors else [0] * len(self.inputs)
        self.differentiable, self.backprop_inputs = differentiable, backprop_inputs
        self.topo_rank = max([i.topo_rank for i in self.inputs], default=-1) + 1
        self.id = len(graph.nodes)
        graph.nodes.append(self)

    def bp_inputs(self):
        return self.backprop_inputs if self.backprop_inputs is not None else self.inputs

    def is_source(self):
        return not self.inputs

    def _c(self, o):
        return o if isinstance(o, Tensor) else scalar(float(o), self.graph)

    def __add__(self, o): return add(self, self._c(o))
    def __radd__(self, o): return add(self._c(o), self)
    def __sub__(self, o): return sub(self, self._c(o))
    def __rsub__(self, o): return sub(self._c(o), self)
    def __mul__(self, o): return mul(self, self._c(o))
    def __rmul__(self, o): return mul(self._c(o), self)
    def __truediv__(self, o): return div(self, self._c(o))
    def __rtruediv__(self, o): return div(self._c(o), self)
    def __neg__(self): return neg(self)
    def eval(self, ctx=None, feeds=None): return (ctx or self.graph).evaluator().push(self).feeds(feeds).run()[0].unwrap()
    def reshape(self, s): return reshape(self, s)
    def flatten(self): return flatten(self)
    def squeeze(self, a): return squeeze(self, a)
    def expand_dims(self, a): return expand_dims(self, a)
    def transpose(self, a): return transpose(self, a)
    def access_elem(self, i): return Tensor(self.graph, "IndexOp", [self], {"index": i})
    def get_variable_id(self): return self.attrs.get("vid") if self.op == "Variable" else None


class Result:
    def __init__(self, value, err=None):
        self.value, self.err = value, err

    def is_ok(self):
        return self.err is None

    def unwrap(self):
        if self.err is not None:
            raise EvalError(self.err.kind, str(self.err))
        return self.value


class Feeder:
    def __init__(self):
        self.items = []

    def push(self, key, value):
        self.items.append((key, value))
        return self


class Evaluator:
    def __init__(self, graph):
        self.graph, self.targets, self.feeder = graph, [], Feeder()

    def push(self, x):
        self.targets.append(x)
        return self

    def extend(self, xs):
        self.targets.extend(xs)
        return self

    def feed(self, k, v):
        self.feeder.push(k, v)
        return self

    def feeds(self, feeds):
        if isinstance(feeds, Feeder):
            self.feeder = feeds
        elif feeds:
            for k, v in (feeds.items() if isinstance(feeds, dict) else feeds):
                self.feeder.push(k, v)
        return self

    def set_feeder(self, f):
        self.feeder = f
        return self

    def run(self):
        return self.graph.eval(self.targets, self.feeder.items)


class Context:
    def __init__(self, env):
        self.env, self.nodes, self.var_nodes = env, [], {}

    def placeholder(self, name, shape):
        return Tensor(self, "Placeholder", [], {"name": name, "shape": list(shape)})

    def variable(self, key):
        if isinstance(key, tuple):
            vid = self.env.name_to_id[(key[0], key[1])]
        elif isinstance(key, str):
            vid = self.env.name_to_id[("", key)]
        else:
            vid = int(key)
        if vid not in self.var_nodes:
            self.var_nodes[vid] = Tensor(self, "Variable", [], {"vid": vid})
        return self.var_nodes[vid]

    def evaluator(self):
        return Evaluator(self)

    def namespace(self, ns):
        return self.env.namespace(ns)

    def default_namespace(self):
        return self.env.namespace("")

    def size(self):
        return len(self.nodes)

    # ---- Graph::eval (evaluation.rs:252-362): every node at most once per run; optimizer ops applied after all gradients exist
    def eval(self, targets, feeds):
        memo, pending = {}, []

        def feed_of(t):
            for k, v in feeds:
                if (isinstance(k, Tensor) and k is t) or (not isinstance(k, Tensor) and k == t.attrs["name"]):
                    return np.asarray(v, dtype=R.OUT_DTYPE)
            raise Panic("Placeholder unfilled")

        def value(t, sel=0):
            if t.op == "Placeholder":
                return feed_of(t)
            if t.op == "Variable":
                return self.env.arrays[t.attrs["vid"]]
            if t.id not in memo:
                order, seen = [], set()
                stack = [(t, False)]
                while stack:       # iterative post-order
                    n, visit = stack.pop()
                    if n.id in memo or n.op in ("Placeholder", "Variable"):
                        continue
                    if visit:
                        try:
                            ins = [value(i, s) for i, s in zip(n.inputs, n.selectors)]
                            memo[n.id] = COMPUTE[n.op](n, ins, pending)
                        except R.OpError as e:
                            memo[n.id] = e
                    elif n.id not in seen:
                        seen.add(n.id)
                        stack.append((n, True))
                        for i in n.inputs:
                            stack.append((i, False))
            r = memo[t.id]
            if isinstance(r, R.OpError):
                raise r
            return r[sel]

        out = []
        for t in targets:
            try:
                out.append(Result(np.array(value(t), dtype=R.OUT_DTYPE, copy=True)))
            except R.OpError as e:
                out.append(Result(None, e))
        for fn in pending:
            fn()
        return out


class _Slot:
    def __init__(self, env, ns, name=None):
        self.env, self.ns, self._name = env, ns, name

    def name(self, n):
        return _Slot(self.env, self.ns, n)

    def set(self, v):
        vid = len(self.env.arrays)
        self.env.arrays.append(np.array(v, dtype=R.OUT_DTYPE, copy=True))
        self.env.name_to_id[(self.ns, self._name if self._name is not None else "anon%d" % vid)] = vid
        return vid


class Namespace:
    def __init__(self, env, ns):
        self.env, self.ns = env, ns

    def slot(self):
        return _Slot(self.env, self.ns)

    def current_var_ids(self):
        return sorted(v for (n, _), v in self.env.name_to_id.items() if n == self.ns)

    def get_array_by_name(self, name):
        v = self.env.name_to_id.get((self.ns, name))
        return None if v is None else self.env.arrays[v]


class VariableEnvironment:
    def __init__(self, device=0):
        self.arrays, self.name_to_id = [], {}

    def slot(self): return _Slot(self, "")
    def namespace(self, ns): return Namespace(self, ns)
    namespace_mut = namespace
    def default_namespace(self): return Namespace(self, "")
    default_namespace_mut = default_namespace
    def get_array_by_id(self, vid): return self.arrays[vid].copy()
    def set_array_by_id(self, vid, v): self.arrays[vid] = np.array(v, dtype=R.OUT_DTYPE, copy=True)
    def run(self, f): return f(Context(self))
    def close(self): pass


def run(f, device=0):
    return VariableEnvironment().run(f)


# ================================================================================================ constructors (tensor_ops/mod.rs)
def _g(ts):
    for t in ts:
        if isinstance(t, Tensor):
            return t.graph
    raise Panic("no tensor")


def as_tensor(v, g):
    return v if isinstance(v, Tensor) else Tensor(g, "Const", [], {"value": np.asarray([float(x) for x in v], dtype=np.float64), "meta": True})


def convert_to_tensor(arr, g): return Tensor(g, "Const", [], {"value": np.asarray(arr, dtype=R.OUT_DTYPE)})
def scalar(v, g): return Tensor(g, "Const", [], {"value": np.asarray(v, dtype=R.OUT_DTYPE)})
def zeros(shape, g): return Tensor(g, "Fill", [as_tensor(shape, g)], {"v": 0.0})
def ones(shape, g): return Tensor(g, "Fill", [as_tensor(shape, g)], {"v": 1.0})
def shape(x): return Tensor(x.graph, "Shape", [x], differentiable=False)
def rank(x): return Tensor(x.graph, "Rank", [x], differentiable=False)
def size(x): return Tensor(x.graph, "Size", [x], differentiable=False)
def identity(x): return Tensor(x.graph, "Identity", [x])
def nth_tensor(x, n): return Tensor(x.graph, "Identity", [x], selectors=[n])
def stop_gradient(x): return Tensor(x.graph, "StopGradient", [x], differentiable=False)


def _mk_unary(name):
    def f(x):
        return Tensor(x.graph, "Unary", [x], {"fn": name})
    f.__name__ = name
    return f


for _n in ["sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "exp2", "exp10", "ln", "log2",
           "log10", "sqrt", "neg", "abs", "sign", "floor", "ceil", "inv", "inv_sqrt", "square", "sigmoid", "relu", "softplus", "lgamma", "digamma"]:
    globals()[_n] = _mk_unary(_n)


def pow(x, a): return Tensor(x.graph, "Unary", [x], {"fn": "pow", "p": float(a)})
def elu(x, alpha): return Tensor(x.graph, "Unary", [x], {"fn": "elu", "p": float(alpha)})
def clip(x, lo, hi): return Tensor(x.graph, "Clip", [x], {"lo": float(lo), "hi": float(hi)})


def _mk_bin(name):
    def f(a, b):
        g = _g([a, b])
        a = a if isinstance(a, Tensor) else scalar(float(a), g)
        b = b if isinstance(b, Tensor) else scalar(float(b), g)
        return Tensor(g, "Bin", [a, b], {"fn": name})
    f.__name__ = name
    return f


for _n in ["add", "sub", "mul", "div"]:
    globals()[_n] = _mk_bin(_n)


def _mk_cmp(name):
    def f(a, b):
        g = _g([a, b])
        a = a if isinstance(a, Tensor) else scalar(float(a), g)
        b = b if isinstance(b, Tensor) else scalar(float(b), g)
        return Tensor(g, "Cmp", [a, b], {"fn": name})
    f.__name__ = name
    return f


for _n in ["equal", "not_equal", "greater", "lesser", "greater_equal", "lesser_equal", "maximum", "minimum"]:
    globals()[_n] = _mk_cmp(_n)


def add_n(xs): return xs[0] if len(xs) == 1 else Tensor(xs[0].graph, "AddN", list(xs))
def leaky_relu(x, alpha): return maximum(x, scalar(alpha, x.graph) * x)


def _mk_red(kind):
    def f(x, axes, keep_dims):
        return Tensor(x.graph, "Reduce", [x, as_tensor(axes, x.graph)], {"kind": kind, "keep": bool(keep_dims)})
    f.__name__ = "reduce_" + kind
    return f


reduce_sum, reduce_mean, reduce_prod, reduce_min, reduce_max = [_mk_red(k) for k in ["sum", "mean", "prod", "min", "max"]]


def reduce_variance(x, axes, keep_dims): return reduce_mean(square(x - reduce_mean(x, axes, True)), axes, keep_dims)
def sum_all(x): return Tensor(x.graph, "SumAll", [x])
def mean_all(x): return sum_all(x) / size(x)
def argmax(x, axis, keep_dim): return Tensor(x.graph, "Arg", [x], {"max": True, "axis": axis, "keep": keep_dim})
def argmin(x, axis, keep_dim): return Tensor(x.graph, "Arg", [x], {"max": False, "axis": axis, "keep": keep_dim})
def reduce_logsumexp(x, axis, keep_dim): return Tensor(x.graph, "LogSumExp", [x], {"axis": axis, "keep": keep_dim})
def softmax(x, axis): return Tensor(x.graph, "Softmax", [x], {"axis": axis})
def log_softmax(x, axis): return Tensor(x.graph, "LogSoftmax", [x], {"axis": axis})
def sigmoid_cross_entropy(y, t): return Tensor(y.graph, "SigmoidXent", [y, t])
def softmax_cross_entropy(y, t): return Tensor(y.graph, "SoftmaxXent", [y, t])
def sparse_softmax_cross_entropy(y, t): return Tensor(y.graph, "SparseXent", [y, t])
def mean_squared_error(y, t): return reduce_mean(square(y - t), [-1], False)
def matmul(a, b): return Tensor(a.graph, "MatMul", [a, b], {"ta": False, "tb": False, "batched": False})
def batch_matmul(a, b): return batch_matmul_t(a, b, False, False)
def batch_matmul_t(a, b, ta, tb): return Tensor(a.graph, "MatMul", [a, b], {"ta": ta, "tb": tb, "batched": True})


def tensordot(a, b, a_axes, b_axes):
    g = a.graph
    pre = Tensor(g, "TensordotPre", [a, b, as_tensor(a_axes, g), as_tensor(b_axes, g)])
    fs, pa, pb, nsa, nsb = [nth_tensor(pre, i) for i in range(5)]
    return reshape(matmul(reshape(transpose(a, pa), nsa), reshape(transpose(b, pb), nsb)), fs)


def reshape(x, shp): return Tensor(x.graph, "Reshape", [x, as_tensor(shp, x.graph)])
def flatten(x): return Tensor(x.graph, "Reshape", [x, scalar(-1.0, x.graph)])
def transpose(x, axes): return Tensor(x.graph, "Transpose", [x, as_tensor(axes, x.graph)], {"inv": False})
def squeeze(x, axes): return Tensor(x.graph, "Squeeze", [x, as_tensor(axes, x.graph)])
def expand_dims(x, axes): return Tensor(x.graph, "ExpandDims", [x, as_tensor(axes, x.graph)])


def slice(x, starts, ends):
    idx = []
    for s, e in zip(starts, ends):          # mod.rs:2181-2190
        idx.append((s, None) if e == -1 else (s, e + 1 if e < -1 else e))
    return Tensor(x.graph, "Slice", [x], {"idx": idx})


def split(x, sizes, axis):
    out, start = [], 0
    for sz in sizes:
        out.append(Tensor(x.graph, "Split", [x], {"axis": axis, "s": start, "e": start + sz}))
        start += sz
    return out


def concat(xs, axis): return Tensor(xs[0].graph, "Concat", list(xs), {"axis": axis})
def tile(x, axis, num): return Tensor(x.graph, "Tile", [x], {"axis": axis, "num": num})
def gather_common(param, indices, axis): return Tensor(param.graph, "Gather", [as_tensor(indices, param.graph), param], {"axis": axis})
def gather(param, indices, axis): return Tensor(param.graph, "Gather", [as_tensor(indices, param.graph), param], {"axis": axis})
def conv2d(x, w, pad, stride): return dilated_conv2d(x, w, pad, stride, 1)
def dilated_conv2d(x, w, pad, stride, dilate): return Tensor(x.graph, "Conv2D", [x, w], {"p": (pad, stride, dilate)})
def conv2d_transpose(x, w, pad, stride): return dilated_conv2d_transpose(x, w, pad, stride, 1)
def dilated_conv2d_transpose(x, w, pad, stride, dilate): return Tensor(x.graph, "Conv2DTranspose", [x, w], {"p": (pad, stride, dilate)})
def max_pool2d(x, pool_size, pad, stride): return Tensor(x.graph, "MaxPool2D", [x], {"size": pool_size, "pad": pad, "stride": stride})


# ---- forced discrete decisions (parity-test tooling, not reference API) -------------------------------------------------------------
# ReLU masks and max-pool argmaxes make a network's gradient a DISCONTINUOUS function of its forward values: one pre-activation within
# rounding of 0, or two window candidates within rounding of each other, and two correct implementations route a gradient value
# differently; after the sqrt(N) cancellation inside a filter gradient a single such flip is a 1e-3 relative difference (measured: the
# f32 C port of the reference against this f64 oracle, VGG stack at 128x128).  The full-size parity tests therefore (1) check that the
# device's decisions equal the oracle's except at verified near-ties, and (2) compare gradients with the oracle evaluated UNDER THE
# DEVICE'S DECISIONS, where everything left is rounding.
def relu_forced(x, mask):
    """relu with the 0/1 mask given: y = x * mask, dy/dx = mask (activation_ops.rs:154-166 with `greater(x, 0)` replaced by `mask`)"""
    return Tensor(x.graph, "ReLUForced", [x], {"mask": np.asarray(mask, dtype=R.OUT_DTYPE)})


def max_pool2d_forced(x, idx, pool_size, pad, stride):
    """max_pool2d with the argmax offsets given (flat offsets into the whole input, max_pool2d.rs:61): y = x.flat[idx]; backward unchanged"""
    return Tensor(x.graph, "MaxPool2DForced", [x], {"size": pool_size, "pad": pad, "stride": stride, "idx": np.asarray(idx).astype(np.int64)})
def dropout(x, ratio, train, seed=0, mask=None): return Tensor(x.graph, "Dropout", [x], {"ratio": ratio, "train": train, "mask": mask})


def normalize(x, axes):
    mean = reduce_mean(x, axes, True)
    centered = x - mean
    variance = reduce_mean(square(centered), axes, True)
    return centered * inv_sqrt(variance + scalar(1e-5, x.graph))


def batch_norm(x, scale, shift): return normalize(x, [0]) * scale + shift


# ================================================================================================ Op::compute dispatch
def _ints(a):
    return [int(v) for v in np.asarray(a).ravel()]


def _c_bin(n, ins, _):
    return [R.binary_arith(n.attrs["fn"], ins[0], ins[1])]


def _c_unary(n, ins, _):
    fn = n.attrs["fn"]
    if fn == "inv_sqrt":
        fn = "invsqrt"
    return [R.unary(fn, ins[0], n.attrs.get("p", 0.0))]


def _c_reduce(n, ins, _):
    return [R.reduce(n.attrs["kind"], ins[0], _ints(ins[1]), n.attrs["keep"])]


def _c_reshape(n, ins, _):
    x, s = np.asarray(ins[0]), [float(v) for v in np.asarray(ins[1]).ravel()]
    prod = 1.0
    for v in s:
        prod *= v
    target = [int(v) if v != -1 else int(x.size // int(-prod)) for v in s]
    if int(np.prod(target, dtype=np.int64)) != x.size:
        raise R.OpError("IncompatibleShape", "reshape failed: %s vs %s" % (x.shape, target))
    return [x.reshape(target)]


def _c_transpose(n, ins, _):
    perm = _ints(ins[1])
    if len(perm) != ins[0].ndim:
        raise R.OpError("IncompatibleShape", "transpose: inputs's ndim and axes's length must match")
    dims = [0] * len(perm)
    for i, d in enumerate(perm):
        if n.attrs["inv"]:
            dims[d] = i
        else:
            dims[i] = d
    return [np.transpose(ins[0], dims)]


def _slice_obj(idx):
    return tuple(np.s_[s:e] for s, e in idx)


def _c_conv(n, ins, _):
    pad, stride, dil = n.attrs["p"]
    y, cols = R.conv2d(ins[0], ins[1], pad, stride, dil, return_cols=True)
    return [y, cols]


def _c_filter_grad(n, ins, _):
    cols, gy, w = _f64(ins[0]), _f64(ins[1]), ins[2]
    B, C, kh, kw, yh, yw = cols.shape
    O = gy.shape[1]
    gw = np.einsum("bop,bkp->ok", gy.reshape(B, O, yh * yw), cols.reshape(B, C * kh * kw, yh * yw), optimize=True)
    return [gw.reshape(np.asarray(w).shape).astype(R.OUT_DTYPE)]


def _c_conv_with_cols(n, ins, _):
    cols, w = _f64(ins[0]), _f64(ins[1])
    B, C, kh, kw, yh, yw = cols.shape
    O = w.shape[0]
    y = np.einsum("ok,bkp->bop", w.reshape(O, -1), cols.reshape(B, C * kh * kw, yh * yw), optimize=True)
    return [y.reshape(B, O, yh, yw).astype(R.OUT_DTYPE)]


def _f64(a):
    return np.asarray(a, dtype=np.float64)


def _c_dropout(n, ins, _):
    x = ins[0]
    if not n.attrs["train"]:
        return [R.dropout(x, None, n.attrs["ratio"], train=False)]
    mask = n.attrs["mask"]
    if mask is None:
        raise Panic("oracle dropout needs an explicit mask (RNG streams are parity-unpinned)")
    mask = np.asarray(mask, dtype=R.OUT_DTYPE)
    return [R.dropout(x, mask, n.attrs["ratio"]), mask]


def _c_update(n, ins, pending):
    env, kind, h = n.graph.env, n.attrs["kind"], n.attrs["h"]
    vids = [i.attrs["vid"] if i.op == "Variable" else None for i in n.inputs]
    g = np.array(ins[1], copy=True)

    def apply():
        A = env.arrays
        if kind == "adam":
            p, m, v, t = R.adam_update(A[vids[0]], g, A[vids[2]], A[vids[3]], A[vids[4]], *h)
            A[vids[0]], A[vids[2]], A[vids[3]], A[vids[4]] = p, m, v, np.asarray(t, dtype=R.OUT_DTYPE).reshape(A[vids[4]].shape)
        elif kind == "sgd":
            A[vids[0]] = R.sgd_update(A[vids[0]], g, h[0])
        elif kind == "momentum":
            A[vids[0]], A[vids[2]] = R.momentum_sgd_update(A[vids[0]], g, A[vids[2]], h[0], h[1])
        else:
            A[vids[0]], A[vids[2]] = R.adagrad_update(A[vids[0]], g, A[vids[2]], h[0])
    pending.append(apply)       # applied after every gradient of the run exists (same ordering as the CUDA engine)
    return [np.zeros((), dtype=R.OUT_DTYPE)]


def _c_gather(n, ins, _):
    return [R.gather(ins[1], ins[0], n.attrs["axis"])]


def _c_tensordot_pre(n, ins, _):
    x0, x1 = np.asarray(ins[0]), np.asarray(ins[1])
    a0 = [a + x0.ndim if a < 0 else a for a in _ints(ins[2])]
    a1 = [a + x1.ndim if a < 0 else a for a in _ints(ins[3])]

    def pre(shp, axes, flip):
        free = [i for i in range(len(shp)) if i not in axes]
        pf = int(np.prod([shp[i] for i in free], dtype=np.int64))
        pa = int(np.prod([shp[i] for i in axes], dtype=np.int64))
        perm = (axes + free) if flip else (free + axes)
        return perm, ([pa, pf] if flip else [pf, pa]), [shp[i] for i in free]
    p0, ns0, f0 = pre(x0.shape, a0, False)
    p1, ns1, f1 = pre(x1.shape, a1, True)
    F = lambda v: np.asarray(v, dtype=np.float64)
    return [F(f0 + f1), F(p0), F(p1), F(ns0), F(ns1)]


def _maybe_reduce(n, ins, _):
    return [R.maybe_reduce_sum(ins[0], _ints(ins[1]))]


def _maybe_broadcast(n, ins, _):
    target = tuple(_ints(ins[1]))
    x = np.asarray(ins[0])
    if x.shape == target:
        return [x]
    if R._is_scalar_shape(x.shape):
        x = x.reshape((1,) * len(target))
    return [R.broadcast_to(x, target)]


def _index_grad(n, ins, _):
    x, gy = np.asarray(ins[0]), ins[1]
    r = np.zeros(x.shape, dtype=R.OUT_DTYPE)
    r.ravel()[n.attrs["index"]] = np.asarray(gy).reshape(())
    return [r]


def _slice_grad(n, ins, _):
    gx = np.zeros(np.asarray(ins[0]).shape, dtype=R.OUT_DTYPE)
    gx[_slice_obj(n.attrs["idx"])] = ins[1]
    return [gx]


def _split_idx(n, x):
    ax = n.attrs["axis"] + x.ndim if n.attrs["axis"] < 0 else n.attrs["axis"]
    return [(n.attrs["s"], n.attrs["e"]) if k == ax else (0, None) for k in range(x.ndim)]


def _split_grad(n, ins, _):
    x = np.asarray(ins[0])
    gx = np.zeros(x.shape, dtype=R.OUT_DTYPE)
    gx[_slice_obj(_split_idx(n, x))] = ins[1]
    return [gx]


def _concat_grad(n, ins, _):
    gy = np.asarray(ins[0])
    ax = n.attrs["axis"] + gy.ndim if n.attrs["axis"] < 0 else n.attrs["axis"]
    start = sum(np.asarray(ins[1 + i]).shape[ax] for i in range(n.attrs["index"]))
    ln = np.asarray(ins[1 + n.attrs["index"]]).shape[ax]
    sl = [np.s_[:]] * gy.ndim
    sl[ax] = np.s_[start:start + ln]
    return [gy[tuple(sl)]]


def _squeeze(n, ins, _):
    x = np.asarray(ins[0])
    axes = sorted(_ints(ins[1]))
    for adjust, i in enumerate(axes):
        ax = (x.ndim + i if i < 0 else i) - adjust
        assert x.shape[ax] == 1, "Can't squeeze a dim whose size != 1"
        x = np.squeeze(x, ax)
    return [x]


def _expand(n, ins, _):
    x = np.asarray(ins[0])
    shp = list(x.shape)
    for i in sorted(_ints(ins[1])):
        shp.insert(x.ndim + i if i < 0 else i, 1)
    return [x.reshape(shp)]


def _reduce_grad_common(n, ins, _):
    gy, target = np.asarray(ins[0]), tuple(_ints(ins[1]))
    if gy.shape == target:
        return [gy]
    if n.attrs["mk"] or R._is_scalar_shape(gy.shape):
        axes = sorted(a + len(target) if a < 0 else a for a in _ints(ins[2]))
        shp = list(gy.shape)
        for a in axes:
            shp.insert(a, 1)
        gy = gy.reshape(shp)
    return [R.broadcast_to(gy, target)]


COMPUTE = {
    "Const": lambda n, ins, _: [n.attrs["value"]],
    "Fill": lambda n, ins, _: [np.full(_ints(ins[0]), n.attrs["v"], dtype=R.OUT_DTYPE)],
    "Shape": lambda n, ins, _: [np.asarray(np.asarray(ins[0]).shape, dtype=np.float64)],
    "Rank": lambda n, ins, _: [np.asarray(float(np.asarray(ins[0]).ndim))],
    "Size": lambda n, ins, _: [np.asarray(float(np.asarray(ins[0]).size))],
    "Identity": lambda n, ins, _: [ins[0]], "StopGradient": lambda n, ins, _: [ins[0]],
    "SetDiff1D": lambda n, ins, _: [np.asarray(sorted(set(np.asarray(ins[0]).ravel().tolist()) - set(np.asarray(ins[1]).ravel().tolist())), dtype=R.OUT_DTYPE)],
    "Map": lambda n, ins, _: [np.asarray(n.attrs["f"](np.asarray(ins[0])), dtype=R.OUT_DTYPE)],
    "Bin": _c_bin, "Unary": _c_unary, "Reduce": _c_reduce,
    "Clip": lambda n, ins, _: [R.unary("clip", ins[0], n.attrs["lo"], n.attrs["hi"])],
    "ClipGrad": lambda n, ins, _: [R.clip_grad(ins[0], ins[1], n.attrs["lo"], n.attrs["hi"])],
    "ELUGrad": lambda n, ins, _: [R.elu_grad(ins[0], ins[1], n.attrs["alpha"])],
    "Cmp": lambda n, ins, _: [R.compare(n.attrs["fn"], ins[0], ins[1])],
    "AddN": lambda n, ins, _: [R.add_n(ins)],
    "SumAll": lambda n, ins, _: [R.sum_all(ins[0])],
    "SumAllGrad": lambda n, ins, _: [np.full(_ints(ins[1]), np.asarray(ins[0]).reshape(()), dtype=R.OUT_DTYPE)],
    "Arg": lambda n, ins, _: [R.arg_reduce(ins[0], n.attrs["axis"], n.attrs["keep"], n.attrs["max"])],
    "LogSumExp": lambda n, ins, _: [R.logsumexp(ins[0], n.attrs["axis"], n.attrs["keep"])],
    "Softmax": lambda n, ins, _: [R.softmax(ins[0], n.attrs["axis"])],
    "LogSoftmax": lambda n, ins, _: [R.log_softmax(ins[0], n.attrs["axis"])],
    "SigmoidXent": lambda n, ins, _: [R.sigmoid_cross_entropy(ins[0], ins[1])],
    "SoftmaxXent": lambda n, ins, _: list(R.softmax_cross_entropy(ins[0], ins[1])),
    "SparseXent": lambda n, ins, _: list(R.sparse_softmax_cross_entropy(ins[0], ins[1])),
    "SparseXentGrad": lambda n, ins, _: [R.sparse_softmax_cross_entropy_grad(ins[0], ins[1], ins[2])],
    "MatMul": lambda n, ins, _: [(R.batch_matmul if n.attrs["batched"] else R.matmul)(ins[0], ins[1], n.attrs["ta"], n.attrs["tb"])],
    "TensordotPre": _c_tensordot_pre, "Reshape": _c_reshape, "Transpose": _c_transpose, "Squeeze": _squeeze, "ExpandDims": _expand,
    "Slice": lambda n, ins, _: [np.asarray(ins[0])[_slice_obj(n.attrs["idx"])]], "SliceGrad": _slice_grad,
    "Split": lambda n, ins, _: [np.asarray(ins[0])[_slice_obj(_split_idx(n, np.asarray(ins[0])))]], "SplitGrad": _split_grad,
    "Concat": lambda n, ins, _: [np.concatenate([np.asarray(i) for i in ins], axis=n.attrs["axis"])], "ConcatGrad": _concat_grad,
    "Tile": lambda n, ins, _: [np.concatenate([np.asarray(ins[0])] * n.attrs["num"], axis=n.attrs["axis"])],
    "Gather": _c_gather,
    "GatherGrad": lambda n, ins, _: [R.gather_grad(ins[0], np.asarray(ins[1]).shape, ins[2], n.attrs["axis"])],
    "IndexOp": lambda n, ins, _: [np.asarray(np.asarray(ins[0]).ravel()[n.attrs["index"]])], "IndexOpGrad": _index_grad,
    "MaybeReduceSum": _maybe_reduce, "MaybeBroadcast": _maybe_broadcast, "ReduceGradCommon": _reduce_grad_common,
    "Conv2D": _c_conv, "Conv2DWithCols": _c_conv_with_cols, "Conv2DFilterGrad": _c_filter_grad,
    "Conv2DTranspose": lambda n, ins, _: [R.conv2d_transpose(ins[0], ins[1], *n.attrs["p"])],
    "Conv2DTransposeFilterGrad": lambda n, ins, _: [R.conv2d_transpose_filter_grad(ins[0], ins[1], np.asarray(ins[2]).shape, *n.attrs["p"])],
    "MaxPool2D": lambda n, ins, _: list(R.max_pool2d(ins[0], n.attrs["size"], n.attrs["pad"], n.attrs["stride"])[:2]),
    "MaxPool2DGrad": lambda n, ins, _: [R.max_pool2d_grad(ins[0], ins[1], n.attrs["size"], n.attrs["pad"], n.attrs["stride"])],
    "MaxPool2DGradGrad": lambda n, ins, _: [R.max_pool2d_grad_grad(ins[0], ins[1], n.attrs["size"], n.attrs["pad"], n.attrs["stride"])],
    "Dropout": _c_dropout, "Update": _c_update,
    "ReLUForced": lambda n, ins, _: [(np.asarray(ins[0]) * n.attrs["mask"]).astype(R.OUT_DTYPE)],
    "MaxPool2DForced": lambda n, ins, _: [np.asarray(ins[0]).ravel()[n.attrs["idx"]], n.attrs["idx"].astype(R.OUT_DTYPE)],
}


# ================================================================================================ Op::grad compositions
def _maybe_reduce_t(target_shape, x):
    return Tensor(x.graph, "MaybeReduceSum", [x, target_shape])


def _grad_bin(y, gy):
    x0, x1, fn = y.inputs[0], y.inputs[1], y.attrs["fn"]
    s0, s1 = shape(x0), shape(x1)
    if fn == "add":
        return [_maybe_reduce_t(s0, gy), _maybe_reduce_t(s1, gy)]                       # binary_ops.rs:154-165
    if fn == "sub":
        return [_maybe_reduce_t(s0, gy), neg(_maybe_reduce_t(s1, gy))]                  # :195-206
    if fn == "mul":
        return [_maybe_reduce_t(s0, gy * x1), _maybe_reduce_t(s1, gy * x0)]             # :218-236
    return [_maybe_reduce_t(s0, gy / x1), _maybe_reduce_t(s1, neg(x0) * pow(x1, -2.0) * gy)]   # :273-289


def _grad_unary(y, gy):
    x, fn, g = y.inputs[0], y.attrs["fn"], y.graph
    S = lambda v: scalar(v, g)
    table = {                                                    # math_ops.rs / activation_ops.rs
        "sin": lambda: cos(x) * gy, "cos": lambda: neg(sin(x) * gy), "tan": lambda: gy / square(cos(x)),
        "asin": lambda: inv_sqrt(S(1.0) - square(x)) * gy, "acos": lambda: neg(inv_sqrt(S(1.0) - square(x))) * gy,
        "atan": lambda: inv(square(x) + S(1.0)) * gy, "sinh": lambda: cosh(x) * gy, "cosh": lambda: sinh(x) * gy,
        "tanh": lambda: gy * (S(1.0) - square(y)), "asinh": lambda: inv(sqrt(square(x) + S(1.0))) * gy,
        "acosh": lambda: inv(sqrt(square(x) - S(1.0))) * gy, "atanh": lambda: inv(S(1.0) - square(x)) * gy,
        "exp": lambda: y * gy, "exp2": lambda: S(math.log(2.0)) * y * gy, "exp10": lambda: S(math.log(10.0)) * y * gy,
        "ln": lambda: gy / x, "log2": lambda: gy / (S(math.log(2.0)) * x), "log10": lambda: gy / (S(math.log(10.0)) * x),
        "sqrt": lambda: gy * (S(0.5) * pow(x, -0.5)), "pow": lambda: gy * S(y.attrs["p"]) * pow(x, y.attrs["p"] - 1.0),
        "neg": lambda: neg(gy), "abs": lambda: gy * sign(x), "inv": lambda: neg(square(y)) * gy,
        "inv_sqrt": lambda: S(-0.5) * pow(x, -1.5) * gy, "square": lambda: S(2.0) * x * gy,
        "sigmoid": lambda: gy * (y - square(y)), "relu": lambda: mul(greater(x, S(0.0)), gy),
        "softplus": lambda: gy * (exp(x) / (exp(x) + S(1.0))),
        "lgamma": lambda: gy * digamma(x),                                               # math_ops.rs:1047-1052; Digamma: None (:1036-1039)
        "elu": lambda: Tensor(g, "ELUGrad", [x, gy], {"alpha": y.attrs["p"]}),
    }
    return [table[fn]()] if fn in table else [None]


def _rgc(gy, x, axes, keep):
    return Tensor(gy.graph, "ReduceGradCommon", [gy, shape(x), axes], {"mk": not keep})


def _grad_reduce(y, gy):
    x, axes, kind, keep = y.inputs[0], y.inputs[1], y.attrs["kind"], y.attrs["keep"]
    if kind == "sum":
        return [_rgc(gy, x, axes, keep), None]                                          # reduction_ops.rs:172-184
    if kind == "mean":                                                                  # :217-238
        return [_rgc(gy, x, axes, keep) / reduce_prod(gather_common(shape(x), axes, 0), [0], False), None]
    if kind == "prod":
        return [_rgc(gy * y, x, axes, keep) / x, None]                                  # :256-273
    return [mul(equal(x, _rgc(y, x, axes, keep)), _rgc(gy, x, axes, keep)), None]       # :332-363


def _grad_conv(y, gy):
    x, w, p = y.inputs[0], y.inputs[1], y.attrs["p"]
    g = y.graph                                                                         # conv2d.rs:556-586
    gx = Tensor(g, "Conv2DTranspose", [gy, w], {"p": p})
    gw = Tensor(g, "Conv2DFilterGrad", [nth_tensor(y, 1), gy, w], {"p": p}, backprop_inputs=[x, gy])
    return [gx, gw]


def _grad_filter_grad(y, gy):                                                           # conv2d.rs:746-775
    cols, g_y, p, g = y.inputs[0], y.inputs[1], y.attrs["p"], y.graph
    gx = Tensor(g, "Conv2DTranspose", [g_y, gy], {"p": p})
    ggy = Tensor(g, "Conv2DWithCols", [cols, gy], {"p": p}, backprop_inputs=[y.bp_inputs()[0], gy])
    return [gx, ggy]


def _grad_conv_with_cols(y, gy):                                                        # conv2d.rs:599-628
    cols, w, p, g = y.inputs[0], y.inputs[1], y.attrs["p"], y.graph
    gx = Tensor(g, "Conv2DTranspose", [gy, w], {"p": p})
    gw = Tensor(g, "Conv2DFilterGrad", [cols, gy, w], {"p": p}, backprop_inputs=[y.bp_inputs()[0], gy])
    return [gx, gw]


def _grad_conv_transpose(y, gy):                                                        # conv2d_transpose.rs:274-300
    x, w, p, g = y.inputs[0], y.inputs[1], y.attrs["p"], y.graph
    return [Tensor(g, "Conv2D", [gy, w], {"p": p}), Tensor(g, "Conv2DTransposeFilterGrad", [gy, x, stop_gradient(w)], {"p": p})]


def _grad_conv_transpose_fg(y, gw):                                                     # conv2d_transpose.rs:453-479
    gy, x, p, g = y.inputs[0], y.inputs[1], y.attrs["p"], y.graph
    return [Tensor(g, "Conv2DTranspose", [x, gw], {"p": p}), Tensor(g, "Conv2D", [gy, gw], {"p": p}), None]


def _grad_sparse_xent(y, gy):                                                           # xent_ops.rs:115-136
    t, log_x = y.inputs[1], nth_tensor(y, 1)
    gx1 = Tensor(y.graph, "SparseXentGrad", [log_x, t, gy])
    x = exp(log_x)
    gx2 = x * gy * (reduce_sum(x * log_x, [1], True) - log_x)
    return [gx1, gx2]


def _grad_softmax_xent(y, gy):                                                          # xent_ops.rs:179-201
    log_x, t = nth_tensor(y, 1), y.inputs[1]
    x = exp(log_x)
    return [(x - t) * gy, gy * (reduce_sum(x * log_x, [-1], True) - log_x) * y]


def _grad_sigmoid_xent(y, gy):                                                          # xent_ops.rs:48-60
    x, t = y.inputs
    e = exp(x)
    return [((e / (scalar(1.0, y.graph) + e)) - t) * gy, neg(gy * t)]


def _grad_concat(y, gy):
    return [Tensor(y.graph, "ConcatGrad", [gy] + y.inputs, {"index": i, "axis": y.attrs["axis"]}) for i in range(len(y.inputs))]


GRAD = {
    "Bin": _grad_bin, "Unary": _grad_unary, "Reduce": _grad_reduce,
    "Identity": lambda y, gy: [gy], "StopGradient": lambda y, gy: [None],
    "Clip": lambda y, gy: [Tensor(y.graph, "ClipGrad", [y.inputs[0], gy], {"lo": y.attrs["lo"], "hi": y.attrs["hi"]})],
    "Cmp": lambda y, gy: ([mul(equal(y.inputs[0], y), gy), mul(equal(y.inputs[1], y), gy)] if y.attrs["fn"] in ("maximum", "minimum") else [None]),
    "AddN": lambda y, gy: [gy] * len(y.inputs),
    "SumAll": lambda y, gy: [Tensor(y.graph, "SumAllGrad", [gy, shape(y.inputs[0])])],                        # reduction_ops.rs:130-136
    "SumAllGrad": lambda y, gy: [sum_all(gy), None],
    "LogSumExp": lambda y, gy: [softmax(y.inputs[0], y.attrs["axis"]) * gy],                                  # math_ops.rs:602-608
    "Softmax": lambda y, gy: [(gy - reduce_sum(y * gy, [y.attrs["axis"]], True)) * y],                        # activation_ops.rs:105-110
    "LogSoftmax": lambda y, gy: [gy - exp(y) * reduce_sum(gy, [1], True)],                                    # xent_ops.rs:24-30
    "SigmoidXent": _grad_sigmoid_xent, "SoftmaxXent": _grad_softmax_xent, "SparseXent": _grad_sparse_xent,
    "MatMul": lambda y, gy: [Tensor(y.graph, "MatMul", [gy, y.inputs[1]], {"ta": False, "tb": True, "batched": y.attrs["batched"]}),
                             Tensor(y.graph, "MatMul", [y.inputs[0], gy], {"ta": True, "tb": False, "batched": y.attrs["batched"]})],   # dot_ops.rs:608-628
    "Reshape": lambda y, gy: [Tensor(y.graph, "Reshape", [gy, shape(y.inputs[0])]), None],                   # array_ops.rs:229-238
    "Transpose": lambda y, gy: [Tensor(y.graph, "Transpose", [gy, y.inputs[1]], {"inv": not y.attrs["inv"]}), None],
    "Squeeze": lambda y, gy: [expand_dims(gy, y.inputs[1]), None], "ExpandDims": lambda y, gy: [squeeze(gy, y.inputs[1]), None],
    "Slice": lambda y, gy: [Tensor(y.graph, "SliceGrad", [y.inputs[0], gy], {"idx": y.attrs["idx"]})],
    "Split": lambda y, gy: [Tensor(y.graph, "SplitGrad", [y.inputs[0], gy], dict(y.attrs))],
    "Concat": _grad_concat, "Tile": lambda y, gy: [reduce_sum(gy, [y.attrs["axis"]], True)],
    "Gather": lambda y, gy: [None, Tensor(y.graph, "GatherGrad", [y.inputs[0], y.inputs[1], gy], {"axis": y.attrs["axis"]})],
    "IndexOp": lambda y, gy: [Tensor(y.graph, "IndexOpGrad", [y.inputs[0], gy], {"index": y.attrs["index"]})],
    "MaybeReduceSum": lambda y, gy: [Tensor(y.graph, "MaybeBroadcast", [gy, shape(y.inputs[0])]), None],    # binary_ops.rs:96-104
    "MaybeBroadcast": lambda y, gy: [_maybe_reduce_t(shape(y.inputs[0]), gy), None],
    "ReduceGradCommon": lambda y, gy: [Tensor(y.graph, "Reduce", [gy, y.inputs[2]], {"kind": "sum", "keep": y.attrs["mk"]}), None, None],
    "Conv2D": _grad_conv, "Conv2DFilterGrad": _grad_filter_grad, "Conv2DWithCols": _grad_conv_with_cols,
    "Conv2DTranspose": _grad_conv_transpose, "Conv2DTransposeFilterGrad": _grad_conv_transpose_fg,
    "MaxPool2D": lambda y, gy: [Tensor(y.graph, "MaxPool2DGrad", [gy, nth_tensor(y, 1)], dict(y.attrs))],
    "MaxPool2DGrad": lambda y, gy: [Tensor(y.graph, "MaxPool2DGradGrad", [gy, y.inputs[1]], dict(y.attrs)), None],
    "Dropout": lambda y, gy: [gy * nth_tensor(y, 1)],
    "ReLUForced": lambda y, gy: [gy * convert_to_tensor(y.attrs["mask"], y.graph)],
    "MaxPool2DForced": lambda y, gy: [Tensor(y.graph, "MaxPool2DGrad", [gy, nth_tensor(y, 1)], {k: y.attrs[k] for k in ("size", "pad", "stride")})],
}


def compute_gradients(ys, xs, gys=None):
    """gradient.rs:20-84 + init_gradient_map :203-248"""
    g = ys[0].graph
    info = {}
    is_x = lambda t: any(t is x for x in xs)
    stack = [(y, False) for y in ys]
    while stack:
        cur, visit = stack.pop()
        if visit:
            on = cur.differentiable and (is_x(cur) or any(info.get(c.id, {"on": False})["on"] for c in cur.bp_inputs()))
            info[cur.id] = {"on": on, "grads": []}
        else:
            stack.append((cur, True))
            for c in cur.bp_inputs():
                if cur.id in info:
                    continue
                if c.is_source() or not c.differentiable:
                    info[c.id] = {"on": c.differentiable and is_x(c), "grads": []}
                else:
                    stack.append((c, False))

    def gradient(i):
        if len(i["grads"]) > 1:
            i["grads"] = [add_n(i["grads"])]
        return i["grads"][0]
    if gys is not None:
        for y, gy in zip(ys, gys):
            info[y.id]["grads"].append(gy)
    else:
        one = scalar(1.0, g)
        for y in ys:
            info[y.id]["grads"].append(one)
    heap = [(-y.topo_rank, k, y) for k, y in enumerate(ys)]
    heapq.heapify(heap)
    cnt = len(heap)
    while heap:
        _, _, y = heapq.heappop(heap)
        gy = gradient(info[y.id])
        gxs = GRAD[y.op](y, gy) if y.op in GRAD else [None] * len(y.inputs)
        for x, gx in zip(y.bp_inputs(), gxs):
            xi = info.get(x.id)
            if xi is None or not xi["on"] or gx is None:
                continue
            first = not xi["grads"]
            xi["grads"].append(gx)
            if not x.is_source() and first:
                cnt += 1
                heapq.heappush(heap, (-x.topo_rank, cnt, x))
    return [gradient(info[x.id]) if (x.id in info and info[x.id]["on"] and info[x.id]["grads"]) else None for x in xs]


def grad(ys, xs):
    """tensor_ops/mod.rs:94-114"""
    gs = compute_gradients([sum_all(y) for y in ys], xs)
    return [gx if gx is not None else zeros(shape(x), x.graph) for x, gx in zip(xs, gs)]


def grad_with_default(ys, xs, ys_grads):
    gs = compute_gradients(list(ys), xs, list(ys_grads))
    return [gx if gx is not None else zeros(shape(x), x.graph) for x, gx in zip(xs, gs)]


def _hessian_vector_product(ys, xs, vectors):
    """tensor_ops/mod.rs:218-236"""
    return grad([gx * v for gx, v in zip(grad(ys, xs), vectors)], xs)


def setdiff1d(a, b): return Tensor(a.graph, "SetDiff1D", [a, b], differentiable=False)       # mod.rs:2044-2057, array_ops.rs:241-279
def map(x, f): return Tensor(x.graph, "Map", [x], {"f": f}, differentiable=False)            # noqa: A001  mod.rs:2931-2945, higher_order_ops.rs:5-36


def jacobians(y, xs, objective_len):
    vv = [grad([y.access_elem(i)], xs) for i in range(objective_len)]
    return [concat([expand_dims(flatten(v[i]), [0]) for v in vv], 0) for i in range(len(xs))]


# ================================================================================================ optimizers (optimizers/*.rs)
class _Opt:
    kind, names = None, []

    def _state(self, g, vid):
        return [g.variable((self.ns, "%d%s" % (vid, s))) for s in self.names]

    def compute_updates(self, params, grads, g):
        return [Tensor(g, "Update", [p, gr] + self._state(g, p.attrs["vid"]), {"kind": self.kind, "h": self.h}) for p, gr in zip(params, grads)]

    def get_update_op(self, params, grads, g):
        return add_n(self.compute_updates(params, grads, g))

    def update(self, params, grads, g, feeder=None):
        for r in g.evaluator().extend(self.compute_updates(params, grads, g)).set_feeder(feeder or Feeder()).run():
            r.unwrap()


class optimizers:
    class Adam(_Opt):
        kind, names = "adam", ["m", "v", "t"]

        def __init__(self, alpha, eps, b1, b2, var_id_list, env, namespace_id):
            self.h, self.ns = (alpha, eps, b1, b2), namespace_id
            for vid in var_id_list:                      # optimizers/adam.rs:84-103
                ns = env.namespace(namespace_id)
                ns.slot().name("%dm" % vid).set(np.zeros_like(env.arrays[vid]))
                ns.slot().name("%dv" % vid).set(np.zeros_like(env.arrays[vid]))
                ns.slot().name("%dt" % vid).set(np.ones((), dtype=R.OUT_DTYPE))

        @staticmethod
        def default(namespace_id, var_id_list, env):
            return optimizers.Adam(0.001, 1e-08, 0.9, 0.999, var_id_list, env, namespace_id)

    class SGD(_Opt):
        kind, names, ns = "sgd", [], ""

        def __init__(self, alpha):
            self.h = (alpha,)

    class MomentumSGD(_Opt):
        kind, names = "momentum", [""]

        def __init__(self, alpha, momentum, var_id_list, env, namespace_id):
            self.h, self.ns = (alpha, momentum), namespace_id
            for vid in var_id_list:
                env.namespace(namespace_id).slot().name("%d" % vid).set(np.zeros_like(env.arrays[vid]))

        @staticmethod
        def default(namespace_id, var_id_list, env):
            return optimizers.MomentumSGD(0.01, 0.9, var_id_list, env, namespace_id)

    class AdaGrad(_Opt):
        kind, names = "adagrad", [""]

        def __init__(self, lr, var_id_list, env, namespace_id):
            self.h, self.ns = (lr,), namespace_id
            for vid in var_id_list:
                env.namespace(namespace_id).slot().name("%d" % vid).set(np.zeros_like(env.arrays[vid]))

        @staticmethod
        def default(namespace_id, var_id_list, env):
            return optimizers.AdaGrad(0.01, var_id_list, env, namespace_id)

    @staticmethod
    def grad_helper(losses, namespace):
        g = losses[0].graph
        xs = [g.variable(v) for v in namespace.current_var_ids()]
        gs = compute_gradients([sum_all(l) for l in losses], xs)
        pairs = [(x, gx) for x, gx in zip(xs, gs) if gx is not None]
        return [p[0] for p in pairs], [p[1] for p in pairs]
