"""Python view of the graph-level C ABI (``include/agx200.h``), shaped like the reference crate so that the re-hosted
tests read like the reference's own (``/root/reference/tests``):

    env = ag.VariableEnvironment()                       # src/variable.rs
    w = env.slot().name("w").set(rng.standard_normal((784, 10)))
    def step(g):                                         # env.run(|g| ...)
        x = g.placeholder("x", [-1, 784]); wt = g.variable(w)
        loss = T.reduce_mean(T.sparse_softmax_cross_entropy(T.matmul(x, wt), y), [0], False)
        grads = T.grad([loss], [wt])
        adam.update([wt], grads, g, ag.Feeder().push(x, batch))
    env.run(step)

Everything numeric happens in ``libagb200.so`` (C++ engine + sm_100a kernels); this file only marshals arguments.
There is no CPU fallback: creating a ``VariableEnvironment`` without a B200 raises ``OpError`` (CudaError).
"""
import ctypes as C

import numpy as np

from . import ffi

_P, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
_pi, _pi64, _pf, _pd = C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_double)


class AgxFeed(C.Structure):
    _fields_ = [("name", C.c_char_p), ("tensor_id", _i), ("data", _P), ("shape", _pi64), ("rank", _i), ("on_device", _i)]


_pfeed = C.POINTER(AgxFeed)
SIGNATURES = {
    "agx_env_new": [_i, C.POINTER(_P)], "agx_env_free": [_P], "agx_env_ctx": [_P, C.POINTER(_P)],
    "agx_env_set": [_P, C.c_char_p, C.c_char_p, _P, _pi64, _i, _pi], "agx_env_find": [_P, C.c_char_p, C.c_char_p, _pi],
    "agx_env_var_count": [_P, _pi], "agx_env_var_ids": [_P, C.c_char_p, _pi, _i, _pi], "agx_env_var_shape": [_P, _i, _pi64, _pi],
    "agx_env_get": [_P, _i, _P, _i64], "agx_env_put": [_P, _i, _P, _i64], "agx_env_var_ptr": [_P, _i, C.POINTER(_P)],
    "agx_env_save": [_P, C.c_char_p], "agx_env_load": [_P, C.c_char_p], "agx_env_set_data_parallel": [_P, _i, _i, _P], "agx_env_set_fusion": [_P, _i], "agx_env_set_plan_cache": [_P, _i], "agx_env_plan_stats": [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int)], "agx_fuse_selftest": [_i, C.c_uint, C.POINTER(C.c_int)],
    "agx_out_append": [_P, _P, _P, _i], "agx_out_error": [_P, C.c_char_p], "agx_custom_op": [_P, C.c_char_p, _P, _i, _P, _P, _P, C.POINTER(_i)],
    "agx_hook": [_P, _i, _i, C.c_char_p, _P, _P, C.POINTER(_i)],
    "agx_graph_new": [_P, C.POINTER(_P)], "agx_graph_free": [_P], "agx_graph_clear": [_P], "agx_graph_size": [_P, _pi],
    "agx_placeholder": [_P, C.c_char_p, _pi64, _i, _pi], "agx_variable": [_P, _i, _pi], "agx_variable_by_name": [_P, C.c_char_p, C.c_char_p, _pi],
    "agx_convert_to_tensor": [_P, _P, _pi64, _i, _pi],
    "agx_call": [_P, C.c_char_p, _pi, _i, _pi64, _i, _pd, _i, _pi, _i, _pi],
    "agx_grad": [_P, _pi, _i, _pi, _i, _pi, _pi], "agx_grad_helper": [_P, _pi, _i, C.c_char_p, _pi, _pi, _i, _pi],
    "agx_tensor_op_name": [_P, _i, C.c_char_p, _i], "agx_tensor_variable_id": [_P, _i, _pi],
    "agx_eval": [_P, _pi, _i, _pfeed, _i, C.POINTER(_P)], "agx_run": [_P, _pi, _i, _pfeed, _i],
    "agx_eval_launch": [_P, _pi, _i, _pfeed, _i, C.POINTER(_P)], "agx_results_fetch": [_P],
    "agx_step_capture": [_P, _pi, _i, _pfeed, _i, C.POINTER(_P)], "agx_step_launch": [_P], "agx_step_free": [_P],
    "agx_results_count": [_P, _pi], "agx_results_status": [_P, _i, _pi, C.POINTER(C.c_char_p)], "agx_results_shape": [_P, _i, _pi64, _pi],
    "agx_results_data": [_P, _i, C.POINTER(_pf), _pi64], "agx_results_free": [_P],
    "agx_opt_adam": [_P, _pi, _i, C.c_char_p, _f, _f, _f, _f, C.POINTER(_P)], "agx_opt_sgd": [_f, C.POINTER(_P)],
    "agx_opt_momentum_sgd": [_P, _pi, _i, C.c_char_p, _f, _f, C.POINTER(_P)], "agx_opt_adagrad": [_P, _pi, _i, C.c_char_p, _f, C.POINTER(_P)],
    "agx_opt_compute_updates": [_P, _P, _pi, _pi, _i, _pi], "agx_opt_get_update_op": [_P, _P, _pi, _pi, _i, _pi],
    "agx_opt_update": [_P, _P, _pi, _pi, _i, _pfeed, _i], "agx_opt_free": [_P],
}
_lib = None


class Panic(RuntimeError):
    """A condition on which the reference panics (status AGX_ERR_PANIC = 200)."""


def lib():
    global _lib
    if _lib is None:
        l = ffi.load_library()
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes, fn.restype = args, C.c_int
        l.agx_last_error.argtypes, l.agx_last_error.restype = [], C.c_char_p
        _lib = l
    return _lib


def _check(status):
    if status == 0:
        return
    msg = lib().agx_last_error().decode("utf-8", "replace")
    if status == 200:
        raise Panic(msg)
    raise ffi.OpError(status, msg)


def _f32c(value):
    """float32, C-contiguous, rank preserved (np.ascontiguousarray would turn a 0-d scalar into shape (1,))."""
    return np.require(np.asarray(value, dtype=np.float32), requirements=["C", "A"])


def _ints(v):
    return (C.c_int * len(v))(*[int(x) for x in v])


def _i64s(v):
    return (C.c_int64 * max(len(v), 1))(*[int(x) for x in v])


class EvalError(Exception):
    """EvalError::OpError (src/lib.rs:241-244)"""

    def __init__(self, code, msg):
        self.code, self.kind = code, ffi.OpError.NAMES.get(code, "Error%d" % code)
        super().__init__("%s: %s" % (self.kind, msg))


# ------------------------------------------------------------------------------------------------ Tensor
class Tensor:
    """Copy handle {id, graph} (src/tensor.rs:22-30) with the operator overloads of src/tensor.rs:817-912."""
    __array_priority__ = 1000

    def __init__(self, graph, tid):
        self.graph, self.id = graph, int(tid)

    def _coerce(self, other):
        return other if isinstance(other, Tensor) else scalar(float(other), self.graph)

    def __add__(self, o): return add(self, self._coerce(o))
    def __radd__(self, o): return add(self._coerce(o), self)
    def __sub__(self, o): return sub(self, self._coerce(o))
    def __rsub__(self, o): return sub(self._coerce(o), self)
    def __mul__(self, o): return mul(self, self._coerce(o))
    def __rmul__(self, o): return mul(self._coerce(o), self)
    def __truediv__(self, o): return div(self, self._coerce(o))
    def __rtruediv__(self, o): return div(self._coerce(o), self)
    def __neg__(self): return neg(self)

    def eval(self, ctx=None, feeds=None):
        """Tensor::eval (src/tensor.rs:94-99): Result -> value or raises EvalError."""
        return (ctx or self.graph).evaluator().push(self).feeds(feeds).run()[0].unwrap()

    # hooks (src/tensor.rs:198-320 -> hook_ops.rs:5-31): identity nodes that show the HOST value when evaluated (explicit D2H sync point)
    def _hook(self, kind, text=None, fn=None):
        cb = None
        if fn is not None:
            def tramp(_user, arr):
                fn(_host_array(arr.contents))
            cb = _HOOK_FN(tramp)
            self.graph._keep.append(cb)
        t = C.c_int()
        _check(lib().agx_hook(self.graph.h, self.id, kind, text.encode() if text else None, C.cast(cb, C.c_void_p) if cb else None, None, C.byref(t)))
        return Tensor(self.graph, t.value)

    def map(self, f): return map(self, f)
    def raw_hook(self, fn): return self._hook(0, fn=fn)
    def show(self): return self._hook(1)
    def show_shape(self): return self._hook(2)
    def print(self, what): return self._hook(3, text=str(what))

    def op_name(self):
        buf = C.create_string_buffer(256)
        _check(lib().agx_tensor_op_name(self.graph.h, self.id, buf, 256))
        return buf.value.decode()

    def get_variable_id(self):
        v = C.c_int()
        _check(lib().agx_tensor_variable_id(self.graph.h, self.id, C.byref(v)))
        return v.value if v.value >= 0 else None

    # method aliases (src/tensor_ops/mod.rs:2992-3079)
    def reshape(self, shape): return reshape(self, shape)
    def flatten(self): return flatten(self)
    def squeeze(self, axes): return squeeze(self, axes)
    def expand_dims(self, axes): return expand_dims(self, axes)
    def transpose(self, axes): return transpose(self, axes)
    def size(self): return size(self)
    def rank(self): return rank(self)
    def shape(self): return shape(self)
    def reduce_sum(self, axes, keep_dims): return reduce_sum(self, axes, keep_dims)
    def reduce_mean(self, axes, keep_dims): return reduce_mean(self, axes, keep_dims)
    def reduce_prod(self, axes, keep_dims): return reduce_prod(self, axes, keep_dims)
    def reduce_min(self, axes, keep_dims): return reduce_min(self, axes, keep_dims)
    def reduce_max(self, axes, keep_dims): return reduce_max(self, axes, keep_dims)
    def access_elem(self, i): return _call(self.graph, "access_elem", [self], [i])


class Result:
    """Result<NdArray, EvalError> of one evaluation target."""

    def __init__(self, value, code, msg):
        self.value, self.code, self.msg = value, code, msg

    def is_ok(self):
        return self.code == 0

    def unwrap(self):
        if self.code != 0:
            raise EvalError(self.code, self.msg)
        return self.value


# ------------------------------------------------------------------------------------------------ feeds / evaluator
class Feeder:
    """src/evaluation.rs:91-114"""

    def __init__(self):
        self.items = []

    def push(self, key, value):
        self.items.append((key, value))
        return self


class DeviceArray:
    """A feed value that already lives in HBM (bench `value` leg): device pointer + shape."""

    def __init__(self, ptr, shape):
        self.ptr, self.shape = int(ptr), tuple(int(s) for s in shape)


class HostPrefetcher:
    """Double-buffered host -> HBM feed staging (the device-side form of Feeder::push with host arrays, evaluation.rs:296).

        pf = HostPrefetcher(env, [x_shape, y_shape]); pf.stage([x0, y0])
        for i in range(steps):
            x, y = pf.acquire()                                   # DeviceArrays of step i
            pending = ev.feed("x", x).feed("y", y).run_deferred()   # kernels + result copies enqueued, no host sync
            pf.stage([x_next, y_next])                              # H2D of step i+1 runs under the kernels of step i
            ... previous.get() ...; previous = pending              # read step i-1's loss while step i runs

    stage() copies from (ideally pinned) host arrays on a second CUDA stream, so the transfer of step i+1 runs under the kernels of
    step i; acquire() makes the compute stream wait for it.  Values and results are identical to feeding the host arrays directly."""

    def __init__(self, env, shapes):
        from . import ffi as _ffi
        self._ffi, self._lib, self._ctx = _ffi, _ffi.load_library(), env.agb_ctx()
        self.shapes = [tuple(int(d) for d in s) for s in shapes]
        self.bufs = []
        for _ in range(2):
            row = []
            for s in self.shapes:
                p = C.c_void_p()
                _ffi.check(self._lib.agb_alloc(self._ctx, max(int(np.prod(s)), 1) * 4, C.byref(p)))
                row.append(DeviceArray(p.value, s))
            self.bufs.append(row)
        self.staged, self.cur = 0, 0        # buffer index the next stage() writes / the last acquire() returned
        self._keep = [None, None]           # host arrays of the in-flight copies (must outlive the async transfer)

    def stage(self, host_arrays):
        keep = []
        for d, a in zip(self.bufs[self.staged], host_arrays):
            a = _f32c(a)
            assert tuple(a.shape) == d.shape, "staged array must have the placeholder's shape"
            self._ffi.check(self._lib.agb_stage_h2d(self._ctx, d.ptr, a.ctypes.data, a.nbytes))
            keep.append(a)
        self._keep[self.staged] = keep
        self.cur, self.staged = self.staged, self.staged ^ 1

    def acquire(self):
        self._ffi.check(self._lib.agb_stage_wait(self._ctx))     # compute waits for the staged copy
        self._ffi.check(self._lib.agb_stage_mark(self._ctx))     # everything enqueued before this step may still read the other buffer
        return list(self.bufs[self.cur])

    def close(self):
        for row in self.bufs:
            for d in row:
                self._lib.agb_free(self._ctx, d.ptr)
        self.bufs = []


def _make_feeds(items):
    keep, arr = [], (AgxFeed * max(len(items), 1))()
    for k, (key, value) in enumerate(items):
        f = arr[k]
        if isinstance(key, Tensor):
            f.name, f.tensor_id = None, key.id
        else:
            f.name, f.tensor_id = str(key).encode(), -1
        if isinstance(value, DeviceArray):
            shp = _i64s(value.shape)
            f.data, f.shape, f.rank, f.on_device = value.ptr, shp, len(value.shape), 1
            keep.append(shp)
        else:
            a = _f32c(value)
            shp = _i64s(a.shape)
            f.data, f.shape, f.rank, f.on_device = a.ctypes.data, shp, a.ndim, 0
            keep += [a, shp]
    return arr, len(items), keep


class StepGraph:
    """A captured training / evaluation step (Evaluator.capture)."""

    def __init__(self, h):
        self.h = h

    def launch(self):
        _check(lib().agx_step_launch(self.h))

    def close(self):
        if self.h:
            lib().agx_step_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Deferred:
    """Pending results of Evaluator.run_deferred()."""

    def __init__(self, ev, res):
        self._ev, self._res, self._out = ev, res, None

    def get(self):
        if self._out is None:
            _check(lib().agx_results_fetch(self._res))
            self._out = self._ev._collect(self._res)
            self._res = None
        return self._out

    def __del__(self):
        try:
            if self._res is not None:
                lib().agx_results_free(self._res)
        except Exception:
            pass


class Evaluator:
    """src/evaluation.rs:58-172"""

    def __init__(self, graph):
        self.graph, self.targets, self.feeder = graph, [], Feeder()

    def push(self, x):
        self.targets.append(x)
        return self

    def extend(self, xs):
        self.targets.extend(xs)
        return self

    def feed(self, key, value):
        self.feeder.push(key, value)
        return self

    def feeds(self, feeds):
        if isinstance(feeds, Feeder):
            self.feeder = feeds
        elif feeds:
            for k, v in (feeds.items() if isinstance(feeds, dict) else feeds):
                self.feeder.push(k, v)
        return self

    def set_feeder(self, feeder):
        self.feeder = feeder
        return self

    def _collect(self, res):
        out = []
        try:
            for i in range(len(self.targets)):
                code, msg = C.c_int(), C.c_char_p()
                lib().agx_results_status(res, i, C.byref(code), C.byref(msg))
                if code.value != 0:
                    out.append(Result(None, code.value, (msg.value or b"").decode("utf-8", "replace")))
                    continue
                shp, rank = (C.c_int64 * 8)(), C.c_int()
                lib().agx_results_shape(res, i, shp, C.byref(rank))
                data, cnt = _pf(), C.c_int64()
                _check(lib().agx_results_data(res, i, C.byref(data), C.byref(cnt)))
                a = np.ctypeslib.as_array(data, shape=(cnt.value,)).copy() if cnt.value else np.zeros((0,), np.float32)
                out.append(Result(a.reshape(tuple(shp[k] for k in range(rank.value))), 0, ""))
        finally:
            lib().agx_results_free(res)
        return out

    def run(self):
        arr, n, keep = _make_feeds(self.feeder.items)
        res = C.c_void_p()
        _check(lib().agx_eval(self.graph.h, _ints([t.id for t in self.targets]), len(self.targets), arr, n, C.byref(res)))
        return self._collect(res)

    def run_deferred(self):
        """Launch the evaluation and queue the results' device -> host copies on the copy stream WITHOUT a host sync; returns a
        Deferred whose .get() yields what run() would have returned.  Calling .get() after the NEXT step has been launched keeps the
        GPU busy while the host reads the loss (the reference's evaluator is synchronous: evaluation.rs:150-172)."""
        arr, n, keep = _make_feeds(self.feeder.items)
        res = C.c_void_p()
        _check(lib().agx_eval_launch(self.graph.h, _ints([t.id for t in self.targets]), len(self.targets), arr, n, C.byref(res)))
        return Deferred(self, res)

    def capture(self):
        """Capture this evaluation (run_async semantics, device-resident feeds only) into a CUDA graph; returns a StepGraph whose
        .launch() replays the whole step — forward, backward, optimizer — without walking the graph on the host."""
        arr, n, keep = _make_feeds(self.feeder.items)
        h = C.c_void_p()
        _check(lib().agx_step_capture(self.graph.h, _ints([t.id for t in self.targets]), len(self.targets), arr, n, C.byref(h)))
        return StepGraph(h)

    def run_async(self):
        """Evaluate for side effects only (training step): nothing is copied back, no host sync."""
        arr, n, keep = _make_feeds(self.feeder.items)
        _check(lib().agx_run(self.graph.h, _ints([t.id for t in self.targets]), len(self.targets), arr, n))


# ------------------------------------------------------------------------------------------------ graph / context
class Context:
    """Context / Graph (src/graph.rs:109-200)."""

    def __init__(self, env):
        self.env = env
        h = C.c_void_p()
        _check(lib().agx_graph_new(env.h, C.byref(h)))
        self.h = h
        self._keep = []        # ctypes callbacks / Python ops of user-defined nodes and hooks: must outlive the graph

    def close(self):
        if self.h:
            lib().agx_graph_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.env.h:
                self.close()
        except Exception:
            pass

    def clear(self):
        _check(lib().agx_graph_clear(self.h))

    def size(self):
        n = C.c_int()
        _check(lib().agx_graph_size(self.h, C.byref(n)))
        return n.value

    def placeholder(self, name, shape):
        t = C.c_int()
        _check(lib().agx_placeholder(self.h, name.encode(), _i64s(shape), len(shape), C.byref(t)))
        return Tensor(self, t.value)

    def variable(self, key):
        """GetVariableTensor (src/variable.rs:117-148): VariableID | "name" | ("namespace", "name")"""
        t = C.c_int()
        if isinstance(key, (int, np.integer)):
            _check(lib().agx_variable(self.h, int(key), C.byref(t)))
        elif isinstance(key, tuple):
            _check(lib().agx_variable_by_name(self.h, key[0].encode(), key[1].encode(), C.byref(t)))
        else:
            _check(lib().agx_variable_by_name(self.h, b"", str(key).encode(), C.byref(t)))
        return Tensor(self, t.value)

    def evaluator(self):
        return Evaluator(self)

    def namespace(self, ns):
        return self.env.namespace(ns)

    def default_namespace(self):
        return self.env.namespace("")


class _Slot:
    def __init__(self, env, ns, name=None):
        self.env, self.ns, self._name = env, ns, name

    def name(self, name):
        return _Slot(self.env, self.ns, name)

    def set(self, value):
        import uuid
        a = _f32c(value)
        v = C.c_int()
        nm = self._name if self._name is not None else str(uuid.uuid4())      # DefaultVariableSlot::set (variable.rs:262-270)
        _check(lib().agx_env_set(self.env.h, self.ns.encode(), nm.encode(), a.ctypes.data, _i64s(a.shape), a.ndim, C.byref(v)))
        return v.value


class Namespace:
    """VariableNamespace(Mut) (src/variable.rs:169-215)"""

    def __init__(self, env, ns):
        self.env, self.ns = env, ns

    def slot(self):
        return _Slot(self.env, self.ns)

    def current_var_ids(self):
        out, n = (C.c_int * 4096)(), C.c_int()
        _check(lib().agx_env_var_ids(self.env.h, self.ns.encode(), out, 4096, C.byref(n)))
        return [out[i] for i in range(n.value)]

    def get_array_by_name(self, name):
        v = C.c_int()
        _check(lib().agx_env_find(self.env.h, self.ns.encode(), name.encode(), C.byref(v)))
        return None if v.value < 0 else self.env.get_array_by_id(v.value)


class VariableEnvironment:
    """src/variable.rs:152-155 — variables live in HBM; host copies only on get/save."""

    def __init__(self, device=0):
        h = C.c_void_p()
        _check(lib().agx_env_new(int(device), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            h, self.h = self.h, None
            lib().agx_env_free(h)

    def slot(self):
        return _Slot(self, "")

    def namespace(self, ns):
        return Namespace(self, ns)

    namespace_mut = namespace

    def default_namespace(self):
        return Namespace(self, "")

    default_namespace_mut = default_namespace

    def get_array_by_id(self, vid):
        shp, rank = (C.c_int64 * 8)(), C.c_int()
        _check(lib().agx_env_var_shape(self.h, vid, shp, C.byref(rank)))
        shape = tuple(shp[i] for i in range(rank.value))
        out = np.empty(shape, np.float32)
        _check(lib().agx_env_get(self.h, vid, out.ctypes.data, out.size))
        return out

    def set_array_by_id(self, vid, value):
        a = _f32c(value)
        _check(lib().agx_env_put(self.h, vid, a.ctypes.data, a.size))

    def var_ptr(self, vid):
        p = C.c_void_p()
        _check(lib().agx_env_var_ptr(self.h, vid, C.byref(p)))
        return p.value

    def agb_ctx(self):
        p = C.c_void_p()
        _check(lib().agx_env_ctx(self.h, C.byref(p)))
        return p

    def save(self, path):
        _check(lib().agx_env_save(self.h, str(path).encode()))

    def initialize(self, path):
        _check(lib().agx_env_load(self.h, str(path).encode()))

    @staticmethod
    def load(path, device=0):
        env = VariableEnvironment(device)
        env.initialize(path)
        return env

    def set_data_parallel(self, rank, world, nccl_id):
        buf = C.create_string_buffer(bytes(nccl_id), 128)
        _check(lib().agx_env_set_data_parallel(self.h, rank, world, buf))

    def set_fusion(self, on):
        """Deferred elementwise expressions (engine/fuse.cc): on by default; off = one launch per node (elementwise values bit-identical; summed
        weight-gradient GEMMs differ by fp32 reassociation)."""
        _check(lib().agx_env_set_fusion(self.h, 1 if on else 0))

    def set_plan_cache(self, on):
        """Automatic step-plan cache (capi.cc): an evaluation seen for the third time with the same graph structure, targets and feed shapes
        replays a CUDA graph captured at its second sight — also across graph objects rebuilt per step (env.run).  On by default."""
        _check(lib().agx_env_set_plan_cache(self.h, 1 if on else 0))

    def plan_stats(self):
        c, r, n = C.c_int64(), C.c_int64(), C.c_int()
        _check(lib().agx_env_plan_stats(self.h, C.byref(c), C.byref(r), C.byref(n)))
        return {"captures": c.value, "replays": r.value, "live_plans": n.value}

    def run(self, f):
        """env.run(|g| ...) (src/variable.rs:670-683): a fresh graph per call."""
        g = Context(self)
        try:
            return f(g)
        finally:
            g.close()


def run(f, device=0):
    """ag::run (src/graph.rs:86-100): a throw-away environment."""
    env = VariableEnvironment(device)
    try:
        return env.run(f)
    finally:
        env.close()


# ------------------------------------------------------------------------------------------------ tensor_ops
def _graph_of(tensors):
    for t in tensors:
        if isinstance(t, Tensor):
            return t.graph
    raise Panic("no tensor argument to take the graph from")


# ---- user-defined ops (the public `Op` trait, src/op.rs:1-48,90-101; tests/test_core.rs:6-36) through the host-callback ABI
class _HostArray(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("shape", C.POINTER(C.c_int64)), ("rank", C.c_int)]


_COMPUTE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(_HostArray), C.c_int, C.c_void_p)
_GRAD_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int))
_HOOK_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(_HostArray))


def _host_array(h):
    shape = tuple(h.shape[i] for i in range(h.rank))
    n = int(np.prod(shape)) if shape else 1
    return np.ctypeslib.as_array(h.data, shape=(n,)).reshape(shape).copy() if n else np.zeros(shape, np.float32)


class OpError(Exception):
    """Raise from Op.compute to return Err(OpError::..) (op.rs:67-73): code 1..5."""

    def __init__(self, code, msg):
        self.code, self.msg = int(code), str(msg)
        super().__init__(msg)


class ComputeContext:
    """op::ComputeContext (src/op.rs:186-309) for a user-defined op: inputs arrive as host ndarrays."""

    def __init__(self, inputs, out):
        self._inputs, self._out = inputs, out

    def num_inputs(self): return len(self._inputs)
    def input(self, i): return self._inputs[i]

    def append_output(self, arr):
        a = _f32c(arr)
        _check(lib().agx_out_append(self._out, a.ctypes.data, _i64s(a.shape), a.ndim))


class GradientContext:
    """op::GradientContext (src/op.rs:342-434)."""

    def __init__(self, g, inputs, y, gy):
        self.g, self._inputs, self._y, self._gy, self.gxs = g, inputs, y, gy, []

    def graph(self): return self.g
    def num_inputs(self): return len(self._inputs)
    def input(self, i): return self._inputs[i]
    def output(self): return self._y
    def output_grad(self): return self._gy
    def append_input_grad(self, gx): self.gxs.append(gx)


class Op:
    """trait Op (src/op.rs:90-101): subclass with name() / compute(ctx) / grad(ctx); build nodes with `build_op`."""

    def name(self): return type(self).__name__
    def compute(self, ctx): raise NotImplementedError
    def grad(self, ctx):
        for _ in range(ctx.num_inputs()):
            ctx.append_input_grad(None)


def build_op(g, op, inputs=()):
    """Tensor::builder(g).append_input(x, false)...build(op) (src/tensor.rs:609-803) for a Python `Op`."""
    inputs = list(inputs)

    def compute(_user, arrs, n, out):
        try:
            op.compute(ComputeContext([_host_array(arrs[i]) for i in range(n)], out))
            return 0
        except OpError as e:
            lib().agx_out_error(out, e.msg.encode())
            return e.code
        except Exception as e:          # a panic must not cross the FFI boundary: reported as NdArrayError
            lib().agx_out_error(out, ("%s: %s" % (type(e).__name__, e)).encode())
            return 1

    def grad(_user, _g, ins, n, y, gy, gxs):
        ctx = GradientContext(g, [Tensor(g, ins[i]) for i in range(n)], Tensor(g, y), Tensor(g, gy))
        op.grad(ctx)
        for i in range(n):
            gx = ctx.gxs[i] if i < len(ctx.gxs) else None
            gxs[i] = gx.id if gx is not None else -1
    c_fn, g_fn = _COMPUTE_FN(compute), _GRAD_FN(grad)
    g._keep.extend([c_fn, g_fn, op])
    t = C.c_int()
    _check(lib().agx_custom_op(g.h, op.name().encode(), _ints([x.id for x in inputs]), len(inputs), C.cast(c_fn, C.c_void_p), C.cast(g_fn, C.c_void_p), None, C.byref(t)))
    return Tensor(g, t.value)


class _MapOp(Op):
    """higher_order_ops.rs:5-36 MapOp: y = f(x) on the host value, no gradient."""

    def __init__(self, f): self.f = f
    def name(self): return "MapOp"
    def compute(self, ctx): ctx.append_output(self.f(ctx.input(0)))


def map(x, f):  # noqa: A001  (the reference's name: tensor_ops::map, mod.rs:2931-2945)
    return build_op(x.graph, _MapOp(f), [x])


def _call(g, fn, tensors=(), ints=(), floats=(), multi=False):
    for t in tensors:
        if t.graph is not g:
            raise Panic("Detected tensors belonging to different graphs")
    out, n = (C.c_int * 64)(), C.c_int()
    fl = (C.c_double * max(len(floats), 1))(*[float(x) for x in floats])
    _check(lib().agx_call(g.h, fn.encode(), _ints([t.id for t in tensors]), len(tensors), _i64s(ints), len(ints), fl, len(floats), out, 64, C.byref(n)))
    res = [Tensor(g, out[i]) for i in range(n.value)]
    return res if multi else res[0]


def as_tensor(v, g):
    """AsTensor (src/tensor.rs:915-940): a Tensor passes through; an integer list becomes a (host-side) constant."""
    if isinstance(v, Tensor):
        return v
    return _call(g, "as_tensor", [], [int(x) for x in v])


def convert_to_tensor(arr, g):
    a = _f32c(arr)
    t = C.c_int()
    _check(lib().agx_convert_to_tensor(g.h, a.ctypes.data, _i64s(a.shape), a.ndim, C.byref(t)))
    return Tensor(g, t.value)


def scalar(v, g): return _call(g, "scalar", [], [], [v])
def zeros(shape, g): return _call(g, "zeros", [as_tensor(shape, g)])
def ones(shape, g): return _call(g, "ones", [as_tensor(shape, g)])


def _u(name):
    def f(x):
        return _call(x.graph, name, [x])
    f.__name__ = name
    return f


for _n in ["sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "exp2", "exp10", "ln", "log2",
           "log10", "sqrt", "neg", "abs", "sign", "floor", "ceil", "inv", "inv_sqrt", "square", "sigmoid", "relu", "softplus", "lgamma", "digamma", "shape", "rank",
           "size", "identity", "stop_gradient", "sum_all", "mean_all", "flatten"]:
    globals()[_n] = _u(_n)
# the reference exports the gamma functions per float type (mod.rs:488-538); the device arithmetic is f32 for both spellings
lgamma_f32 = lgamma_f64 = globals()["lgamma"]
digamma_f32 = digamma_f64 = globals()["digamma"]


def _b(name):
    def f(a, b):
        g = _graph_of([a, b])
        a = a if isinstance(a, Tensor) else scalar(float(a), g)
        b = b if isinstance(b, Tensor) else scalar(float(b), g)
        return _call(g, name, [a, b])
    f.__name__ = name
    return f


for _n in ["add", "sub", "mul", "div", "equal", "not_equal", "greater", "lesser", "greater_equal", "lesser_equal", "maximum", "minimum",
           "sigmoid_cross_entropy", "softmax_cross_entropy", "sparse_softmax_cross_entropy", "mean_squared_error", "matmul", "batch_matmul",
           "setdiff1d", "assign"]:
    globals()[_n] = _b(_n)


def pow(x, a): return _call(x.graph, "pow", [x], [], [a])
def elu(x, alpha): return _call(x.graph, "elu", [x], [], [alpha])
def leaky_relu(x, alpha): return _call(x.graph, "leaky_relu", [x], [], [alpha])
def clip(x, lo, hi): return _call(x.graph, "clip", [x], [], [lo, hi])
def add_n(xs): return _call(xs[0].graph, "add_n", list(xs))
def nth_tensor(x, n): return _call(x.graph, "nth_tensor", [x], [n])
def _red(name):
    def f(x, axes, keep_dims):
        return _call(x.graph, name, [x, as_tensor(axes, x.graph)], [int(keep_dims)])
    f.__name__ = name
    return f


for _n in ["reduce_sum", "reduce_mean", "reduce_prod", "reduce_min", "reduce_max", "reduce_variance"]:
    globals()[_n] = _red(_n)


def argmax(x, axis, keep_dim): return _call(x.graph, "argmax", [x], [axis, int(keep_dim)])
def argmin(x, axis, keep_dim): return _call(x.graph, "argmin", [x], [axis, int(keep_dim)])
def reduce_logsumexp(x, axis, keep_dim): return _call(x.graph, "reduce_logsumexp", [x], [axis, int(keep_dim)])
def softmax(x, axis): return _call(x.graph, "softmax", [x], [axis])
def log_softmax(x, axis): return _call(x.graph, "log_softmax", [x], [axis])
def batch_matmul_t(a, b, trans_a, trans_b): return _call(a.graph, "batch_matmul_t", [a, b], [int(trans_a), int(trans_b)])
def tensordot(a, b, a_axes, b_axes): return _call(a.graph, "tensordot", [a, b, as_tensor(a_axes, a.graph), as_tensor(b_axes, a.graph)])
def reshape(x, shape): return _call(x.graph, "reshape", [x, as_tensor(shape, x.graph)])
def transpose(x, axes): return _call(x.graph, "transpose", [x, as_tensor(axes, x.graph)])
def squeeze(x, axes): return _call(x.graph, "squeeze", [x, as_tensor(axes, x.graph)])
def expand_dims(x, axes): return _call(x.graph, "expand_dims", [x, as_tensor(axes, x.graph)])
def slice(x, starts, ends): return _call(x.graph, "slice", [x], list(starts) + list(ends))
def split(x, sizes, axis): return _call(x.graph, "split", [x], [axis] + list(sizes), multi=True)
def concat(xs, axis): return _call(xs[0].graph, "concat", list(xs), [axis])
def tile(x, axis, num): return _call(x.graph, "tile", [x], [axis, num])
def gather_common(param, indices, axis): return _call(param.graph, "gather_common", [param, as_tensor(indices, param.graph)], [axis])
def gather(param, indices, axis): return _call(param.graph, "gather", [param, as_tensor(indices, param.graph)], [axis])
def conv2d(x, w, pad, stride): return _call(x.graph, "conv2d", [x, w], [pad, stride])
def dilated_conv2d(x, w, pad, stride, dilate): return _call(x.graph, "dilated_conv2d", [x, w], [pad, stride, dilate])
def conv2d_transpose(x, w, pad, stride): return _call(x.graph, "conv2d_transpose", [x, w], [pad, stride])
def dilated_conv2d_transpose(x, w, pad, stride, dilate): return _call(x.graph, "dilated_conv2d_transpose", [x, w], [pad, stride, dilate])
def max_pool2d(x, pool_size, pad, stride): return _call(x.graph, "max_pool2d", [x], [pool_size, pad, stride])
def dropout(x, dropout_ratio, train, seed=0): return _call(x.graph, "dropout", [x], [int(train), int(seed)], [dropout_ratio])


# random_* generator ops (src/tensor_ops/mod.rs:2426-2676).  `seed=0` = the crate's default rng (fixed seed, ndarray_ext.rs:250-264); the
# `_rng` variants of the reference take an ArrayRng, here a seed.  Values come from a device Philox stream (parity-unpinned).
def _rand(kind, shape, g, p0, p1, seed): return _call(g, "random", [as_tensor(shape, g)], [kind, int(seed)], [float(p0), float(p1)])
def random_uniform(shape, min, max, g, seed=0): return _rand(0, shape, g, min, max, seed)
def random_normal(shape, mean, stddev, g, seed=0): return _rand(1, shape, g, mean, stddev, seed)
def standard_uniform(shape, g, seed=0): return _rand(0, shape, g, 0.0, 1.0, seed)
def standard_normal(shape, g, seed=0): return _rand(1, shape, g, 0.0, 1.0, seed)
def bernoulli(shape, p, g, seed=0): return _rand(2, shape, g, p, 0.0, seed)
def random_exp(shape, lambda_, g, seed=0): return _rand(3, shape, g, lambda_, 0.0, seed)
def log_normal(shape, mean, stddev, g, seed=0): return _rand(4, shape, g, mean, stddev, seed)
def gamma(shape, shape_param, scale, g, seed=0): return _rand(5, shape, g, shape_param, scale, seed)


random_gamma = gamma                         # the reference's name (mod.rs:2639)


class ArrayRng:
    """ndarray_ext::ArrayRng (ndarray_ext.rs:243-264) as far as the graph API needs it: a seed for the `_rng` constructors.  The default is
    the crate's fixed default seed; the values come from the device Philox stream (parity-unpinned)."""

    def __init__(self, seed=0):
        self.seed = int(seed)

    @staticmethod
    def default():
        return ArrayRng(0)


def _seed_of(arr_rng): return arr_rng.seed if isinstance(arr_rng, ArrayRng) else int(arr_rng)
def random_uniform_rng(arr_rng, shape, min, max, g): return random_uniform(shape, min, max, g, _seed_of(arr_rng))
def random_normal_rng(arr_rng, shape, mean, stddev, g): return random_normal(shape, mean, stddev, g, _seed_of(arr_rng))
def standard_uniform_rng(arr_rng, shape, g): return standard_uniform(shape, g, _seed_of(arr_rng))
def standard_normal_rng(arr_rng, shape, g): return standard_normal(shape, g, _seed_of(arr_rng))
def bernoulli_rng(arr_rng, shape, p, g): return bernoulli(shape, p, g, _seed_of(arr_rng))
def random_exp_rng(arr_rng, shape, lambda_, g): return random_exp(shape, lambda_, g, _seed_of(arr_rng))
def log_normal_rng(arr_rng, shape, mean, stddev, g): return log_normal(shape, mean, stddev, g, _seed_of(arr_rng))
def random_gamma_rng(arr_rng, shape, shape_param, scale, g): return gamma(shape, shape_param, scale, g, _seed_of(arr_rng))
def dropout_rng(x, dropout_ratio, train, rng): return dropout(x, dropout_ratio, train, _seed_of(rng) or 0x5EED + 1)


def normalize(x, axes): return _call(x.graph, "normalize", [x, as_tensor(axes, x.graph)])
def batch_norm(x, scale, shift): return _call(x.graph, "batch_norm", [x, scale, shift])
def control_dependencies(x, deps): return _call(x.graph, "control_dependencies", [x] + list(deps))


def _hessian_vector_product(ys, xs, vectors):
    """(Experimental in the reference) mod.rs:218-236: grad of sum_i grad(ys, xs)[i] * vectors[i] with respect to xs."""
    return grad([gx * v for gx, v in zip(grad(ys, xs), vectors)], xs)


def grad(ys, xs):
    """T::grad (src/tensor_ops/mod.rs:94-114)"""
    g = ys[0].graph
    out = (C.c_int * len(xs))()
    _check(lib().agx_grad(g.h, _ints([y.id for y in ys]), len(ys), _ints([x.id for x in xs]), len(xs), None, out))
    return [Tensor(g, out[i]) for i in range(len(xs))]


def grad_with_default(ys, xs, ys_grads):
    g = ys[0].graph
    out = (C.c_int * len(xs))()
    _check(lib().agx_grad(g.h, _ints([y.id for y in ys]), len(ys), _ints([x.id for x in xs]), len(xs), _ints([t.id for t in ys_grads]), out))
    return [Tensor(g, out[i]) for i in range(len(xs))]


def jacobians(y, xs, objective_len):
    """src/tensor_ops/mod.rs:188-217"""
    vec_vec = [grad([y.access_elem(i)], xs) for i in range(objective_len)]
    return [concat([expand_dims(flatten(v[i]), [0]) for v in vec_vec], 0) for i in range(len(xs))]


# ------------------------------------------------------------------------------------------------ optimizers
class _Optimizer:
    """trait Optimizer (src/optimizers/mod.rs:49-99)"""

    def __init__(self, h):
        self.h = h

    def compute_updates(self, params, grads, g):
        out = (C.c_int * len(params))()
        _check(lib().agx_opt_compute_updates(self.h, g.h, _ints([p.id for p in params]), _ints([t.id for t in grads]), len(params), out))
        return [Tensor(g, out[i]) for i in range(len(params))]

    def get_update_op(self, params, grads, g):
        t = C.c_int()
        _check(lib().agx_opt_get_update_op(self.h, g.h, _ints([p.id for p in params]), _ints([x.id for x in grads]), len(params), C.byref(t)))
        return Tensor(g, t.value)

    def update(self, params, grads, g, feeder=None):
        arr, n, keep = _make_feeds(feeder.items if feeder else [])
        _check(lib().agx_opt_update(self.h, g.h, _ints([p.id for p in params]), _ints([x.id for x in grads]), len(params), arr, n))


class optimizers:
    class Adam(_Optimizer):
        def __init__(self, alpha, eps, b1, b2, var_id_list, env, namespace_id):
            h = C.c_void_p()
            _check(lib().agx_opt_adam(env.h, _ints(var_id_list), len(var_id_list), namespace_id.encode(), alpha, eps, b1, b2, C.byref(h)))
            super().__init__(h)

        @staticmethod
        def default(namespace_id, var_id_list, env):      # src/optimizers/adam.rs:58-72
            return optimizers.Adam(0.001, 1e-08, 0.9, 0.999, var_id_list, env, namespace_id)

    class SGD(_Optimizer):
        def __init__(self, alpha):
            h = C.c_void_p()
            _check(lib().agx_opt_sgd(alpha, C.byref(h)))
            super().__init__(h)

    class MomentumSGD(_Optimizer):
        def __init__(self, alpha, momentum, var_id_list, env, namespace_id):
            h = C.c_void_p()
            _check(lib().agx_opt_momentum_sgd(env.h, _ints(var_id_list), len(var_id_list), namespace_id.encode(), alpha, momentum, C.byref(h)))
            super().__init__(h)

        @staticmethod
        def default(namespace_id, var_id_list, env):      # src/optimizers/momentum_sgd.rs:42-54
            return optimizers.MomentumSGD(0.01, 0.9, var_id_list, env, namespace_id)

    class AdaGrad(_Optimizer):
        def __init__(self, lr, var_id_list, env, namespace_id):
            h = C.c_void_p()
            _check(lib().agx_opt_adagrad(env.h, _ints(var_id_list), len(var_id_list), namespace_id.encode(), lr, C.byref(h)))
            super().__init__(h)

        @staticmethod
        def default(namespace_id, var_id_list, env):      # src/optimizers/adagrad.rs:13-30
            return optimizers.AdaGrad(0.01, var_id_list, env, namespace_id)

    @staticmethod
    def grad_helper(losses, namespace):
        """src/optimizers/mod.rs:21-46"""
        g = losses[0].graph
        vs, gs, n = (C.c_int * 4096)(), (C.c_int * 4096)(), C.c_int()
        _check(lib().agx_grad_helper(g.h, _ints([l.id for l in losses]), len(losses), namespace.ns.encode(), vs, gs, 4096, C.byref(n)))
        return [Tensor(g, vs[i]) for i in range(n.value)], [Tensor(g, gs[i]) for i in range(n.value)]
