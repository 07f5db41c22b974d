"""The reference's gradient / evaluation tests, re-hosted as backend-neutral graph builders.

Every case cites the reference test it restates (/root/reference/tests/*.rs or src/...:line).  A case is a function
``case(ag, env, g, rng) -> (z, grads, var_ids, feeds)`` that registers its variables in ``env`` and builds its graph in ``g``
using only names shared by the CUDA engine (rust_autograd_b200.autograd) and the oracle (oracle.ref_graph), so that
  * tests/test_oracle_graph.py runs it on the oracle in float64 with the reference's own finite-difference check
    (ag::test_helper::check_theoretical_grads, src/test_helper.rs:9-148: eps 1e-3, tol 1e-3 / 1e-2), and
  * tests/test_engine_gpu.py runs it on the GPU engine and compares z and every gradient with the oracle.
"""
import numpy as np


def _u(name, lo, hi, shape=(3,), tol=1e-3):
    def case(ag, env, g, rng):
        v = env.slot().set(rng.uniform(lo, hi, shape))
        vt = g.variable(v)
        z = getattr(ag, name)(vt)
        return z, ag.grad([z], [vt]), [v], None
    case.__name__ = name
    case.tol = tol
    return case


def _mk(tol=1e-3):
    def deco(f):
        f.tol = tol
        return f
    return deco


CASES = []


def case(tol=1e-3):
    def deco(f):
        f.tol = tol
        CASES.append(f)
        return f
    return deco


# tests/test_tensor_ops_grad.rs:78-413: unary math on rng.random_uniform(&[3], lo, hi)
for _n, _lo, _hi, _tol in [("asinh", 0., 0.2, 1e-3), ("acosh", 1.1, 1.3, 1e-3), ("atanh", 0., 0.2, 1e-3), ("sinh", 0., 0.2, 1e-3), ("cosh", 0., 0.2, 1e-3),
                           ("tanh", 0., 0.2, 1e-3), ("asin", 0., 0.2, 1e-2), ("acos", 0., 0.2, 1e-3), ("atan", 0., 0.2, 1e-3), ("sin", 0., 0.2, 1e-3),
                           ("cos", 0., 0.2, 1e-3), ("tan", 0., 0.2, 1e-2), ("sqrt", 0.9, 1.1, 1e-3), ("exp", 0.9, 1.1, 1e-2), ("ln", 1., 1.1, 1e-2),
                           ("abs", 0.2, 1.0, 1e-3), ("neg", -1., 1., 1e-3), ("square", -1., 1., 1e-3), ("inv", 0.5, 1.5, 1e-3), ("sigmoid", -1., 1., 1e-3),
                           ("softplus", -1., 1., 1e-3), ("exp2", 0.5, 1., 1e-2), ("exp10", 0., 0.3, 1e-2), ("log2", 1., 2., 1e-3), ("log10", 1., 2., 1e-3),
                           ("inv_sqrt", 0.5, 1.5, 1e-3), ("lgamma", 1., 1.01, 1e-3)]:     # lgamma: :858-876 (f64 there)
    CASES.append(_u(_n, _lo, _hi, tol=_tol))


@case()
def get(ag, env, g, rng):                   # :9-31 access_elem
    v = env.slot().set(np.array([1., 2., 3.]))
    vt = g.variable(v)
    z = (2. * vt).access_elem(1)
    return z, ag.grad([z], [vt]), [v], None


@case()
def add_n(ag, env, g, rng):                 # :34-55
    vs = [env.slot().set(np.array([1., 2., 3.])) for _ in range(3)]
    ts = [g.variable(v) for v in vs]
    z = ag.add_n(ts)
    return z, ag.grad([z], [ts[1]]), [vs[1]], None


@case()
def clip(ag, env, g, rng):                  # :58-75
    v = env.slot().set(np.array([1., 2., 3.]))
    vt = g.variable(v)
    z = ag.clip(vt, 1.5, 2.5)
    return z, ag.grad([z], [vt]), [v], None


@case()
def pow(ag, env, g, rng):                   # :330-348
    v = env.slot().set(rng.uniform(0.9, 1.1, (3,)))
    vt = g.variable(v)
    z = ag.pow(vt, 1.1)
    return z, ag.grad([z], [vt]), [v], None


@case()
def expand_dims(ag, env, g, rng):           # :415-433
    v = env.slot().set(rng.standard_normal((3,)))
    vt = g.variable(v)
    z = ag.expand_dims(vt, [0, 2])
    return z, ag.grad([z], [vt]), [v], None


@case()
def squeeze(ag, env, g, rng):               # :436-454
    v = env.slot().set(rng.standard_normal((3, 1, 2, 1)))
    vt = g.variable(v)
    z = ag.squeeze(vt, [3, 1])
    return z, ag.grad([z], [vt]), [v], None


@case(5e-3)
def matmul(ag, env, g, rng):                # :457-476
    v = env.slot().set(rng.standard_normal((2, 3)))
    a = ag.convert_to_tensor(rng.standard_normal((4, 2)), g)
    vt = g.variable(v)
    z = ag.matmul(a, vt)
    return z, ag.grad([z], [vt]), [v], None


@case()
def batch_matmul(ag, env, g, rng):          # :479-498
    v = env.slot().set(rng.standard_normal((2, 2, 3)))
    a = ag.convert_to_tensor(rng.standard_normal((2, 4, 2)), g)
    vt = g.variable(v)
    z = ag.batch_matmul(a, vt)
    return z, ag.grad([z], [vt]), [v], None


@case()
def implicit_broadcast(ag, env, g, rng):    # :501-520
    b = env.slot().set(rng.standard_normal((1, 3)))
    x = ag.convert_to_tensor(rng.standard_normal((4, 3)), g)
    bt = g.variable(b)
    z = x + bt
    return z, ag.grad([z], [bt]), [b], None


@case()
def wx_plus_b(ag, env, g, rng):             # :523-544
    w = env.slot().set(rng.standard_normal((2, 3)))
    b = env.slot().set(rng.standard_normal((1, 3)))
    x = ag.convert_to_tensor(rng.standard_normal((4, 2)), g)
    wt, bt = g.variable(w), g.variable(b)
    z = ag.matmul(x, wt) + bt
    return z, ag.grad([z], [wt, bt]), [w, b], None


def _reduce(name, keep, data=None):
    def c(ag, env, g, rng):
        v = env.slot().set(np.array(data) if data is not None else rng.standard_normal((3, 2)))
        vt = g.variable(v)
        z = getattr(ag, "reduce_" + name)(vt, [1], keep)
        return z, ag.grad([z], [vt]), [v], None
    c.__name__ = "reduce_%s%s" % (name, "_keep" if keep else "")
    c.tol = 1e-3
    return c


# :547-729
CASES += [_reduce("min", False, [[0., 1.], [3., 2.]]), _reduce("min", True, [[0., 1.], [3., 2.]]), _reduce("max", False, [[0., 1.], [3., 2.]]),
          _reduce("max", True, [[0., 1.], [3., 2.]]), _reduce("mean", False), _reduce("mean", True), _reduce("sum", False), _reduce("sum", True),
          _reduce("prod", False)]


@case()
def maximum(ag, env, g, rng):               # :732-751
    v1, v2 = env.slot().set(np.array([1., 2., 3.])), env.slot().set(np.array([4., 5., 6.]))
    a, b = g.variable(v1), g.variable(v2)
    z = ag.maximum(a, b)
    return z, ag.grad([z], [a, b]), [v1, v2], None


@case()
def minimum(ag, env, g, rng):               # :754-773
    v1, v2 = env.slot().set(np.array([1., 2., 3.])), env.slot().set(np.array([4., 5., 6.]))
    a, b = g.variable(v1), g.variable(v2)
    z = ag.minimum(a, b)
    return z, ag.grad([z], [a, b]), [v1, v2], None


@case()
def transpose(ag, env, g, rng):             # :901-919
    v = env.slot().set(rng.standard_normal((1, 2, 3, 4)))
    vt = g.variable(v)
    z = ag.transpose(vt, [2, 3, 0, 1])
    return z, ag.grad([z], [vt]), [v], None


@case()
def reshape_after_transpose(ag, env, g, rng):   # :922-941
    v = env.slot().set(rng.standard_normal((2, 3, 4)))
    vt = g.variable(v)
    z = ag.reshape(ag.transpose(vt, [2, 1, 0]), [4, 6])
    return z, ag.grad([z], [vt]), [v], None


@case()
def transpose_then_reshape_then_mm(ag, env, g, rng):   # :944-966
    v = env.slot().set(rng.standard_normal((1, 2, 3, 4, 5)))
    v2 = env.slot().set(rng.standard_normal((8, 2)))
    vt, v2t = g.variable(v), g.variable(v2)
    z = ag.matmul(ag.reshape(ag.transpose(vt, [4, 2, 3, 0, 1]), [15, 8]), v2t)
    return z, ag.grad([z], [vt]), [v], None


@case()
def add(ag, env, g, rng):                   # :969-989
    a, b = env.slot().set(rng.standard_normal((2, 2))), env.slot().set(rng.standard_normal((2, 2)))
    at, bt = g.variable(a), g.variable(b)
    z = at + bt
    return z, ag.grad([z], [at, bt]), [a, b], None


@case()
def mul(ag, env, g, rng):                   # :992-1012
    a, b = env.slot().set(rng.standard_normal((2, 2))), env.slot().set(rng.standard_normal((2, 2)))
    at, bt = g.variable(a), g.variable(b)
    z = at * bt
    return z, ag.grad([z], [at, bt]), [a, b], None


@case()
def div_broadcast(ag, env, g, rng):         # tests/test_binary_ops_grad.rs (division with broadcasting)
    a, b = env.slot().set(rng.standard_normal((3, 2))), env.slot().set(rng.uniform(1., 2., (1, 2)))
    at, bt = g.variable(a), g.variable(b)
    z = at / bt
    return z, ag.grad([z], [at, bt]), [a, b], None


@case()
def scalar_arith(ag, env, g, rng):          # tests/test_binary_ops_grad.rs:218-424 (scalar on either side)
    a = env.slot().set(rng.uniform(1., 2., (2, 3)))
    at = g.variable(a)
    z = (2. * at + 1.) / (3. - at * 0.5) - 1. / at
    return z, ag.grad([z], [at]), [a], None


@case()
def elu(ag, env, g, rng):                   # :1036-1054
    v = env.slot().set(rng.standard_normal((2, 2)))
    vt = g.variable(v)
    z = ag.elu(vt, 1.)
    return z, ag.grad([z], [vt]), [v], None


@case()
def relu(ag, env, g, rng):                  # :1057-1074
    v = env.slot().set(np.array([0.2, 0.5]))
    vt = g.variable(v)
    z = ag.relu(vt)
    return z, ag.grad([z], [vt]), [v], None


@case()
def logsumexp(ag, env, g, rng):             # :1098-1116
    v = env.slot().set(rng.standard_normal((2, 3)))
    vt = g.variable(v)
    z = ag.reduce_logsumexp(vt, 1, True)
    return z, ag.grad([z], [vt]), [v], None


@case()
def log_softmax(ag, env, g, rng):           # :1119-1137
    v = env.slot().set(rng.standard_normal((1, 3)))
    vt = g.variable(v)
    z = ag.log_softmax(vt, 1)
    return z, ag.grad([z], [vt]), [v], None


@case()
def softmax(ag, env, g, rng):               # activation_ops.rs:98-111
    v = env.slot().set(rng.standard_normal((3, 4)))
    vt = g.variable(v)
    z = ag.softmax(vt, 1) * ag.convert_to_tensor(rng.standard_normal((3, 4)), g)
    return z, ag.grad([z], [vt]), [v], None


@case()
def softmax_cross_entropy(ag, env, g, rng):     # :1140-1159
    v = env.slot().set(rng.standard_normal((1, 3)))
    t = ag.convert_to_tensor(np.array([[1., 0., 0.]]), g)
    vt = g.variable(v)
    z = ag.softmax_cross_entropy(vt, t)
    return z, ag.grad([z], [vt]), [v], None


@case()
def sigmoid_cross_entropy(ag, env, g, rng):     # :1162-1181
    v = env.slot().set(rng.standard_normal((1, 3)))
    t = ag.convert_to_tensor(rng.standard_normal((1, 3)), g)
    vt = g.variable(v)
    z = ag.sigmoid_cross_entropy(vt, t)
    return z, ag.grad([z], [vt]), [v], None


@case()
def sparse_softmax_cross_entropy(ag, env, g, rng):   # :1184-1203
    v = env.slot().set(rng.standard_normal((2, 3)))
    t = ag.convert_to_tensor(np.array([1., 0.]), g)
    vt = g.variable(v)
    z = ag.sparse_softmax_cross_entropy(vt, t)
    return z, ag.grad([z], [vt]), [v], None


@case()
def gather(ag, env, g, rng):                # :1206-1225
    v = env.slot().set(rng.standard_normal((5, 4, 8, 2)))
    vt = g.variable(v)
    x = ag.convert_to_tensor(np.array([[5., 4., 3.], [2., 1., 0.]]), g)
    z = ag.gather(vt, x, 2)
    return z, ag.grad([z], [vt]), [v], None


@case()
def concat(ag, env, g, rng):                # :1228-1248
    v1, v2 = env.slot().set(rng.standard_normal((1, 2))), env.slot().set(rng.standard_normal((1, 2)))
    a, b = g.variable(v1), g.variable(v2)
    z = ag.concat([a, b], 1)
    return z, ag.grad([z], [a]), [v1], None


@case()
def slice(ag, env, g, rng):                 # :1251-1269
    v = env.slot().set(rng.standard_normal((4, 4)))
    vt = g.variable(v)
    z = ag.slice(vt, [0, 0], [-1, 2])
    return z, ag.grad([z], [vt]), [v], None


@case()
def split(ag, env, g, rng):                 # :1272-1290
    v = env.slot().set(rng.standard_normal((3, 7, 5)))
    vt = g.variable(v)
    z = ag.split(vt, [2, 3, 2], 1)[1]
    return z, ag.grad([z], [vt]), [v], None


@case()
def flatten(ag, env, g, rng):               # :1293-1311
    v = env.slot().set(rng.standard_normal((4, 4)))
    vt = g.variable(v)
    z = ag.flatten(vt)
    return z, ag.grad([z], [vt]), [v], None


@case()
def reshape(ag, env, g, rng):               # :1314-1333
    v = env.slot().set(rng.standard_normal((4, 4)))
    vt = g.variable(v)
    z = ag.reshape(vt, [4, 2, 2])
    return z, ag.grad([z], [vt]), [v], None


@case(1e-2)
def conv2d_transpose(ag, env, g, rng):      # :1358-1371
    x, w = env.slot().set(rng.standard_normal((3, 2, 2, 2))), env.slot().set(rng.standard_normal((2, 3, 2, 2)))
    xt, wt = g.variable(x), g.variable(w)
    y = ag.conv2d_transpose(xt, wt, 0, 1)
    return y, ag.grad([y], [wt]), [w], None


@case(1e-2)
def conv2d_transpose_filter_grad(ag, env, g, rng):   # :1374-1396 (second order)
    x, w = env.slot().set(rng.standard_normal((2, 2, 2, 2))), env.slot().set(rng.standard_normal((2, 3, 2, 2)))
    xt, wt = g.variable(x), g.variable(w)
    y = ag.conv2d_transpose(xt, wt, 0, 1)
    gw = ag.grad([y], [wt])[0]
    return gw, ag.grad([gw], [wt]), [w], None


@case(1e-2)
def conv2d_filter_grad(ag, env, g, rng):    # :1399-1420 (second order)
    x, w = env.slot().set(rng.standard_normal((2, 3, 5, 5))), env.slot().set(rng.standard_normal((2, 3, 2, 2)))
    xt, wt = g.variable(x), g.variable(w)
    y = ag.conv2d(xt, wt, 0, 1)
    gw = ag.grad([y], [wt])[0]
    return gw, ag.grad([gw], [wt]), [w], None


@case(1e-2)
def conv2d_grad(ag, env, g, rng):           # :1423-1446 (grad of dgrad w.r.t. gy)
    x, w = env.slot().set(rng.standard_normal((2, 3, 5, 5))), env.slot().set(rng.standard_normal((2, 3, 2, 2)))
    gy = env.slot().set(np.ones((2, 2, 4, 4)))
    xt, wt, gyt = g.variable(x), g.variable(w), g.variable(gy)
    y = ag.conv2d(xt, wt, 0, 1)
    gx = ag.grad_with_default([y], [xt], [gyt])[0]
    return gx, ag.grad([gx], [gyt]), [gy], None


@case(1e-2)
def conv2d_xw_grad(ag, env, g, rng):        # :1449-1470 (grad of wgrad w.r.t. x)
    x, w = env.slot().set(rng.standard_normal((2, 3, 5, 5))), env.slot().set(rng.standard_normal((2, 3, 2, 2)))
    xt, wt = g.variable(x), g.variable(w)
    y = ag.conv2d(xt, wt, 0, 1)
    gw = ag.grad([y], [wt])[0]
    return gw, ag.grad([gw], [xt]), [x], None


@case(1e-2)
def conv2d(ag, env, g, rng):                # :1473-1493
    x, w = env.slot().set(rng.standard_normal((2, 3, 5, 5))), env.slot().set(rng.standard_normal((2, 3, 3, 3)))
    xt, wt = g.variable(x), g.variable(w)
    y = ag.conv2d(xt, wt, 1, 2)
    return y, ag.grad([y], [xt, wt]), [x, w], None


@case(1e-2)
def dilated_conv2d(ag, env, g, rng):        # tensor_ops/mod.rs:2763 (no reference test; same check as conv2d)
    x, w = env.slot().set(rng.standard_normal((2, 2, 7, 7))), env.slot().set(rng.standard_normal((3, 2, 3, 3)))
    xt, wt = g.variable(x), g.variable(w)
    y = ag.dilated_conv2d(xt, wt, 2, 1, 2)
    return y, ag.grad([y], [xt, wt]), [x, w], None


@case(1e-2)
def max_pool2d(ag, env, g, rng):            # :1496-1505
    x = env.slot().set(np.linspace(0., 1., 9))
    xt = g.variable(x)
    y = ag.max_pool2d(ag.reshape(xt, [1, 1, 3, 3]), 2, 0, 1)
    return y, ag.grad([y], [xt]), [x], None


@case(1e-2)
def max_pool2d_grad(ag, env, g, rng):       # :1508-1532 (second order)
    x = env.slot().set(np.linspace(0., 1., 36))
    gy = env.slot().set(rng.standard_normal((2, 2, 2, 2)))
    xt, gyt = g.variable(x), g.variable(gy)
    y = ag.max_pool2d(ag.reshape(xt, [2, 2, 3, 3]), 2, 0, 1)
    gx = ag.grad_with_default([y], [xt], [gyt])[0]
    return gx, ag.grad([gx], [gyt]), [gy], None


@case(1e-2)
def tensordot(ag, env, g, rng):             # :1535-1546
    a = env.slot().set(rng.standard_normal((3, 4, 5)))
    at = g.variable(a)
    b = ag.convert_to_tensor(rng.standard_normal((4, 3, 2)), g)
    c = ag.tensordot(at, b, [1, 0], [0, 1])
    return c, ag.grad([c], [at]), [a], None


@case(1e-2)
def primitive_back_propagation_through_time(ag, env, g, rng):   # :1549-1600 (the lstm_lm-style unrolled RNN)
    lookup = env.slot().set(rng.standard_normal((5, 3)))
    wo = env.slot().set(rng.standard_normal((3, 5)))
    wh = env.slot().set(rng.standard_normal((3, 3)))
    max_sent, batch = 3, 2
    lt, wot, wht = g.variable(lookup), g.variable(wo), g.variable(wh)
    sentences = g.placeholder("sents", [-1, max_sent])
    h = g.placeholder("h", [-1, 3])
    loss_buf = []
    for i in range(max_sent):
        cur = ag.squeeze(ag.slice(sentences, [0, i], [-1, i + 1]), [-1])
        nex = ag.squeeze(ag.slice(sentences, [0, (i + 1) % max_sent], [-1, (i + 1) % max_sent + 1]), [-1])
        x = ag.gather(lt, cur, 0)
        h = ag.tanh(ag.matmul(x, wht) + ag.matmul(h, wht))       # (sic) both terms use wh
        loss_buf.append(ag.sparse_softmax_cross_entropy(ag.matmul(h, wot), nex))
    loss = ag.add_n(loss_buf)
    feeds = [("sents", np.array([[2., 3., 1.], [0., 2., 0.]])), ("h", np.zeros((batch, 3)))]
    return loss, ag.grad([loss], [lt, wot, wht]), [lookup, wo, wh], feeds


@case(1e-2)
def lstm_step(ag, env, g, rng):             # examples/lstm_lm.rs:84-139 (toy LSTM LM: dim 4, vocab 5, seq 3, batch 2)
    D, V, B, S = 4, 5, 2, 3
    n = {k: env.slot().name(k).set(rng.standard_normal(shp) * 0.3) for k, shp in
         [("lookup", (V, D)), ("wx", (D, 4 * D)), ("wh", (D, 4 * D)), ("b", (1, 4 * D)), ("wo", (D, V))]}
    t = {k: g.variable(v) for k, v in n.items()}
    sents = g.placeholder("sents", [-1, S])
    h, c = g.placeholder("h0", [-1, D]), g.placeholder("c0", [-1, D])
    losses = []
    for i in range(S - 1):
        ids = ag.squeeze(ag.slice(sents, [0, i], [-1, i + 1]), [-1])
        nxt = ag.squeeze(ag.slice(sents, [0, i + 1], [-1, i + 2]), [-1])
        x = ag.gather(t["lookup"], ids, 0)
        xh = ag.matmul(x, t["wx"]) + ag.matmul(h, t["wh"]) + t["b"]
        gates = [ag.slice(xh, [0, k * D], [-1, (k + 1) * D]) for k in range(4)]
        i_, f_, o_, u_ = ag.sigmoid(gates[0]), ag.sigmoid(gates[1]), ag.sigmoid(gates[2]), ag.tanh(gates[3])
        c = f_ * c + i_ * u_
        h = o_ * ag.tanh(c)
        losses.append(ag.sparse_softmax_cross_entropy(ag.matmul(h, t["wo"]), nxt))
    loss = ag.add_n(losses)
    vs = list(t.values())
    feeds = [("sents", np.array([[2., 3., 1.], [0., 2., 0.]])), ("h0", np.zeros((B, D))), ("c0", np.zeros((B, D)))]
    return loss, ag.grad([loss], vs), list(n.values()), feeds


@case(3e-2)     # ReLU / max-pool kinks inside +-eps make central differences noisy here
def cnn_block(ag, env, g, rng):             # examples/cnn_mnist.rs:36-51 at toy size (conv + bias + relu + pool + fc + xent)
    w1 = env.slot().set(rng.standard_normal((4, 1, 3, 3)) * 0.3)
    b1 = env.slot().set(rng.standard_normal((1, 4, 6, 6)) * 0.1)
    w2 = env.slot().set(rng.standard_normal((4 * 3 * 3, 3)) * 0.3)
    ts = [g.variable(v) for v in (w1, b1, w2)]
    x = g.placeholder("x", [-1, 36])
    y = g.placeholder("y", [-1, 1])
    z = ag.relu(ag.conv2d(ag.reshape(x, [-1, 1, 6, 6]), ts[0], 1, 1) + ts[1])
    z = ag.reshape(ag.max_pool2d(z, 2, 0, 2), [-1, 36])
    loss = ag.reduce_mean(ag.sparse_softmax_cross_entropy(ag.matmul(z, ts[2]), y), [0], False)
    feeds = [("x", rng.uniform(0, 1, (5, 36))), ("y", np.array([[0.], [2.], [1.], [1.], [0.]]))]
    return loss, ag.grad([loss], ts), [w1, b1, w2], feeds


@case()
def normalize_batch_norm(ag, env, g, rng):  # tensor_ops/mod.rs:2325-2375 composites
    x = env.slot().set(rng.standard_normal((4, 3)))
    s, b = env.slot().set(rng.uniform(0.5, 1.5, (1, 3))), env.slot().set(rng.standard_normal((1, 3)))
    xt, st, bt = g.variable(x), g.variable(s), g.variable(b)
    z = ag.batch_norm(xt, st, bt) * ag.convert_to_tensor(rng.standard_normal((4, 3)), g)
    return z, ag.grad([z], [xt, st, bt]), [x, s, b], None


@case()
def reduce_variance_mse(ag, env, g, rng):   # tensor_ops/mod.rs:1291,1845
    a, b = env.slot().set(rng.standard_normal((3, 4))), env.slot().set(rng.standard_normal((3, 4)))
    at, bt = g.variable(a), g.variable(b)
    z = ag.reduce_variance(at, [1], False) + ag.mean_squared_error(at, bt)
    return z, ag.grad([z], [at, bt]), [a, b], None


@case()
def tile_leaky(ag, env, g, rng):            # tile (mod.rs:1039) + leaky_relu (:1695)
    a = env.slot().set(rng.standard_normal((1, 3)))      # Tile::grad = reduce_sum(gy, [axis], keep) (array_ops.rs:693-695): exact only for a unit axis
    at = g.variable(a)
    z = ag.leaky_relu(ag.tile(at, 0, 4), 0.1) * ag.convert_to_tensor(rng.standard_normal((4, 3)), g)
    return z, ag.grad([z], [at]), [a], None


@case()
def expr8_higher_order(ag, env, g, rng):    # tests/test_tensor_ops_eval.rs / doc tests: d2/dx2 of x^2 * y style expressions
    x, y = env.slot().set(np.array(3.)), env.slot().set(np.array(2.))
    xt, yt = g.variable(x), g.variable(y)
    z = 2. * xt * xt + 3. * yt + 1.
    gx = ag.grad([z], [xt])[0]
    return gx, ag.grad([gx], [xt]), [x], None
