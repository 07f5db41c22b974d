"""The BASELINE.json workloads as graph builders.  Each builder takes the tensor_ops module `T` (the CUDA engine's
``rust_autograd_b200.autograd`` — or any module with the same names) and follows the reference example it names, so the
same definition drives tests and benchmarks.  Synthetic data only: there is no network for datasets."""
import numpy as np

VGG_LAYERS = [(3, 64), (64, 64), "pool", (64, 128), (128, 128), "pool", (128, 256), (256, 256), (256, 256), "pool"]


def glorot_uniform(rng, shape):
    """ndarray_ext.rs:335-340: U(+-sqrt(6 / (fan_in + fan_out))) on the first two dims"""
    s = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-s, s, shape).astype(np.float32)


# ------------------------------------------------------------------------------------------------ examples/mlp_mnist.rs
def mlp_init(env, rng):
    env.slot().name("w").set(glorot_uniform(rng, (784, 10)))
    env.slot().name("b").set(np.zeros((1, 10), np.float32))


def mlp_loss(T, g):
    x, y = g.placeholder("x", [-1, 784]), g.placeholder("y", [-1, 1])
    z = T.matmul(x, g.variable("w")) + g.variable("b")
    return T.reduce_mean(T.sparse_softmax_cross_entropy(z, y), [0], False), z


# ------------------------------------------------------------------------------------------------ examples/cnn_mnist.rs:36-115
def cnn_mnist_init(env, rng):
    ns = env.default_namespace_mut()
    ns.slot().name("w1").set((rng.standard_normal((32, 1, 3, 3)) * 0.1).astype(np.float32))
    ns.slot().name("w2").set((rng.standard_normal((64, 32, 3, 3)) * 0.1).astype(np.float32))
    ns.slot().name("w3").set(glorot_uniform(rng, (64 * 7 * 7, 10)))
    ns.slot().name("b1").set(np.zeros((1, 32, 28, 28), np.float32))
    ns.slot().name("b2").set(np.zeros((1, 64, 14, 14), np.float32))
    ns.slot().name("b3").set(np.zeros((1, 10), np.float32))


def cnn_mnist_logits(T, g, train, masks=None, taps=None):
    """`taps` (a dict) collects the three dropout nodes so a parity test can read the device masks back (nth_tensor(d, 1)) and hand them to
    the oracle through `masks` (the RNG streams are parity-unpinned)."""
    def drop(t, i):
        if masks is not None:
            d = T.dropout(t, 0.25, train, mask=masks[i])         # oracle: explicit mask
        elif train is not None:
            d = T.dropout(t, 0.25, train)
        else:
            return t
        if taps is not None:
            taps["drop%d" % i] = d
        return d
    x = g.placeholder("x", [-1, 28 * 28]).reshape([-1, 1, 28, 28])
    z1 = T.conv2d(x, g.variable("w1"), 1, 1) + g.variable("b1")
    z2 = drop(T.max_pool2d(T.relu(z1), 2, 0, 2), 0)
    z3 = T.conv2d(z2, g.variable("w2"), 1, 1) + g.variable("b2")
    z4 = drop(T.max_pool2d(T.relu(z3), 2, 0, 2), 1)
    z5 = T.reshape(z4, [-1, 64 * 7 * 7])
    return drop(T.matmul(z5, g.variable("w3")) + g.variable("b3"), 2)


def cnn_mnist_loss(T, g, train=True, masks=None, taps=None):
    logits = cnn_mnist_logits(T, g, train, masks, taps)
    return T.reduce_mean(T.sparse_softmax_cross_entropy(logits, g.placeholder("y", [-1, 1])), [0], False), logits


# ------------------------------------------------------------------------------------------------ examples/lstm_lm.rs:18-125
def lstm_init(env, rng, dim, vocab, scale=0.01):
    """init_vars (lstm_lm.rs:55-82): N(0, 0.01) tables / gate weights, zero bias, U(0, 0.01) prediction weights"""
    ns = env.default_namespace_mut()
    ns.slot().name("lookup_table").set((rng.standard_normal((vocab, dim)) * scale).astype(np.float32))
    ns.slot().name("wx").set((rng.standard_normal((dim, 4 * dim)) * scale).astype(np.float32))
    ns.slot().name("wh").set((rng.standard_normal((dim, 4 * dim)) * scale).astype(np.float32))
    ns.slot().name("b").set(np.zeros((1, 4 * dim), np.float32))
    ns.slot().name("w_pred").set(rng.uniform(0, scale, (dim, vocab)).astype(np.float32))


def lstm_loss(T, g, dim, seq):
    """Unrolled LSTM language model (lstm_lm.rs:18-52,96-113): per step gather -> two gate GEMMs -> 4 slices -> sigmoid/tanh cell ->
    prediction GEMM -> sparse softmax cross-entropy; loss = add_n of the per-step losses.  h0 / c0 are zeros_like the first x."""
    v = {k: g.variable(k) for k in ("lookup_table", "wx", "wh", "b", "w_pred")}
    sents = g.placeholder("sents", [-1, seq])
    h = c = None
    losses = []
    for i in range(seq - 1):
        cur = T.slice(sents, [0, i], [-1, i + 1])
        nxt = T.slice(sents, [0, i + 1], [-1, i + 2])
        x = T.squeeze(T.gather(v["lookup_table"], cur, 0), [1])
        if h is None:
            h = T.zeros(T.shape(x), g)
            c = T.zeros(T.shape(x), g)
        xh = T.matmul(x, v["wx"]) + T.matmul(h, v["wh"]) + v["b"]
        i_, f_, c_, o_ = (T.slice(xh, [0, k * dim], [-1, (k + 1) * dim]) for k in range(4))
        c = T.sigmoid(f_) * c + T.sigmoid(i_) * T.tanh(c_)
        h = T.sigmoid(o_) * T.tanh(c)
        losses.append(T.sparse_softmax_cross_entropy(T.matmul(h, v["w_pred"]), nxt))
    return T.add_n(losses), None


# ------------------------------------------------------------------------------------------------ VGG-style stack (configs[3])
def vgg_init(env, rng, size=128, classes=10, layers=VGG_LAYERS):
    ns = env.default_namespace_mut()
    h, i, c_last = size, 0, 3
    for l in layers:
        if l == "pool":
            h //= 2
            continue
        cin, cout = l
        ns.slot().name("conv%d_w" % i).set((rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32))
        ns.slot().name("conv%d_b" % i).set(np.zeros((1, cout, 1, 1), np.float32))
        i, c_last = i + 1, cout
    ns.slot().name("fc_w").set(glorot_uniform(rng, (c_last * h * h, classes)))
    ns.slot().name("fc_b").set(np.zeros((1, classes), np.float32))
    return c_last * h * h


def vgg_loss(T, g, size=128, layers=VGG_LAYERS, taps=None, forced=None):
    """`taps` (a dict) collects the ReLU and max-pool nodes ("relu%d", "pool%d") so parity tests can read the activations and the argmax
    outputs (nth_tensor(p, 1)).  `forced` (oracle only: {"relu%d": 0/1 mask, "pool%d": argmax offsets}) evaluates the network under given
    discrete decisions (oracle/ref_graph.py relu_forced / max_pool2d_forced)."""
    x, y = g.placeholder("x", [-1, 3, size, size]), g.placeholder("y", [-1, 1])
    h, i, c_last, t, n_pool = size, 0, 3, x, 0
    for l in layers:
        if l == "pool":
            t = T.max_pool2d(t, 2, 0, 2) if forced is None else T.max_pool2d_forced(t, forced["pool%d" % n_pool], 2, 0, 2)
            if taps is not None:
                taps["pool%d" % n_pool] = t
            h //= 2
            n_pool += 1
            continue
        z = T.conv2d(t, g.variable("conv%d_w" % i), 1, 1) + g.variable("conv%d_b" % i)
        t = T.relu(z) if forced is None else T.relu_forced(z, forced["relu%d" % i])
        if taps is not None:
            taps["relu%d" % i] = t
        i, c_last = i + 1, l[1]
    logits = T.matmul(T.reshape(t, [-1, c_last * h * h]), g.variable("fc_w")) + g.variable("fc_b")
    return T.reduce_mean(T.sparse_softmax_cross_entropy(logits, y), [0], False), logits


def vgg_flops_per_sample(size=128, classes=10, layers=VGG_LAYERS):
    """fwd + bwd FLOPs of the contractions (fprop + dgrad + wgrad per conv; the first conv has no dgrad), per sample."""
    h, total, first, c_last = size, 0.0, True, 3
    for l in layers:
        if l == "pool":
            h //= 2
            continue
        f = 2.0 * l[1] * h * h * l[0] * 9
        total += f * (2 if first else 3)
        first, c_last = False, l[1]
    return total + 3 * 2.0 * c_last * h * h * classes
