"""Quick engine bring-up on the GPU: MLP training step, CNN step, a few grads against closed forms."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rust_autograd_b200 import autograd as ag
T = ag
from oracle import ref_ops as R

rng = np.random.default_rng(0)
env = ag.VariableEnvironment()
w0 = rng.standard_normal((784, 10)).astype(np.float32) * 0.05
w = env.slot().name("w").set(w0)
b = env.slot().name("b").set(np.zeros((1, 10), np.float32))
adam = ag.optimizers.Adam.default("adam", env.default_namespace().current_var_ids(), env)
xb = rng.uniform(size=(200, 784)).astype(np.float32)
yb = rng.integers(0, 10, (200, 1)).astype(np.float32)

def step(g):
    x = g.placeholder("x", [-1, 784]); y = g.placeholder("y", [-1, 1])
    wt, bt = g.variable("w"), g.variable("b")
    z = T.matmul(x, wt) + bt
    loss = T.sparse_softmax_cross_entropy(z, y)
    mean_loss = T.reduce_mean(loss, [0], False)
    grads = T.grad([mean_loss], [wt, bt])
    res = g.evaluator().push(mean_loss).extend(grads).feed("x", xb).feed("y", yb).run()
    l, gw, gb = [r.unwrap() for r in res]
    adam.update([wt, bt], grads, g, ag.Feeder().push("x", xb).push("y", yb))
    return l, gw, gb, g.size()

l, gw, gb, n = env.run(step)
z = R.binary_arith("add", R.matmul(xb, w0), np.zeros((1, 10), np.float32))
loss_r, logx = R.sparse_softmax_cross_entropy(z, yb)
gz = R.sparse_softmax_cross_entropy_grad(logx, yb, np.full((200, 1), 1 / 200, np.float32))
gw_r = R.matmul(xb, gz, True, False)
print("graph nodes", n, "loss", l, float(loss_r.mean()), "gw err", np.abs(gw - gw_r).max() / np.abs(gw_r).max(), "gb err", np.abs(gb - gz.sum(0, keepdims=True)).max())
w1 = env.get_array_by_id(w)
w1_r = R.adam_update(w0, gw_r, np.zeros_like(w0), np.zeros_like(w0), np.float32(1))[0]
print("adam err", np.abs(w1 - w1_r).max(), "t", env.namespace("adam").get_array_by_name("%dt" % w))
losses = []
t0 = time.time()
for i in range(50):
    losses.append(float(env.run(step)[0]))
print("50 steps %.1f ms/step; loss %.4f -> %.4f" % ((time.time() - t0) * 20, losses[0], losses[-1]))

# CNN (examples/cnn_mnist.rs)
env2 = ag.VariableEnvironment()
ns = env2.default_namespace()
ns.slot().name("w1").set(rng.standard_normal((32, 1, 3, 3)).astype(np.float32) * 0.1)
ns.slot().name("w2").set(rng.standard_normal((64, 32, 3, 3)).astype(np.float32) * 0.1)
ns.slot().name("w3").set(rng.uniform(-1, 1, (64 * 7 * 7, 10)).astype(np.float32) * np.sqrt(6 / (64 * 7 * 7)))
ns.slot().name("b1").set(np.zeros((1, 32, 28, 28), np.float32))
ns.slot().name("b2").set(np.zeros((1, 64, 14, 14), np.float32))
ns.slot().name("b3").set(np.zeros((1, 10), np.float32))
adam2 = ag.optimizers.Adam.default("adam", ns.current_var_ids(), env2)
xc = rng.uniform(size=(200, 784)).astype(np.float32)

def cnn_step(g):
    x = g.placeholder("x", [-1, 784]); y = g.placeholder("y", [-1, 1])
    x4 = x.reshape([-1, 1, 28, 28])
    z1 = T.conv2d(x4, g.variable("w1"), 1, 1) + g.variable("b1")
    z2 = T.dropout(T.max_pool2d(T.relu(z1), 2, 0, 2), 0.25, True)
    z3 = T.conv2d(z2, g.variable("w2"), 1, 1) + g.variable("b2")
    z4 = T.dropout(T.max_pool2d(T.relu(z3), 2, 0, 2), 0.25, True)
    z5 = T.reshape(z4, [-1, 64 * 7 * 7])
    logits = T.dropout(T.matmul(z5, g.variable("w3")) + g.variable("b3"), 0.25, True)
    loss = T.sparse_softmax_cross_entropy(logits, y)
    mean_loss = T.reduce_mean(loss, [0], False)
    params, grads = ag.optimizers.grad_helper([mean_loss], g.default_namespace())
    upd = adam2.get_update_op(params, grads, g)
    res = g.evaluator().push(mean_loss).push(upd).feed("x", xc).feed("y", yb).run()
    return float(res[0].unwrap()), g.size()

ls = []
t0 = time.time()
for i in range(30):
    l, n = env2.run(cnn_step)
    ls.append(l)
print("cnn: nodes", n, "30 steps %.1f ms/step; loss %.4f -> %.4f" % ((time.time() - t0) / 30 * 1e3, ls[0], ls[-1]))
print("ENGINE SMOKE OK")
