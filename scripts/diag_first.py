"""debug: first-layer forward on the weights of test_vgg_training_steps_track_oracle after two Adam steps"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_autograd_b200 as agb
from rust_autograd_b200 import autograd as ag, ffi, workloads as W
from oracle import ref_ops as R
rng = np.random.default_rng(7)
xs = [rng.standard_normal((4, 3, 64, 64)).astype(np.float32) for _ in range(3)]
ys = [rng.integers(0, 10, (4, 1)).astype(np.float32) for _ in range(3)]
env = ag.VariableEnvironment()
ffi.check(ffi.load_library().agb_set_math_mode(env.agb_ctx(), 0))
W.vgg_init(env, np.random.default_rng(0), size=64)
adam = ag.optimizers.Adam.default("adam", env.default_namespace().current_var_ids(), env)
names = None
for i, (x, y) in enumerate(zip(xs, ys)):
    def step(g):
        loss, _ = W.vgg_loss(ag, g, size=64)
        params, grads = ag.optimizers.grad_helper([loss], g.default_namespace())
        r = g.evaluator().push(loss).push(adam.get_update_op(params, grads, g)).feed("x", x).feed("y", y).run()
        print("step", i, "loss", float(np.asarray(r[0].unwrap()).ravel()[0]), flush=True)
    if i == 2:
        break
    env.run(step)
ids = env.default_namespace().current_var_ids()
arrs = [np.asarray(env.get_array_by_id(k)).copy() for k in ids]
for k, a in zip(ids, arrs):
    print(k, a.shape, float(np.abs(a).max()))
w1 = [a for a in arrs if a.shape == (64, 3, 3, 3)][0]
b1 = [a for a in arrs if a.size == 64 and a.ndim >= 1][0].reshape(64)
dev = agb.Device(0)
x = xs[2]
ref = np.maximum(R.conv2d(x, w1, 1, 1, 1) + b1.reshape(1, 64, 1, 1), 0)
for mode in (0, 1, 2):
    dev.set_math_mode(mode)
    o = dev.conv2d(dev.upload(x), dev.upload(w1), 1, 1, 1, bias=dev.upload(b1), relu=True, channels_last=True).numpy()
    d = np.abs(o - ref)
    print("mode", mode, "max abs err", float(d.max()), "rel", float(d.max() / np.abs(ref).max()), "argmax", np.unravel_index(d.argmax(), d.shape), "nonzero mismatch", int(((o > 0) != (ref > 0)).sum()))
dev.close()
