"""Three eager training steps of one small config (mlp | cnn), for an ncu launch list: `ncu --metrics gpu__time_duration.sum --csv python scripts/small_step.py cnn`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rust_autograd_b200 import autograd as ag, ffi, workloads as W  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cnn"
env = ag.VariableEnvironment()
lib, ctx = ffi.load_library(), env.agb_ctx()
ffi.check(lib.agb_set_math_mode(ctx, 1))
rng = np.random.default_rng(0)
(W.mlp_init if which == "mlp" else W.cnn_mnist_init)(env, rng)
adam = ag.optimizers.Adam.default("adam", env.default_namespace().current_var_ids(), env)
g = ag.Context(env)
loss, _ = W.mlp_loss(ag, g) if which == "mlp" else W.cnn_mnist_loss(ag, g, train=True)
params, grads = ag.optimizers.grad_helper([loss], g.default_namespace())
upd = adam.get_update_op(params, grads, g)
x = rng.uniform(size=(200, 784)).astype(np.float32)
y = rng.integers(0, 10, (200, 1)).astype(np.float32)
for _ in range(3):
    g.evaluator().push(loss).push(upd).feed("x", x).feed("y", y).run_async()
ffi.check(lib.agb_sync(ctx))
