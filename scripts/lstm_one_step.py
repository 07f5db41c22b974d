import sys, os, numpy as np, ctypes as C
sys.path.insert(0, '/root/repo')
from rust_autograd_b200 import autograd as ag, ffi, workloads as W
env = ag.VariableEnvironment(); lib, ctx = ffi.load_library(), env.agb_ctx()
ffi.check(lib.agb_set_math_mode(ctx, 1))
D, V, S, B = 1024, 8192, 64, 128
W.lstm_init(env, np.random.default_rng(0), D, V)
adam = ag.optimizers.Adam.default("adam", env.default_namespace().current_var_ids(), env)
g = ag.Context(env)
loss, _ = W.lstm_loss(ag, g, D, S)
params, grads = ag.optimizers.grad_helper([loss], g.default_namespace())
upd = adam.get_update_op(params, grads, g)
sents = np.random.default_rng(1).integers(0, V, (B, S)).astype(np.float32)
for _ in range(2):
    g.evaluator().push(loss).push(upd).feed("sents", sents).run_async()
ffi.check(lib.agb_sync(ctx))
