"""N eager training steps of the bench workload (VGG stack, batch 256, Adam) with device-resident inputs: the target of ncu captures.
    python scripts/one_step.py [tf32|3xtf32] [steps | count | profile]      ("count": print the number of kernel launches of one step;
    "profile": three warm-up steps, then one step inside a cudaProfilerStart / Stop range)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rust_autograd_b200 import autograd as ag, ffi, workloads as W  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
arg = sys.argv[2] if len(sys.argv) > 2 else "4"
lib = ffi.load_library()
env = ag.VariableEnvironment(0)
env.set_plan_cache(False)
ctx = env.agb_ctx()
ffi.check(lib.agb_set_math_mode(ctx, {"3xtf32": 0, "tf32": 1}[mode]))
rng = np.random.default_rng(0)
W.vgg_init(env, rng)
adam = ag.optimizers.Adam.default("adam", env.default_namespace().current_var_ids(), env)
B = 256
x = rng.standard_normal((B, 3, 128, 128)).astype(np.float32)
y = rng.integers(0, 10, (B, 1)).astype(np.float32)
feeds = {}
for name, a in (("x", x), ("y", y)):
    p = C.c_void_p(); ffi.check(lib.agb_alloc(ctx, a.nbytes, C.byref(p))); ffi.check(lib.agb_h2d(ctx, p, a.ctypes.data, a.nbytes))
    feeds[name] = ag.DeviceArray(p.value, a.shape)
ffi.check(lib.agb_sync(ctx))
g = ag.Context(env)
loss, _ = W.vgg_loss(ag, g)
params, grads = ag.optimizers.grad_helper([loss], g.default_namespace())
upd = adam.get_update_op(params, grads, g)


def step():
    g.evaluator().push(loss).push(upd).feed("x", feeds["x"]).feed("y", feeds["y"]).run_async()


if arg == "count":
    step(); step()
    ffi.check(lib.agb_sync(ctx))
    a, b = C.c_int64(), C.c_int64()
    ffi.check(lib.agb_launch_count(ctx, C.byref(a)))
    step()
    ffi.check(lib.agb_sync(ctx))
    ffi.check(lib.agb_launch_count(ctx, C.byref(b)))
    print(b.value - a.value)
elif arg == "profile":          # ncu --profile-from-start off: only the last step lies inside the cudaProfilerStart / Stop range
    rt = C.CDLL("libcudart.so")
    for _ in range(3):
        step()
    ffi.check(lib.agb_sync(ctx))
    rt.cudaProfilerStart()
    step()
    ffi.check(lib.agb_sync(ctx))
    rt.cudaProfilerStop()
else:
    for _ in range(int(arg)):
        step()
    ffi.check(lib.agb_sync(ctx))
