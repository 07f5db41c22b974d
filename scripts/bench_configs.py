"""Step time of the small BASELINE configs (configs[0..1]: MLP-MNIST, CNN-MNIST at batch 200) on the CUDA engine vs the CPU oracle.
These are launch-latency bound: reported as µs/step and samples/s, not against a roofline."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rust_autograd_b200 import autograd as ag, ffi, workloads as W  # noqa: E402


def run(name, init, loss_fn, feeds, mode, steps=200, warm=20, device=0, emit=True):
    env = ag.VariableEnvironment(device)
    lib, ctx = ffi.load_library(), env.agb_ctx()
    ffi.check(lib.agb_set_math_mode(ctx, mode))
    init(env, np.random.default_rng(0))
    adam = ag.optimizers.Adam.default("adam", env.default_namespace().current_var_ids(), env)
    g = ag.Context(env)
    loss, _ = loss_fn(ag, g)
    params, grads = ag.optimizers.grad_helper([loss], g.default_namespace())
    upd = adam.get_update_op(params, grads, g)
    dev = {}
    for k, a in feeds.items():
        import ctypes as C
        p = C.c_void_p(); ffi.check(lib.agb_alloc(ctx, a.nbytes, C.byref(p))); ffi.check(lib.agb_h2d(ctx, p, a.ctypes.data, a.nbytes))
        dev[k] = ag.DeviceArray(p.value, a.shape)
    ffi.check(lib.agb_sync(ctx))

    def step():
        ev = g.evaluator().push(loss).push(upd)
        for k, v in dev.items():
            ev.feed(k, v)
        ev.run_async()
    for _ in range(warm):
        step()
    ffi.check(lib.agb_sync(ctx))
    import ctypes as C
    l0, l1 = C.c_int64(), C.c_int64()
    ffi.check(lib.agb_launch_count(ctx, C.byref(l0)))
    t0 = time.time()
    for _ in range(steps):
        step()
    ffi.check(lib.agb_sync(ctx))
    dt = (time.time() - t0) / steps
    ffi.check(lib.agb_launch_count(ctx, C.byref(l1)))
    b = next(iter(feeds.values())).shape[0]
    row = {"config": name, "mode": {0: "3xtf32", 1: "tf32", 2: "fp32"}[mode], "us_per_step": dt * 1e6, "samples_per_s": b / dt,
           "launches_per_step": (l1.value - l0.value) / steps}
    # the same step captured into a CUDA graph (Evaluator.capture): no host-side graph walk per step
    ev = g.evaluator().push(loss).push(upd)
    for k, v in dev.items():
        ev.feed(k, v)
    sg = ev.capture()
    for _ in range(warm):
        sg.launch()
    ffi.check(lib.agb_sync(ctx))
    t0 = time.time()
    for _ in range(steps):
        sg.launch()
    ffi.check(lib.agb_sync(ctx))
    dt = (time.time() - t0) / steps
    row.update({"graph_us_per_step": dt * 1e6, "graph_samples_per_s": b / dt})
    sg.close()
    if "--profile" in sys.argv:      # where the eager step's device time goes: event pairs around every call of each kernel class (agb_prof_*)
        names = ["gemm", "conv_fprop", "conv_dgrad", "conv_wgrad", "ewise", "reduce", "softmax", "pool", "optim"]
        ffi.check(lib.agb_prof_enable(ctx, 1)); ffi.check(lib.agb_prof_reset(ctx))
        n = 3
        for _ in range(n):
            step()
        prof = {}
        for i, nm in enumerate(names):
            ms, calls, work = C.c_double(), C.c_int64(), C.c_double()
            ffi.check(lib.agb_prof_collect(ctx, i, C.byref(ms), C.byref(calls), C.byref(work)))
            if calls.value:
                prof[nm] = {"ms_per_step": ms.value / n, "calls_per_step": calls.value / n, "work_per_s": work.value / max(ms.value, 1e-9) * 1e3}
        ffi.check(lib.agb_prof_enable(ctx, 0))
        row["profile"] = prof
    if emit:
        print(json.dumps(row), flush=True)
    g.close(); env.close()
    return row


def small_configs(device=0, mode=1, lstm_steps=5):
    """bench.py's "configs" block: BASELINE configs[0..2] on one GPU, eager and replayed as a step graph (us per step)."""
    rng = np.random.default_rng(0)
    B = 200
    x = rng.uniform(size=(B, 784)).astype(np.float32); y = rng.integers(0, 10, (B, 1)).astype(np.float32)
    D, V, S, Bl = 1024, 8192, 64, 128
    sents = rng.integers(0, V, (Bl, S)).astype(np.float32)
    rows = [run("mlp_mnist_b200", W.mlp_init, W.mlp_loss, {"x": x, "y": y}, mode, steps=100, warm=10, device=device, emit=False),
            run("cnn_mnist_b200", W.cnn_mnist_init, lambda T, g: W.cnn_mnist_loss(T, g, train=True), {"x": x, "y": y}, mode, steps=100, warm=10, device=device, emit=False),
            run("lstm_lm_b128_s64_d1024_v8192", lambda env, r: W.lstm_init(env, r, D, V), lambda T, g: W.lstm_loss(T, g, D, S), {"sents": sents}, mode,
                steps=lstm_steps, warm=3, device=device, emit=False)]
    return {r["config"]: {"eager_us_per_step": r["us_per_step"], "graph_us_per_step": r["graph_us_per_step"], "graph_samples_per_s": r["graph_samples_per_s"],
                          "launches_per_step": r["launches_per_step"], "mode": r["mode"]} for r in rows}


def main():
    rng = np.random.default_rng(0)
    B = 200
    x = rng.uniform(size=(B, 784)).astype(np.float32); y = rng.integers(0, 10, (B, 1)).astype(np.float32)
    for mode in ((0, 1) if "--lstm-only" not in sys.argv else ()):
        run("mlp_mnist_b200", W.mlp_init, W.mlp_loss, {"x": x, "y": y}, mode)
        run("cnn_mnist_b200", W.cnn_mnist_init, lambda T, g: W.cnn_mnist_loss(T, g, train=True), {"x": x, "y": y}, mode)
    # LSTM language model: batch 128 x seq 64, hidden 1024, vocab 8192 -> 0.81 TFLOP per step
    D, V, S, Bl = 1024, 8192, 64, 128
    sents = rng.integers(0, V, (Bl, S)).astype(np.float32)
    for mode in ((1, 0) if "--tf32-only" not in sys.argv else (1,)):
        run("lstm_lm_b128_s64_d1024_v8192", lambda env, r: W.lstm_init(env, r, D, V), lambda T, g: W.lstm_loss(T, g, D, S), {"sents": sents}, mode, steps=10, warm=3)


if __name__ == "__main__":
    main()
