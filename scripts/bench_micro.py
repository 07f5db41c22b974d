"""The driver-visible slice of the op microbench sweep (BASELINE.json configs[4], `metric`: "GEMM TFLOP/s, reduce HBM GB/s vs peak"):
GEMM 8192^3 in TF32 and f32-faithful 3xTF32, reduce_sum over 2^28 elements, softmax over 2^28 elements in 4096-column rows, Adam on
2^26 parameters.  bench.py runs it after the training legs and puts the rows under "micro"; stand-alone it prints them.
CUDA events on the library's stream, median of `iters`; the bandwidth kernels work on 1 GiB arrays (>> the 126 MB L2) and L2 is
flushed between repetitions."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _time(dev, fn, iters, warm=3, flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            dev.flush_l2()
        a, b = dev.event(), dev.event()
        dev.record(a)
        fn()
        dev.record(b)
        ts.append(dev.elapsed_ms(a, b))
    return float(np.median(ts))


def run(peaks, device=0, iters=5):
    import rust_autograd_b200 as agb
    from rust_autograd_b200 import ffi
    dev = agb.Device(device)
    lib = dev.lib
    hbm = peaks["hbm_gbs"]
    tf_burst, tf_sus = peaks["bf16_tflops"] / 2.0, peaks["bf16_tflops_sustained"] / 2.0
    out = {}
    n = 8192
    a, b, c = dev.fill((n, n), 0.5), dev.fill((n, n), 0.25), dev.empty((n, n))
    for mode, nm, div in ((1, "tf32", 1.0), (0, "3xtf32", 3.0)):
        dev.set_math_mode(mode)
        ms = _time(dev, lambda: dev.gemm(a, b, out=c), iters, flush=False)
        tf = 2.0 * n ** 3 / ms / 1e9
        out["gemm_8192_" + nm] = {"ms": ms, "tflops": tf, "frac_burst": tf / (tf_burst / div), "frac_sustained": tf / (tf_sus / div),
                                  "peak_burst": tf_burst / div, "peak_sustained": tf_sus / div}
    for t in (a, b, c):
        t.free()
    dev.set_math_mode(1)
    n = 1 << 28
    x, z = dev.fill((n,), 1.0), dev.empty((n,))
    ms = _time(dev, lambda: ffi.check(lib.agb_reduce(dev.ctx, 0, x.ptr, z.ptr, 1, n, 1)), iters)
    out["reduce_sum_2^28"] = {"ms": ms, "gbs": 4.0 * n / ms / 1e6, "frac": 4.0 * n / ms / 1e6 / hbm}
    ms = _time(dev, lambda: ffi.check(lib.agb_softmax(dev.ctx, x.ptr, z.ptr, n // 4096, 4096, 1)), iters)
    out["softmax_2^28_rows4096"] = {"ms": ms, "gbs": 8.0 * n / ms / 1e6, "frac": 8.0 * n / ms / 1e6 / hbm}
    x.free(); z.free()
    n = 1 << 26
    p, g, m, v, t = dev.fill((n,), 1.0), dev.fill((n,), 0.5), dev.fill((n,), 0.0), dev.fill((n,), 0.0), dev.fill((1,), 1.0)
    ms = _time(dev, lambda: dev.adam([p], [g], [m], [v], [t]), iters)
    out["adam_2^26"] = {"ms": ms, "gbs": 28.0 * n / ms / 1e6, "frac": 28.0 * n / ms / 1e6 / hbm}
    for t_ in (p, g, m, v, t):
        t_.free()
    dev.close()
    return out


if __name__ == "__main__":
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    pk = json.load(open(f)) if os.path.exists(f) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    print(json.dumps(run(pk)))
