"""GPU micro-probe (op microbench sweep, BASELINE configs[4]): times individual kernels with CUDA events on the
library's own stream and prints achieved TFLOP/s / GB/s.  Writes gpurun_out/probe.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_autograd_b200 as agb  # noqa: E402


def timeit(dev, fn, iters=10, warm=3, flush=True):
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


def main():
    which = sys.argv[1:] or ["gemm", "ewise", "conv"]
    dev = agb.Device(0)
    out = {}
    rng = np.random.default_rng(0)
    if "gemm" in which:
        for n in (1024, 2048, 4096, 8192):
            a = dev.upload(rng.standard_normal((n, n)).astype(np.float32))
            b = dev.upload(rng.standard_normal((n, n)).astype(np.float32))
            c = dev.empty((n, n))
            for mode, name in ((1, "tf32"), (0, "3xtf32"), (2, "fp32")):
                if mode == 2 and n > 4096:
                    continue
                dev.set_math_mode(mode)
                ms = timeit(dev, lambda: dev.gemm(a, b, out=c), flush=False)
                out["gemm_%s_%d" % (name, n)] = {"ms": ms, "tflops": 2 * n ** 3 / ms / 1e9}
                print("gemm", name, n, "%.3f ms  %.1f TFLOP/s" % (ms, 2 * n ** 3 / ms / 1e9), flush=True)
            for t in (a, b, c):
                t.free()
    if "ewise" in which:
        n = 1 << 28
        x, y = dev.fill((n,), 1.0), dev.fill((n,), 2.0)
        z = dev.empty((n,))
        import ctypes as C
        from rust_autograd_b200 import ffi
        lib = dev.lib
        tests = {
            "unary_relu": (8 * n, lambda: ffi.check(lib.agb_unary(dev.ctx, ffi.U["relu"], 0.0, 0.0, x.desc(), z.desc()))),
            "binary_add": (12 * n, lambda: ffi.check(lib.agb_binary(dev.ctx, ffi.B["add"], 0.0, 0.0, x.desc(), y.desc(), z.desc()))),
            "reduce_sum": (4 * n, lambda: ffi.check(lib.agb_reduce(dev.ctx, 0, x.ptr, z.ptr, 1, n, 1))),
            "reduce_sum_rows": (4 * n, lambda: ffi.check(lib.agb_reduce(dev.ctx, 0, x.ptr, z.ptr, 1 << 14, 1 << 14, 1))),
            "reduce_sum_cols": (4 * n, lambda: ffi.check(lib.agb_reduce(dev.ctx, 0, x.ptr, z.ptr, 1, 1 << 14, 1 << 14))),
            "softmax_4096": (8 * n, lambda: ffi.check(lib.agb_softmax(dev.ctx, x.ptr, z.ptr, n // 4096, 4096, 1))),
            "softmax_32768": (8 * n, lambda: ffi.check(lib.agb_softmax(dev.ctx, x.ptr, z.ptr, n // 32768, 32768, 1))),
        }
        for k, (bytes_, fn) in tests.items():
            ms = timeit(dev, fn)
            out[k] = {"ms": ms, "gbs": bytes_ / ms / 1e6}
            print(k, "%.3f ms  %.0f GB/s" % (ms, bytes_ / ms / 1e6), flush=True)
        m, v, t = dev.fill((n // 4,), 0.0), dev.fill((n // 4,), 0.0), dev.fill((1,), 1.0)
        p, g = dev.fill((n // 4,), 1.0), dev.fill((n // 4,), 0.5)
        ms = timeit(dev, lambda: dev.adam([p], [g], [m], [v], [t]))
        out["adam"] = {"ms": ms, "gbs": 28 * (n // 4) / ms / 1e6}
        print("adam %.3f ms %.0f GB/s" % (ms, out["adam"]["gbs"]), flush=True)
        for t_ in (x, y, z, m, v, p, g):
            t_.free()
    if "conv" in which:
        layers = [(32, 3, 128, 64), (32, 64, 128, 64), (32, 64, 64, 128), (32, 128, 64, 128), (32, 128, 32, 256), (32, 256, 32, 256)]
        for mode, name in ((1, "tf32"), (0, "3xtf32"), (2, "fp32")):
            dev.set_math_mode(mode)
            for (B, C, H, O) in layers:
                x = dev.upload(rng.standard_normal((B, C, H, H)).astype(np.float32))
                w = dev.upload((rng.standard_normal((O, C, 3, 3)) * 0.1).astype(np.float32))
                gy = dev.upload(rng.standard_normal((B, O, H, H)).astype(np.float32))
                fl = 2.0 * B * O * H * H * C * 9
                for kind, fn in (("fprop", lambda: dev.conv2d(x, w, 1, 1, 1).free()), ("dgrad", lambda: dev.conv2d_transpose(gy, w, 1, 1, 1).free()),
                                 ("wgrad", lambda: dev.conv2d_filter_grad(x, gy, (O, C, 3, 3), 1, 1, 1).free())):
                    ms = timeit(dev, fn, iters=5, warm=2, flush=False)
                    out["conv_%s_%s_B%d_C%d_H%d_O%d" % (kind, name, B, C, H, O)] = {"ms": ms, "tflops": fl / ms / 1e9}
                    print("conv", kind, name, (B, C, H, O), "%.3f ms  %.1f TFLOP/s" % (ms, fl / ms / 1e9), flush=True)
                for t_ in (x, w, gy):
                    t_.free()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
