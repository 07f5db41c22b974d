"""Op microbench sweep (BASELINE.json configs[4]): GEMM / batch_matmul M=N=K 256..8192, conv ResNet/VGG-shaped layers,
reduce_sum / softmax over 2^20..2^30 elements, elementwise, Adam.  CUDA-event timing on the library's stream, L2 flushed
between repetitions for the bandwidth kernels.  Roofline fractions use MEASURED_PEAKS.json (HBM copy GB/s; dense TF32 =
bf16/2).  Output: gpurun_out/ops_r2.json (copied to profiles/ as one JSON row per line)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_autograd_b200 as agb  # noqa: E402
from rust_autograd_b200 import ffi  # noqa: E402


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


def cl(dev, shape):
    """channels-last device tensor (logical [B,C,H,W])"""
    B, Cc, H, W = shape
    t = dev.empty((B, H, W, Cc))
    ffi.check(dev.lib.agb_fill(dev.ctx, t.desc(), 0.01))
    return agb.DArray(dev, t.ptr, (B, Cc, H, W), (H * W * Cc, 1, W * Cc, Cc), owner=t)


def main():
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    hbm, tf32 = pk["hbm_gbs"], pk["bf16_tflops"] / 2
    dev = agb.Device(0)
    lib = dev.lib
    rng = np.random.default_rng(0)
    out = {"peaks": {"hbm_gbs": hbm, "tf32_tflops_burst": tf32}, "rows": []}

    def row(name, ms, flops=None, bytes_=None):
        r = {"op": name, "ms": ms}
        if flops:
            r["tflops"] = flops / ms / 1e9
            r["frac_tf32_peak"] = r["tflops"] / tf32
        if bytes_:
            r["gbs"] = bytes_ / ms / 1e6
            r["frac_hbm"] = r["gbs"] / hbm
        out["rows"].append(r)
        print(json.dumps(r), flush=True)

    for n in (256, 512, 1024, 2048, 4096, 8192):
        a, b, c = dev.fill((n, n), 0.5), dev.fill((n, n), 0.25), dev.empty((n, n))
        for mode, nm in ((1, "tf32"), (0, "3xtf32"), (2, "fp32")):
            if mode == 2 and n > 4096:
                continue
            dev.set_math_mode(mode)
            row("matmul_%s_%d" % (nm, n), timeit(dev, lambda: dev.gemm(a, b, out=c), flush=False), flops=2.0 * n ** 3)
        for t in (a, b, c):
            t.free()
    for bt, n in ((64, 512), (16, 1024), (4, 2048)):
        a, b, c = dev.fill((bt, n, n), 0.5), dev.fill((bt, n, n), 0.25), dev.empty((bt, n, n))
        for mode, nm in ((1, "tf32"), (0, "3xtf32")):
            dev.set_math_mode(mode)
            row("batch_matmul_%s_%dx%d" % (nm, bt, n), timeit(dev, lambda: dev.gemm(a, b, out=c), flush=False), flops=2.0 * bt * n ** 3)
        for t in (a, b, c):
            t.free()
    layers = [(32, 64, 56, 64), (32, 128, 28, 128), (32, 256, 14, 256), (32, 512, 7, 512), (256, 64, 128, 64), (256, 128, 64, 128), (256, 256, 32, 256)]
    for mode, nm in ((1, "tf32"), (0, "3xtf32")):
        dev.set_math_mode(mode)
        for (B, Cc, H, O) in layers:
            x, gy = cl(dev, (B, Cc, H, H)), cl(dev, (B, O, H, H))
            y, gx = cl(dev, (B, O, H, H)), cl(dev, (B, Cc, H, H))
            w = dev.fill((O, Cc, 3, 3), 0.01)
            gw = dev.empty((O, Cc, 3, 3))
            fl = 2.0 * B * O * H * H * Cc * 9
            tag = "%s_B%d_C%d_H%d_O%d" % (nm, B, Cc, H, O)
            row("conv2d_fprop_" + tag, timeit(dev, lambda: ffi.check(lib.agb_conv2d_fprop_f32(dev.ctx, x.desc(), w.desc(), y.desc(), 1, 1, 1)), iters=5, flush=False), flops=fl)
            row("conv2d_transpose_" + tag, timeit(dev, lambda: ffi.check(lib.agb_conv2d_dgrad_f32(dev.ctx, gy.desc(), w.desc(), gx.desc(), 1, 1, 1)), iters=5, flush=False), flops=fl)
            row("conv2d_filter_grad_" + tag, timeit(dev, lambda: ffi.check(lib.agb_conv2d_wgrad_f32(dev.ctx, x.desc(), gy.desc(), gw.desc(), 1, 1, 1)), iters=5, flush=False), flops=fl)
            x = gy = y = gx = None
    # stride-2 stage transitions (ResNet-shaped): forward and filter gradient on the tensor cores, dgrad on the CUDA cores
    dev.set_math_mode(1)
    for (B, Cc, H, O) in ((32, 64, 56, 128), (32, 128, 28, 256), (32, 256, 14, 512)):
        Ho = (H + 2 - 3) // 2 + 1
        x, gy = cl(dev, (B, Cc, H, H)), cl(dev, (B, O, Ho, Ho))
        Hx = 2 * (Ho - 1) - 2 + 3            # conv2d_transpose output size follows the reference formula (conv2d_transpose.rs:55-56)
        y, gx = cl(dev, (B, O, Ho, Ho)), cl(dev, (B, Cc, Hx, Hx))
        w, gw = dev.fill((O, Cc, 3, 3), 0.01), dev.empty((O, Cc, 3, 3))
        fl = 2.0 * B * O * Ho * Ho * Cc * 9
        tag = "tf32_s2_B%d_C%d_H%d_O%d" % (B, Cc, H, O)
        row("conv2d_fprop_" + tag, timeit(dev, lambda: ffi.check(lib.agb_conv2d_fprop_f32(dev.ctx, x.desc(), w.desc(), y.desc(), 1, 2, 1)), iters=5, flush=False), flops=fl)
        row("conv2d_transpose_" + tag, timeit(dev, lambda: ffi.check(lib.agb_conv2d_dgrad_f32(dev.ctx, gy.desc(), w.desc(), gx.desc(), 1, 2, 1)), iters=5, flush=False), flops=fl)
        row("conv2d_filter_grad_" + tag, timeit(dev, lambda: ffi.check(lib.agb_conv2d_wgrad_f32(dev.ctx, x.desc(), gy.desc(), gw.desc(), 1, 2, 1)), iters=5, flush=False), flops=fl)
        x = gy = y = gx = None
    # first layer (C = 3, warp-MMA kernels), max-pool forward / gather-form backward with the ReLU gate, long softmax rows
    for mode, nm in ((1, "tf32"), (0, "3xtf32"), (2, "fp32")):
        dev.set_math_mode(mode)
        B, H, O = 256, 128, 64
        x = dev.fill((B, 3, H, H), 0.01)
        y, gy = cl(dev, (B, O, H, H)), cl(dev, (B, O, H, H))
        w, gw = dev.fill((O, 3, 3, 3), 0.01), dev.empty((O, 3, 3, 3))
        ybytes = 4.0 * B * O * H * H
        row("conv2d_fprop_firstlayer_%s_B256_C3_H128_O64" % nm, timeit(dev, lambda: ffi.check(lib.agb_conv2d_fprop_f32(dev.ctx, x.desc(), w.desc(), y.desc(), 1, 1, 1)), iters=5), bytes_=ybytes + 4.0 * B * 3 * H * H)
        row("conv2d_filter_grad_firstlayer_%s_B256_C3_H128_O64" % nm, timeit(dev, lambda: ffi.check(lib.agb_conv2d_wgrad_f32(dev.ctx, x.desc(), gy.desc(), gw.desc(), 1, 1, 1)), iters=5), bytes_=ybytes + 4.0 * B * 3 * H * H)
        x = y = gy = None
    dev.set_math_mode(1)
    for (B, Cc, H) in ((256, 64, 128), (256, 128, 64), (256, 256, 32)):
        x, gx = cl(dev, (B, Cc, H, H)), cl(dev, (B, Cc, H, H))
        y, idx, gy = cl(dev, (B, Cc, H // 2, H // 2)), cl(dev, (B, Cc, H // 2, H // 2)), cl(dev, (B, Cc, H // 2, H // 2))
        nx = 4.0 * B * Cc * H * H
        row("max_pool2d_fwd_B%d_C%d_H%d" % (B, Cc, H), timeit(dev, lambda: ffi.check(lib.agb_maxpool2d_fwd(dev.ctx, x.desc(), y.desc(), None, idx.ptr, 2, 0, 2)), iters=5), bytes_=nx * 1.5)
        row("max_pool2d_grad_scatter_B%d_C%d_H%d" % (B, Cc, H), timeit(dev, lambda: ffi.check(lib.agb_maxpool2d_bwd(dev.ctx, gy.desc(), None, idx.ptr, gx.desc())), iters=5), bytes_=nx * 1.5)
        row("max_pool2d_grad_tiled_relu_gate_B%d_C%d_H%d" % (B, Cc, H), timeit(dev, lambda: ffi.check(lib.agb_maxpool2d_bwd_fused(dev.ctx, gy.desc(), None, idx.ptr, y.ptr, None, gx.desc(), 2, 2)), iters=5), bytes_=nx * 1.75)
        x = gx = y = idx = gy = None
    for cols in (1024, 4096, 16384, 32768, 131072):
        n = 1 << 28
        x, z = dev.fill((n,), 1.0), dev.empty((n,))
        row("softmax_rows%d_2^28" % cols, timeit(dev, lambda: ffi.check(lib.agb_softmax(dev.ctx, x.ptr, z.ptr, n // cols, cols, 1)), iters=5), bytes_=8.0 * n)
        x.free(); z.free()
    for logn in (20, 24, 28, 30):
        n = 1 << logn
        x, z = dev.fill((n,), 1.0), dev.empty((n,))
        row("reduce_sum_2^%d" % logn, timeit(dev, lambda: ffi.check(lib.agb_reduce(dev.ctx, 0, x.ptr, z.ptr, 1, n, 1))), bytes_=4.0 * n)
        row("reduce_sum_rows4096_2^%d" % logn, timeit(dev, lambda: ffi.check(lib.agb_reduce(dev.ctx, 0, x.ptr, z.ptr, n // 4096, 4096, 1))), bytes_=4.0 * n)
        row("softmax_rows4096_2^%d" % logn, timeit(dev, lambda: ffi.check(lib.agb_softmax(dev.ctx, x.ptr, z.ptr, n // 4096, 4096, 1))), bytes_=8.0 * n)
        if logn <= 28:
            y = dev.fill((n,), 2.0)
            row("relu_2^%d" % logn, timeit(dev, lambda: ffi.check(lib.agb_unary(dev.ctx, ffi.U["relu"], 0.0, 0.0, x.desc(), z.desc()))), bytes_=8.0 * n)
            row("add_2^%d" % logn, timeit(dev, lambda: ffi.check(lib.agb_binary(dev.ctx, ffi.B["add"], 0.0, 0.0, x.desc(), y.desc(), z.desc()))), bytes_=12.0 * n)
            y.free()
        x.free(); z.free()
    n = 1 << 26
    p, g, m, v, t = dev.fill((n,), 1.0), dev.fill((n,), 0.5), dev.fill((n,), 0.0), dev.fill((n,), 0.0), dev.fill((1,), 1.0)
    row("adam_2^26", timeit(dev, lambda: dev.adam([p], [g], [m], [v], [t])), bytes_=28.0 * n)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ops_r2.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
