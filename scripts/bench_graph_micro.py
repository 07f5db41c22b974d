"""In-graph timings of the launch-latency-bound kernels of the LSTM language model: N copies of one launch are captured into a CUDA graph
(agb_graph_begin / agb_graph_end) and the replay is timed with CUDA events, so neither Python / ctypes launch overhead nor ncu's cold caches
are in the figure.  Rows: gpurun_out/graph_micro.jsonl."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_autograd_b200 as agb  # noqa: E402
from rust_autograd_b200 import ffi  # noqa: E402


def graph_time(dev, fn, n=64, reps=20):
    """us per launch of fn() when n of them replay back to back inside one CUDA graph"""
    for _ in range(3):
        fn()
    dev.sync()
    ffi.check(dev.lib.agb_graph_begin(dev.ctx))
    for _ in range(n):
        fn()
    g = C.c_void_p()
    ffi.check(dev.lib.agb_graph_end(dev.ctx, C.byref(g)))
    for _ in range(3):
        ffi.check(dev.lib.agb_graph_launch(dev.ctx, g))
    ts = []
    for _ in range(reps):
        a, b = dev.event(), dev.event()
        dev.record(a)
        ffi.check(dev.lib.agb_graph_launch(dev.ctx, g))
        dev.record(b)
        ts.append(dev.elapsed_ms(a, b))
    ffi.check(dev.lib.agb_graph_destroy(g))
    return float(np.median(ts)) * 1e3 / n


def main():
    dev = agb.Device(0)
    rows = []

    def row(name, us, **kw):
        r = dict(op=name, us_per_launch=us, **kw)
        rows.append(r)
        print(json.dumps(r), flush=True)
    U, B = ffi.F_UNARY, ffi.F_BINARY

    def prepared(rows_, cols_, leaves, prog, out_regs):
        lv = (ffi.AgbFuseLeaf * len(leaves))()
        for i, (x, reg) in enumerate(leaves):
            lv[i] = ffi.AgbFuseLeaf(x.ptr, 0 if x.shape[0] == 1 and rows_ != 1 else x.strides[0], 0 if x.shape[1] == 1 and cols_ != 1 else x.strides[1], reg)
        ins = (ffi.AgbFuseInstr * len(prog))()
        for i, (kind, op, dst, a, b, p0) in enumerate(prog):
            ins[i] = ffi.AgbFuseInstr(kind, (ffi.U if kind == ffi.F_UNARY else ffi.B)[op], dst, a, b, p0)
        ys = [dev.empty((rows_, cols_)) for _ in out_regs]
        outs = (ffi.AgbFuseOut * len(out_regs))()
        for i, (yy, reg) in enumerate(zip(ys, out_regs)):
            outs[i] = ffi.AgbFuseOut(yy.ptr, cols_, reg)
        keep = (lv, ins, outs, ys)
        return lambda: ffi.check(dev.lib.agb_fused_ewise(dev.ctx, rows_, cols_, len(leaves), lv, len(prog), ins, len(out_regs), outs)) or keep
    Bt, D = 128, 1024
    xw, hw, bias, c0 = dev.fill((Bt, 4 * D), 0.1), dev.fill((Bt, 4 * D), 0.2), dev.fill((1, 4 * D), 0.05), dev.fill((Bt, D), 0.3)
    leaves = [(xw.slice(1, k * D, (k + 1) * D), k) for k in range(4)] + [(hw.slice(1, k * D, (k + 1) * D), 4 + k) for k in range(4)] + \
             [(bias.slice(1, k * D, (k + 1) * D), 8 + k) for k in range(4)] + [(c0, 12)]
    prog = [(B, "add", 13 + k, k, 4 + k, 0.0) for k in range(4)] + [(B, "add", 17 + k, 13 + k, 8 + k, 0.0) for k in range(4)] + \
           [(U, "sigmoid", 21, 17, 0, 0.0), (U, "sigmoid", 22, 18, 0, 0.0), (U, "tanh", 23, 19, 0, 0.0), (U, "sigmoid", 24, 20, 0, 0.0),
            (B, "mul", 25, 22, 12, 0.0), (B, "mul", 26, 21, 23, 0.0), (B, "add", 27, 25, 26, 0.0), (U, "tanh", 28, 27, 0, 0.0), (B, "mul", 29, 24, 28, 0.0)]
    cell = prepared(Bt, D, leaves, prog, [21, 22, 24, 23, 27, 28, 29])
    row("fused_lstm_cell_128x1024", graph_time(dev, cell), bytes=4 * Bt * D * 20)
    # a trivial 1-instruction program and a plain single-op kernel of the same size: the floor of a launch inside a graph
    one = prepared(Bt, D, [(c0, 0)], [(U, "square", 1, 0, 0, 0.0)], [1])
    row("fused_square_128x1024", graph_time(dev, one))
    y1 = dev.empty((Bt, D))
    dc0, dy1 = c0.desc(), y1.desc()
    row("unary_square_128x1024", graph_time(dev, lambda: ffi.check(dev.lib.agb_unary(dev.ctx, ffi.U["square"], 0.0, 0.0, dc0, dy1))))
    for n in (1 << 22, 1 << 24):
        gy, y = dev.fill((1, n), 0.5), dev.fill((1, n), 0.25)
        p3 = [(U, "square", 2, 1, 0, 0.0), (B, "sub", 3, 1, 2, 0.0), (B, "mul", 4, 0, 3, 0.0)]
        us = graph_time(dev, prepared(1, n, [(gy, 0), (y, 1)], p3, [4]), n=8)
        row("fused_sigmoid_grad_%d" % n, us, gbs=12.0 * n / us / 1e3)
        gy.free(); y.free()
    # the recurrence GEMMs: h [128, 1024] x wh [1024, 4096] and its gradient g [128, 4096] x wh^T
    dev.set_math_mode(1)
    h, wh, o = dev.fill((Bt, D), 0.01), dev.fill((D, 4 * D), 0.01), dev.empty((Bt, 4 * D))
    row("gemm_128x1024x4096_tf32", graph_time(dev, lambda: dev.gemm(h, wh, out=o)))
    g, oh = dev.fill((Bt, 4 * D), 0.01), dev.empty((Bt, D))
    row("gemm_128x4096x1024_bT_tf32", graph_time(dev, lambda: dev.gemm(g, wh, trans_b=True, out=oh)))
    dev.set_math_mode(0)
    row("gemm_128x1024x4096_3xtf32", graph_time(dev, lambda: dev.gemm(h, wh, out=o)))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "graph_micro.jsonl"), "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    dev.close()


if __name__ == "__main__":
    main()
