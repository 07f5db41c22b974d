"""Micro-timings of the kernels behind the evaluator-level regrouping: the fused elementwise program on a large and on
an LSTM-cell-sized tensor, row stacking, the stacked scatter-add and the device RNG.  Same method and row format as scripts/bench_ops.py
(CUDA events on the library's stream, L2 flushed between repetitions); rows are appended to gpurun_out/ops_new_r1.jsonl."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import rust_autograd_b200 as agb  # noqa: E402
from rust_autograd_b200 import ffi  # noqa: E402
from bench_ops import timeit  # noqa: E402


def main():
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = pk["hbm_gbs"]
    dev = agb.Device(0)
    rows = []

    def row(name, ms, bytes_=None, note=None):
        r = {"op": name, "ms": ms}
        if bytes_:
            r["gbs"] = bytes_ / ms / 1e6
            r["frac_hbm"] = r["gbs"] / hbm
        if note:
            r["note"] = note
        rows.append(r)
        print(json.dumps(r), flush=True)
    U, B = ffi.F_UNARY, ffi.F_BINARY

    def prepared(rows_, cols_, leaves, prog, out_regs):
        """ctypes argument blocks built once: the timed call is the C entry point alone (no Python marshalling between the events)"""
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
    # Sigmoid::grad chain gy * (y - square(y)) (activation_ops.rs:150): 2 leaves, 3 instructions, 1 stored value = 12 B/elem algorithmic
    # (the unfused sequence square -> sub -> mul moves 32 B/elem)
    for n in (1 << 22, 1 << 24):
        gy, y = dev.fill((1, n), 0.5), dev.fill((1, n), 0.25)
        prog = [(U, "square", 2, 1, 0, 0.0), (B, "sub", 3, 1, 2, 0.0), (B, "mul", 4, 0, 3, 0.0)]
        ms = timeit(dev, prepared(1, n, [(gy, 0), (y, 1)], prog, [4]))
        row("fused_sigmoid_grad_%d" % n, ms, 12.0 * n, "one launch; 12 B/elem algorithmic")
        t1, t2, t3 = dev.empty((1, n)), dev.empty((1, n)), dev.empty((1, n))
        dy, dgy, d1, d2, d3 = y.desc(), gy.desc(), t1.desc(), t2.desc(), t3.desc()

        def three():
            ffi.check(dev.lib.agb_unary(dev.ctx, ffi.U["square"], 0.0, 0.0, dy, d1))
            ffi.check(dev.lib.agb_binary(dev.ctx, ffi.B["sub"], 0.0, 0.0, dy, d1, d2))
            ffi.check(dev.lib.agb_binary(dev.ctx, ffi.B["mul"], 0.0, 0.0, dgy, d2, d3))
        row("unfused_sigmoid_grad_%d" % n, timeit(dev, three), 32.0 * n, "the three single-op launches the program replaces (32 B/elem moved)")
        for t in (t1, t2, t3):
            t.free()
        gy.free(); y.free()
    # the LSTM cell forward on [128, 1024] (13 leaves, 7 stored values): launch-latency territory
    Bt, D = 128, 1024
    xw, hw, bias, c0 = dev.fill((Bt, 4 * D), 0.1), dev.fill((Bt, 4 * D), 0.2), dev.fill((1, 4 * D), 0.05), dev.fill((Bt, D), 0.3)
    leaves = [(xw.slice(1, k * D, (k + 1) * D), k) for k in range(4)] + [(hw.slice(1, k * D, (k + 1) * D), 4 + k) for k in range(4)] + \
             [(bias.slice(1, k * D, (k + 1) * D), 8 + k) for k in range(4)] + [(c0, 12)]
    prog = [(B, "add", 13 + k, k, 4 + k, 0.0) for k in range(4)] + [(B, "add", 17 + k, 13 + k, 8 + k, 0.0) for k in range(4)] + \
           [(U, "sigmoid", 21, 17, 0, 0.0), (U, "sigmoid", 22, 18, 0, 0.0), (U, "tanh", 23, 19, 0, 0.0), (U, "sigmoid", 24, 20, 0, 0.0),
            (B, "mul", 25, 22, 12, 0.0), (B, "mul", 26, 21, 23, 0.0), (B, "add", 27, 25, 26, 0.0), (U, "tanh", 28, 27, 0, 0.0), (B, "mul", 29, 24, 28, 0.0)]
    ms = timeit(dev, prepared(Bt, D, leaves, prog, [21, 22, 24, 23, 27, 28, 29]), iters=20, flush=False)
    row("fused_lstm_cell_128x1024", ms, 4.0 * Bt * D * (13 + 7), "13 leaves (sliced / broadcast views), 17 instructions, 7 stored values; L2-resident")
    # row stacking: 63 blocks of [128, 8192]
    blocks = [dev.fill((128, 8192), float(i)) for i in range(63)]
    ms = timeit(dev, lambda: dev.concat_rows(blocks).free())
    row("concat_rows_63x128x8192", ms, 8.0 * 63 * 128 * 8192)
    for b in blocks:
        b.free()
    # stacked scatter-add: 8064 token ids into an [8192, 1024] table
    idx = dev.upload(np.random.default_rng(0).integers(0, 8192, 8064).astype(np.float32))
    gyr, table = dev.fill((8064, 1024), 1.0), dev.fill((8192, 1024), 0.0)
    ms = timeit(dev, lambda: dev.scatter_add(table, gyr, idx, 0))
    row("scatter_add_8064x1024", ms, 4.0 * 8064 * 1024 * 3, "read gy, atomic read-modify-write of the table rows")
    # device RNG (4 B/elem written)
    for kind, p0, p1 in (("uniform", 0.0, 1.0), ("normal", 0.0, 1.0), ("gamma", 2.0, 1.0)):
        n = 1 << 26
        ms = timeit(dev, lambda: dev.random(kind, (n,), p0, p1, seed=3).free())
        row("random_%s_%d" % (kind, n), ms, 4.0 * n)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ops_new_r1.jsonl"), "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    dev.close()


if __name__ == "__main__":
    main()
