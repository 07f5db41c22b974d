"""Where a fused elementwise launch on a [128, 1024] tensor spends its time: in-graph time against the number of leaves, instructions and outputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import rust_autograd_b200 as agb
from rust_autograd_b200 import ffi
from bench_graph_micro import graph_time
dev = agb.Device(0)
U, B = ffi.F_UNARY, ffi.F_BINARY
R_, C_ = 128, 1024


def prepared(leaves, prog, out_regs):
    lv = (ffi.AgbFuseLeaf * len(leaves))()
    for i, (x, reg) in enumerate(leaves):
        lv[i] = ffi.AgbFuseLeaf(x.ptr, x.strides[0], x.strides[1], reg)
    ins = (ffi.AgbFuseInstr * len(prog))()
    for i, (kind, op, dst, a, b, p0) in enumerate(prog):
        ins[i] = ffi.AgbFuseInstr(kind, (ffi.U if kind == ffi.F_UNARY else ffi.B)[op], dst, a, b, p0)
    ys = [dev.empty((R_, C_)) for _ in out_regs]
    outs = (ffi.AgbFuseOut * len(out_regs))()
    for i, (yy, reg) in enumerate(zip(ys, out_regs)):
        outs[i] = ffi.AgbFuseOut(yy.ptr, C_, reg)
    keep = (lv, ins, outs, ys)
    return lambda: ffi.check(dev.lib.agb_fused_ewise(dev.ctx, R_, C_, len(leaves), lv, len(prog), ins, len(out_regs), outs)) or keep


xs = [dev.fill((R_, C_), 0.1 * (i + 1)) for i in range(13)]
for nl, ni, no, op in ((1, 1, 1, "square"), (1, 17, 1, "square"), (1, 17, 1, "tanh"), (13, 12, 1, "add"), (1, 1, 7, "square"), (13, 12, 7, "add"), (13, 17, 7, "cell")):
    leaves = [(xs[i], i) for i in range(nl)]
    if op == "add":
        prog = [(B, "add", 13, 0, 1, 0.0)] + [(B, "add", 13, 13, i, 0.0) for i in range(2, 13)]
        outs = [13] * no
    elif op == "cell":
        prog = [(B, "add", 13 + k, k, 4 + k, 0.0) for k in range(4)] + [(B, "add", 17 + k, 13 + k, 8 + k, 0.0) for k in range(4)] + \
               [(U, "sigmoid", 21, 17, 0, 0.0), (U, "sigmoid", 22, 18, 0, 0.0), (U, "tanh", 23, 19, 0, 0.0), (U, "sigmoid", 24, 20, 0, 0.0),
                (B, "mul", 25, 22, 12, 0.0), (B, "mul", 26, 21, 23, 0.0), (B, "add", 27, 25, 26, 0.0), (U, "tanh", 28, 27, 0, 0.0), (B, "mul", 29, 24, 28, 0.0)]
        outs = [21, 22, 24, 23, 27, 28, 29]
    else:
        prog = [(U, op, 1, 0, 0, 0.0)] + [(U, op, 1, 1, 0, 0.0) for _ in range(ni - 1)]
        outs = [1] * no
    print("leaves %2d instr %2d outs %d (%s): %.2f us" % (nl, len(prog), no, op, graph_time(dev, prepared(leaves, prog, outs))), flush=True)
dev.close()
