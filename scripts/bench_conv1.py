"""One-layer conv microbench (VGG conv1_2 shape by default) for ncu captures: fprop, dgrad(+mask), wgrad in TF32 mode."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import rust_autograd_b200 as agb  # noqa: E402
from rust_autograd_b200 import ffi  # noqa: E402
from bench_ops import cl, timeit  # noqa: E402


def main():
    B, Cc, H, O = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (256, 64, 128, 64))]
    iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    dev = agb.Device(0)
    lib = dev.lib
    dev.set_math_mode(1)
    x, gy, y, gx, m = cl(dev, (B, Cc, H, H)), cl(dev, (B, O, H, H)), cl(dev, (B, O, H, H)), cl(dev, (B, Cc, H, H)), cl(dev, (B, Cc, H, H))
    w, gw = dev.fill((O, Cc, 3, 3), 0.01), dev.empty((O, Cc, 3, 3))
    cs = dev.empty((Cc,))
    fl = 2.0 * B * O * H * H * Cc * 9
    for name, fn in (("fprop", lambda: ffi.check(lib.agb_conv2d_fprop_f32(dev.ctx, x.desc(), w.desc(), y.desc(), 1, 1, 1))),
                     ("dgrad_mask", lambda: ffi.check(lib.agb_conv2d_dgrad_fused_f32(dev.ctx, gy.desc(), w.desc(), m.desc(), cs.ptr, gx.desc(), 1, 1, 1))),
                     ("wgrad", lambda: ffi.check(lib.agb_conv2d_wgrad_f32(dev.ctx, x.desc(), gy.desc(), gw.desc(), 1, 1, 1)))):
        ms = timeit(dev, fn, iters=iters, warm=2, flush=False)
        print("%s B%d C%d H%d O%d: %.3f ms  %.1f TFLOP/s" % (name, B, Cc, H, O, ms, fl / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
