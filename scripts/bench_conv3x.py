"""3xTF32 conv layer timings (VGG shapes, B=256): fprop / dgrad / wgrad per layer."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import rust_autograd_b200 as agb
from rust_autograd_b200 import ffi
from bench_ops import timeit, cl
dev = agb.Device(0); lib = dev.lib
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev.set_math_mode(mode)
for (B, Cc, H, O) in [(256, 64, 128, 64), (256, 64, 64, 128), (256, 128, 64, 128), (256, 128, 32, 256), (256, 256, 32, 256)]:
    x, gy = cl(dev, (B, Cc, H, H)), cl(dev, (B, O, H, H)); y, gx = cl(dev, (B, O, H, H)), cl(dev, (B, Cc, H, H))
    w = dev.fill((O, Cc, 3, 3), 0.01); gw = dev.empty((O, Cc, 3, 3))
    fl = 2.0 * B * O * H * H * Cc * 9
    f = timeit(dev, lambda: ffi.check(lib.agb_conv2d_fprop_f32(dev.ctx, x.desc(), w.desc(), y.desc(), 1, 1, 1)), iters=3, flush=False)
    d = timeit(dev, lambda: ffi.check(lib.agb_conv2d_dgrad_f32(dev.ctx, gy.desc(), w.desc(), gx.desc(), 1, 1, 1)), iters=3, flush=False)
    g = timeit(dev, lambda: ffi.check(lib.agb_conv2d_wgrad_f32(dev.ctx, x.desc(), gy.desc(), gw.desc(), 1, 1, 1)), iters=3, flush=False)
    print("C%d O%d H%d: fprop %.3f ms %.0f TF/s | dgrad %.3f ms %.0f | wgrad %.3f ms %.0f" % (Cc, O, H, f, fl / f / 1e9, d, fl / d / 1e9, g, fl / g / 1e9), flush=True)
    x = gy = y = gx = None
dev.close()
