"""First-layer conv (3 -> 64 @ 128x128, batch 256, NCHW input, channels-last output) forward + filter gradient: timing and the target of
single-kernel ncu captures.  argv: math mode (1 = TF32, 0 = 3xTF32)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import rust_autograd_b200 as agb
from rust_autograd_b200 import ffi
from bench_ops import cl, timeit
dev = agb.Device(0); lib = dev.lib
dev.set_math_mode(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
B, H, O = 256, 128, 64
x = dev.fill((B, 3, H, H), 0.01)
y, gy = cl(dev, (B, O, H, H)), cl(dev, (B, O, H, H))
w, gw = dev.fill((O, 3, 3, 3), 0.01), dev.empty((O, 3, 3, 3))
nbytes = 4.0 * B * O * H * H + 4.0 * B * 3 * H * H
f = timeit(dev, lambda: ffi.check(lib.agb_conv2d_fprop_f32(dev.ctx, x.desc(), w.desc(), y.desc(), 1, 1, 1)), iters=5)
g = timeit(dev, lambda: ffi.check(lib.agb_conv2d_wgrad_f32(dev.ctx, x.desc(), gy.desc(), gw.desc(), 1, 1, 1)), iters=5)
print("first layer: fprop %.3f ms (%.0f GB/s)  wgrad %.3f ms (%.0f GB/s)" % (f, nbytes / f / 1e6, g, nbytes / g / 1e6), flush=True)
dev.close()
