"""Cost / gain of the ReLU sign-bit side channel per VGG layer pair: forward with and without writing the bits, masked dgrad reading the activation
vs the bits (TF32 mode, batch 256, channels-last)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import rust_autograd_b200 as agb
from rust_autograd_b200 import ffi
from bench_ops import cl, timeit
import ctypes as C
dev = agb.Device(0); lib = dev.lib
dev.set_math_mode(1)
for (B, Cin, H, O, O2) in ((256, 3, 128, 64, 64), (256, 64, 64, 128, 128), (256, 128, 32, 256, 256), (256, 256, 32, 256, 256)):
    x = dev.fill((B, Cin, H, H), 0.01) if Cin <= 4 else cl(dev, (B, Cin, H, H))
    w, bias = dev.fill((O, Cin, 3, 3), 0.01), dev.fill((O,), 0.0)
    y = cl(dev, (B, O, H, H))
    bits = dev.empty((B * O * H * H // 32,))
    wr = C.c_int(0)
    f0 = timeit(dev, lambda: ffi.check(lib.agb_conv2d_fprop_fused_f32(dev.ctx, x.desc(), w.desc(), bias.ptr, 1, y.desc(), 1, 1, 1)), iters=5, flush=False)
    f1 = timeit(dev, lambda: ffi.check(lib.agb_conv2d_fprop_fused_bits_f32(dev.ctx, x.desc(), w.desc(), bias.ptr, 1, y.desc(), bits.ptr, C.byref(wr), 1, 1, 1)), iters=5, flush=False)
    gy, gx = cl(dev, (B, O2, H, H)), cl(dev, (B, O, H, H))
    w2, cs = dev.fill((O2, O, 3, 3), 0.01), dev.empty((O,))
    d0 = timeit(dev, lambda: ffi.check(lib.agb_conv2d_dgrad_fused_f32(dev.ctx, gy.desc(), w2.desc(), y.desc(), cs.ptr, gx.desc(), 1, 1, 1)), iters=5, flush=False)
    d1 = timeit(dev, lambda: ffi.check(lib.agb_conv2d_dgrad_fused_bits_f32(dev.ctx, gy.desc(), w2.desc(), y.desc(), bits.ptr, cs.ptr, gx.desc(), 1, 1, 1)), iters=5, flush=False)
    print("%d->%d @%d (next %d->%d): fprop %.3f -> %.3f ms with bits (written %d); masked dgrad %.3f -> %.3f ms; net %+.3f ms" %
          (Cin, O, H, O, O2, f0, f1, wr.value, d0, d1, (f1 - f0) + (d1 - d0)), flush=True)
    x = y = gy = gx = bits = None
dev.close()
