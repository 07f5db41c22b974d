"""max_pool2d forward / gated backward at the three VGG pooling geometries (channels-last, int32 argmax): GB/s against the measured HBM copy rate"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import rust_autograd_b200 as agb
from rust_autograd_b200 import ffi
from bench_ops import cl, timeit
hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
dev = agb.Device(0); lib = dev.lib
for (B, Cc, H) in ((256, 64, 128), (256, 128, 64), (256, 256, 32)):
    x, gx = cl(dev, (B, Cc, H, H)), cl(dev, (B, Cc, H, H))
    y, idx, gy = cl(dev, (B, Cc, H // 2, H // 2)), cl(dev, (B, Cc, H // 2, H // 2)), cl(dev, (B, Cc, H // 2, H // 2))
    cs = dev.empty((Cc,))
    nx = 4.0 * B * Cc * H * H
    f = timeit(dev, lambda: ffi.check(lib.agb_maxpool2d_fwd(dev.ctx, x.desc(), y.desc(), None, idx.ptr, 2, 0, 2)), iters=5)
    b = timeit(dev, lambda: ffi.check(lib.agb_maxpool2d_bwd_fused(dev.ctx, gy.desc(), None, idx.ptr, y.ptr, cs.ptr, gx.desc(), 2, 2)), iters=5)
    print("pool B%d C%d H%d: fwd %.3f ms %.0f GB/s (%.2f)   bwd+gate+sums %.3f ms %.0f GB/s (%.2f)" % (B, Cc, H, f, nx * 1.5 / f / 1e6, nx * 1.5 / f / 1e6 / hbm, b, nx * 1.75 / b / 1e6, nx * 1.75 / b / 1e6 / hbm), flush=True)
    x = gx = y = idx = gy = None
dev.close()
