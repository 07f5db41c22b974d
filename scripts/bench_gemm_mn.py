"""GEMMs with transposed (MN-major) operands at the shapes of the LSTM language model's stacked / long-K products (TF32 mode)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import rust_autograd_b200 as agb
from bench_ops import timeit
dev = agb.Device(0)
dev.set_math_mode(1)
for (m, k, n, ta, tb) in ((1024, 8064, 4096, True, False), (4096, 8064, 1024, True, False), (8064, 4096, 1024, False, True), (8064, 8192, 1024, False, True),
                          (1024, 8064, 8192, True, False), (8064, 1024, 4096, False, False), (4096, 4096, 4096, True, True), (128, 4096, 1024, False, True)):
    a = dev.fill((k, m) if ta else (m, k), 0.01); b = dev.fill((n, k) if tb else (k, n), 0.01); c = dev.empty((m, n))
    ms = timeit(dev, lambda: dev.gemm(a, b, trans_a=ta, trans_b=tb, out=c), flush=False)
    print("gemm m%d k%d n%d ta%d tb%d: %.3f ms %.0f TFLOP/s" % (m, k, n, ta, tb, ms, 2.0 * m * n * k / ms / 1e9), flush=True)
    a.free(); b.free(); c.free()
dev.close()
