"""3xTF32 GEMM timings: n^3 for n in argv (default 8192 4096 2048), all four transpose forms at the first n; useful TFLOP/s = 2 n^3 / t."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import rust_autograd_b200 as agb
from bench_ops import timeit
dev = agb.Device(0)
dev.set_math_mode(0)
ns = [int(v) for v in sys.argv[1:]] or [8192, 4096, 2048]
for i, n in enumerate(ns):
    a, b, c = dev.fill((n, n), 0.5), dev.fill((n, n), 0.25), dev.empty((n, n))
    for ta, tb in ([(False, False), (True, False), (False, True), (True, True)] if i == 0 else [(False, False)]):
        t = timeit(dev, lambda: dev.gemm(a, b, trans_a=ta, trans_b=tb, out=c), iters=5, flush=False)
        print("gemm3x n=%d ta=%d tb=%d: %.3f ms %.1f TFLOP/s" % (n, ta, tb, t, 2.0 * n ** 3 / t / 1e9), flush=True)
    a = b = c = None
dev.close()
