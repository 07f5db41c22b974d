"""The three GEMMs of the VGG classifier (FC 65536 -> 10 at batch 256): forward, weight gradient, input gradient; CUDA-event timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import rust_autograd_b200 as agb
from bench_ops import timeit
dev = agb.Device(0)
dev.set_math_mode(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
B, K, N = 256, 65536, 10
a, w, g = dev.fill((B, K), 0.01), dev.fill((K, N), 0.01), dev.fill((B, N), 0.01)
y, gw, ga = dev.empty((B, N)), dev.empty((K, N)), dev.empty((B, K))
for det in (1, 0):
    dev.set_deterministic(det)
    f = timeit(dev, lambda: dev.gemm(a, w, out=y), flush=False)
    wg = timeit(dev, lambda: dev.gemm(a, g, trans_a=True, out=gw), flush=False)
    ig = timeit(dev, lambda: dev.gemm(g, w, trans_b=True, out=ga), flush=False)
    print("deterministic %d: forward %.1f us, weight grad %.1f us, input grad %.1f us" % (det, f * 1e3, wg * 1e3, ig * 1e3), flush=True)
dev.close()
