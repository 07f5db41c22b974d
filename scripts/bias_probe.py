"""Signed-error probe of the GEMM modes: mean and rms of (C - C64) relative to the row's sum of |products| (positive operands: the
coherent bias of truncating accumulation shows as a non-zero MEAN) and relative to max|C| (gaussian operands)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_autograd_b200 as agb
dev = agb.Device(0)
rng = np.random.default_rng(0)
for K in (1024, 8192):
    for dist in ("pos", "gauss"):
        M = N = 256
        a = (rng.uniform(0, 1, (M, K)) if dist == "pos" else rng.standard_normal((M, K))).astype(np.float32)
        b = (rng.uniform(0, 1, (K, N)) if dist == "pos" else rng.standard_normal((K, N))).astype(np.float32)
        ref = a.astype(np.float64) @ b.astype(np.float64)
        scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
        da, db = dev.upload(a), dev.upload(b)
        for mode, nm in ((0, "3xtf32"), (1, "tf32"), (2, "fp32")):
            dev.set_math_mode(mode)
            c = dev.gemm(da, db).numpy().astype(np.float64)
            e = (c - ref) / scale
            print("K=%d %s %-6s mean %.2e rms %.2e max|err|/max|C| %.2e" % (K, dist, nm, e.mean(), np.sqrt((e ** 2).mean()), np.abs(c - ref).max() / np.abs(ref).max()), flush=True)
dev.close()
