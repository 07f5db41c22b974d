"""Bring-up helper: tries tcgen05 MN-major descriptor variants (AGB_MN_VARIANT=lbo,sbo,layout,kadv,atom32) in fresh
processes and reports the GEMM error of each; then sweeps K to characterise accumulation error per math mode."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import rust_autograd_b200 as agb
dev = agb.Device(0)
rng = np.random.default_rng(0)
for mode in (1, 0):
    dev.set_math_mode(mode)
    for (m, n, k, ta, tb) in [(128, 128, 128, False, False), (128, 128, 128, True, True), (256, 384, 96, True, False)]:
        a = rng.standard_normal((k, m) if ta else (m, k)).astype(np.float32)
        b = rng.standard_normal((n, k) if tb else (k, n)).astype(np.float32)
        c = dev.gemm(dev.upload(a), dev.upload(b), ta, tb).numpy()
        ref = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
        print("mode", mode, (m, n, k, ta, tb), "relerr %%.3e" %% (np.abs(c - ref).max() / np.abs(ref).max()), flush=True)
''' % ROOT

SWEEP = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import rust_autograd_b200 as agb
dev = agb.Device(0)
rng = np.random.default_rng(0)
for k in (256, 1024, 4096, 16384):
    a = rng.standard_normal((256, k)).astype(np.float32)
    b = rng.standard_normal((256, k)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    for mode in (0, 1, 2):
        dev.set_math_mode(mode)
        c = dev.gemm(dev.upload(a), dev.upload(b), False, True).numpy()
        e = c - ref
        print("K", k, "mode", mode, "max|e|/max|ref| %%.3e  rms %%.3e  mean(e*sign(ref))/rms(ref) %%.3e" %% (
            np.abs(e).max() / np.abs(ref).max(), np.sqrt((e ** 2).mean()) / np.sqrt((ref ** 2).mean()),
            (e * np.sign(ref)).mean() / np.sqrt((ref ** 2).mean())), flush=True)
''' % ROOT


def run(code, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=180)
    return (r.stdout + r.stderr[-600:]).strip()


if __name__ == "__main__":
    variants = ["4096,512,1,1024,1", "4096,1024,2,1024,0", "4096,1024,1,1024,1", "512,4096,1,1024,1", "4096,512,1,1024,0",
                "4096,256,1,1024,1", "1024,4096,2,1024,0"]
    for v in variants:
        print("=== AGB_MN_VARIANT=%s" % v, flush=True)
        print(run(CHILD, {"AGB_MN_VARIANT": v}), flush=True)
    print("=== K sweep (K-major operands)", flush=True)
    print(run(SWEEP, {}), flush=True)
