"""Bring-up helper for the tcgen05 conv kernels: runs each piece in a fresh process (a faulting kernel poisons the context)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEAD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import rust_autograd_b200 as agb
from oracle import ref_ops as R
dev = agb.Device(0); dev.set_math_mode(1)
rng = np.random.default_rng(0)
def rel(a, b): return float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))
''' % ROOT
CASES = {
 "gemm_tn64_occ2": "a=rng.standard_normal((64,256)).astype(np.float32); b=rng.standard_normal((256,128)).astype(np.float32); print('gemm', rel(dev.gemm(dev.upload(a),dev.upload(b)).numpy(), R.matmul(a,b)))",
 "gemm_big": "a=rng.standard_normal((512,512)).astype(np.float32); b=rng.standard_normal((512,512)).astype(np.float32); print('gemm', rel(dev.gemm(dev.upload(a),dev.upload(b)).numpy(), R.matmul(a,b)))",
 "fprop": "x=rng.standard_normal((2,64,32,32)).astype(np.float32); w=(rng.standard_normal((64,64,3,3))*0.1).astype(np.float32); print('fprop', rel(dev.conv2d(dev.upload(x),dev.upload(w),1,1,1).numpy(), R.conv2d(x,w,1,1,1)))",
 "dgrad": "g=rng.standard_normal((2,64,32,32)).astype(np.float32); w=(rng.standard_normal((64,64,3,3))*0.1).astype(np.float32); print('dgrad', rel(dev.conv2d_transpose(dev.upload(g),dev.upload(w),1,1,1).numpy(), R.conv2d_transpose(g,w,1,1,1)))",
 "wgrad_pair": "x=rng.standard_normal((2,64,32,32)).astype(np.float32); g=rng.standard_normal((2,64,32,32)).astype(np.float32); print('wgrad', rel(dev.conv2d_filter_grad(dev.upload(x),dev.upload(g),(64,64,3,3),1,1,1).numpy(), R.conv2d_filter_grad(x,g,(64,64,3,3),1,1,1)))",
 "wgrad_128": "x=rng.standard_normal((2,128,32,32)).astype(np.float32); g=rng.standard_normal((2,128,32,32)).astype(np.float32); print('wgrad', rel(dev.conv2d_filter_grad(dev.upload(x),dev.upload(g),(128,128,3,3),1,1,1).numpy(), R.conv2d_filter_grad(x,g,(128,128,3,3),1,1,1)))",
}
for name, body in CASES.items():
    code = HEAD + body + "\ndev.sync()\n"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    print("=== %s rc=%d\n%s%s" % (name, r.returncode, r.stdout[-400:], r.stderr[-500:]), flush=True)
    if r.returncode != 0 and "--sanitize" in sys.argv:
        r = subprocess.run(["compute-sanitizer", "--tool", "memcheck", "--print-limit", "3", sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        print("--- sanitizer:\n" + (r.stdout + r.stderr)[-2500:], flush=True)
