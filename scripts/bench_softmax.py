"""Softmax-family bandwidth rows only (the sweep of scripts/bench_ops.py takes minutes): 2^28 elements in rows of 128 .. 131072 columns,
softmax / log_softmax / logsumexp and the fused sparse cross-entropy; rows appended to gpurun_out/softmax_rows.jsonl."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import rust_autograd_b200 as agb  # noqa: E402
from rust_autograd_b200 import ffi  # noqa: E402
from bench_ops import timeit  # noqa: E402


def main():
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = pk["hbm_gbs"]
    dev = agb.Device(0)
    lib = dev.lib
    n = 1 << 28
    x, z = dev.fill((n,), 1.0), dev.empty((n,))
    out = []
    for cols in (128, 256, 512, 1024, 2048, 4096, 8192, 12288, 16384, 32768, 131072):
        rows = n // cols
        ms = timeit(dev, lambda: ffi.check(lib.agb_softmax(dev.ctx, x.ptr, z.ptr, rows, cols, 1)), iters=5)
        r = {"op": "softmax_rows%d_2^28" % cols, "ms": ms, "gbs": 8.0 * rows * cols / ms / 1e6}
        r["frac_hbm"] = r["gbs"] / hbm
        out.append(r)
        print(json.dumps(r), flush=True)
    for cols in (1024, 16384):
        rows = n // cols
        ms = timeit(dev, lambda: ffi.check(lib.agb_logsumexp(dev.ctx, x.ptr, z.ptr, rows, cols, 1)), iters=5)
        r = {"op": "logsumexp_rows%d_2^28" % cols, "ms": ms, "gbs": 4.0 * rows * cols / ms / 1e6}
        r["frac_hbm"] = r["gbs"] / hbm
        out.append(r)
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "softmax_rows.jsonl"), "w") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")
    dev.close()


if __name__ == "__main__":
    main()
