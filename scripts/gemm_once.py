"""One 8192^3 GEMM (mode from argv: 0 = 3xTF32, 1 = TF32), three launches — the target of single-kernel ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_autograd_b200 as agb
dev = agb.Device(0)
dev.set_math_mode(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
a, b, c = dev.fill((n, n), 0.5), dev.fill((n, n), 0.25), dev.empty((n, n))
for _ in range(3):
    dev.gemm(a, b, out=c)
dev.sync()
dev.close()
