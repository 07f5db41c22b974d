"""Diagnostic: per-tensor relative errors of the full-size parity cases (tests/test_fullsize_gpu.py) in every math mode."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_fullsize_gpu as T
from oracle import ref_graph as OG
from rust_autograd_b200 import autograd as ag, workloads as W

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "vgg"):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((8, 3, 128, 128)).astype(np.float32)
    y = rng.integers(0, 10, (8, 1)).astype(np.float32)
    ref, ng, npool = T._vgg_run(OG, None, x, y)
    for mode in (0, 1, 2):
        for fuse in ((True, False) if mode == 0 else (True,)):
            got, _, _ = T._vgg_run(ag, mode, x, y, fuse)
            print("vgg mode", mode, "fuse", fuse, " ".join("%d:%s:%.1e/%.1e" % (k, "x".join(map(str, a.shape)), T.rel(a, b), T.rel_l2(a, b)) for k, (a, b) in enumerate(zip(got[:2 + ng + npool], ref))))
            for k in range(npool):
                iv, ir = got[2 + ng + npool + k], ref[2 + ng + npool + k]
                print("   pool", k, "idx mismatch frac", float((iv != ir).mean()))
if which in ("all", "lstm"):
    D, V, S, B = 1024, 8192, 9, 128
    sents = np.random.default_rng(9).integers(0, V, (B, S)).astype(np.float32)
    def run(mod, m, fuse=True):
        env = mod.VariableEnvironment()
        if m is not None:
            T.set_mode(env, m); env.set_fusion(fuse)
        W.lstm_init(env, np.random.default_rng(0), D, V, scale=0.05)
        def body(g):
            loss, _ = W.lstm_loss(mod, g, D, S)
            vs = [g.variable(k) for k in ("wx", "wh", "b", "lookup_table", "w_pred")]
            return [np.asarray(r.unwrap()) for r in g.evaluator().push(loss).extend(mod.grad([loss], vs)).feed("sents", sents).run()]
        try:
            return env.run(body)
        finally:
            env.close()
    ref = run(OG, None)
    for mode in (0, 1, 2):
        for fuse in (True, False):
            got = run(ag, mode, fuse)
            print("lstm mode", mode, "fuse", fuse, " ".join("%s:%.1e" % ("x".join(map(str, a.shape)), T.rel(a, b)) for a, b in zip(got, ref)))
