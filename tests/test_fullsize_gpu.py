"""-m gpu parity tests at the geometry the benchmarks time (VERDICT r1 "next" #1): the evaluator paths behind the headline numbers —
conv_rows + fused 2x2 pool epilogue, conv_cols, per-tap 256-wide tiles, all-taps filter gradient, masked dgrad with bias side sums at
128^2 / 64^2 / 32^2; the CNN-MNIST example at batch 200; the LSTM language model at hidden 1024 / vocab 8192 with row-stacked and
long-K GEMMs — compared with the oracle (oracle/ref_graph.py) on the same seeded inputs, in both tensor-core modes.

Tolerances (BASELINE.json north_star): 3xTF32 f32 tensors <= 2e-5 of the tensor's largest magnitude (1e-5 per contraction, graph depth
doubles it); TF32 loss / logits <= 1e-2, gradients in relative L2 (forward rounding legitimately flips ReLU masks and pool argmaxes of
pre-activations within rounding of a tie).  Pool argmax indices: exact except at near-ties (the two candidates differ by <= 1e-5 of the
map's largest value), and those must be rarer than 1e-4 of the windows in 3xTF32 mode."""
import numpy as np
import pytest

from oracle import ref_graph as OG

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ag():
    from rust_autograd_b200 import autograd
    return autograd


def rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def rel_l2(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))


def set_mode(env, mode):
    from rust_autograd_b200 import ffi
    ffi.check(ffi.load_library().agb_set_math_mode(env.agb_ctx(), mode))


# ------------------------------------------------------------------------------------------------ VGG stack at the bench geometry
_VGG_CACHE = {}


@pytest.mark.parametrize("mode", [0, 1], ids=["3xtf32", "tf32"])
def test_vgg_bench_geometry_matches_oracle(ag, mode):
    """workloads.vgg_loss at 3x128x128 (the bench's graph: examples/cnn_mnist.rs:36-51 widened to configs[3]), batch 8, through the protocol
    of oracle/parity.py: (1) every ReLU mask and pool argmax (max_pool2d.rs:21-88) equals the oracle's except at verified near-ties;
    (2) loss, logits and all 16 parameter gradients equal the oracle evaluated under the device's decisions.  The device run whose
    gradients are compared requests exactly what a training step requests, so conv_rows + fused pool epilogue, conv_cols, the 256-wide
    per-tap tiles, the all-taps filter gradient and the masked dgrad with bias side sums are the kernels under test."""
    from oracle import parity as P
    rng = np.random.default_rng(3)
    x = rng.standard_normal((8, 3, 128, 128)).astype(np.float32)
    y = rng.integers(0, 10, (8, 1)).astype(np.float32)
    res, _VGG_CACHE["ref"] = P.vgg_parity(ag, set_mode, mode, x, y, ref_unforced=_VGG_CACHE.get("ref"))
    tol = 2e-5 if mode == 0 else 1e-2
    assert res["forward_independent_of_targets"]
    d = res["decisions"]
    assert d["mismatches_are_near_ties"], d
    assert d["relu_mismatch_frac"] <= (1e-5 if mode == 0 else 5e-3) and d["pool_mismatch_frac"] <= (1e-4 if mode == 0 else 2e-2), d
    assert res["loss_rel"] <= tol and res["logits_rel"] <= tol, res
    assert len(res["grad_rel"]) == 16 and res["max_grad_rel"] <= tol, res["grad_rel"]


def test_vgg_training_steps_track_oracle(ag):
    """Three Adam steps of the bench graph (64x64 images, batch 4: the oracle does this in seconds; optimizers/mod.rs:66-82, adam.rs:11-58).
    The per-step losses are continuous in the weights and must agree to 1e-4.  The weights themselves are compared in relative L2: a single
    near-tie decision (see oracle/parity.py) changes the sign of a few near-zero gradient entries, and Adam's normalised first steps move
    such an entry by +-alpha regardless of its size, so an entry-wise bound tighter than 2 * alpha * steps is not a property of the
    algorithm; bit-level Adam parity incl. state is pinned on smooth graphs (test_optimizers_match_oracle, test_kernels_gpu adam tests)."""
    from rust_autograd_b200 import workloads as W
    rng = np.random.default_rng(7)
    xs = [rng.standard_normal((4, 3, 64, 64)).astype(np.float32) for _ in range(3)]
    ys = [rng.integers(0, 10, (4, 1)).astype(np.float32) for _ in range(3)]

    def train(mod, mode):
        env = mod.VariableEnvironment()
        if mode is not None:
            set_mode(env, mode)
        W.vgg_init(env, np.random.default_rng(0), size=64)
        # alpha = 1e-4: with the default 1e-3 this random-label problem explodes (loss 6 -> 87 -> 19) and the third loss then hangs on which side of a
        # near-tie a handful of step-2 decisions fell (both first-layer kernels are within 4e-7 of the oracle, yet their step-3 losses differ by 8e-4)
        adam = mod.optimizers.Adam(1e-4, 1e-08, 0.9, 0.999, env.default_namespace().current_var_ids(), env, "adam")
        losses = []
        for x, y in zip(xs, ys):
            def step(g):
                loss, _ = W.vgg_loss(mod, g, size=64)
                params, grads = mod.optimizers.grad_helper([loss], g.default_namespace())
                r = g.evaluator().push(loss).push(adam.get_update_op(params, grads, g)).feed("x", x).feed("y", y).run()
                losses.append(float(np.asarray(r[0].unwrap()).ravel()[0]))
            env.run(step)
        n_w = len(env.default_namespace().current_var_ids())
        n = n_w + len(env.namespace("adam").current_var_ids())
        out = [np.asarray(env.get_array_by_id(i)).copy() for i in range(n)]
        env.close()
        return losses, out, n_w
    l_ref, v_ref, n_w = train(OG, None)
    l_got, v_got, _ = train(ag, 0)
    assert np.allclose(l_got, l_ref, rtol=1e-4, atol=0), (l_got, l_ref)
    assert len(v_got) == len(v_ref)
    for k, (a, b) in enumerate(zip(v_got, v_ref)):
        if b.size == 1:                                   # the "{vid}t" step counters: exact
            assert np.array_equal(a, b), k
        elif k < n_w:                                     # weights: every entry within the three steps' travel, and close in L2
            assert float(np.abs(a.astype(np.float64) - b).max()) <= 6.1e-4 and (b.size < 1024 or rel_l2(a, b) <= 5e-3), (k, a.shape, rel_l2(a, b))
        else:                                             # Adam moments: linear in the UNFORCED gradients of three steps (near-tie ReLU / pool decisions differ: the
            assert rel_l2(a, b) <= 5e-2, (k, a.shape, rel_l2(a, b))      # single-step gradients under forced decisions are pinned to 2e-5 in test_vgg_bench_geometry_matches_oracle)


# ------------------------------------------------------------------------------------------------ CNN-MNIST at batch 200 (configs[1])
@pytest.mark.parametrize("mode", [0, 1], ids=["3xtf32", "tf32"])
def test_cnn_mnist_batch200_matches_oracle(ag, mode):
    """examples/cnn_mnist.rs:36-115 at BASELINE's batch 200 with dropout in training mode: the device masks are read back
    (nth_tensor(dropout, 1), random_ops.rs:218-245) and handed to the oracle; loss, logits and the six gradients."""
    from rust_autograd_b200 import workloads as W
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (200, 784)).astype(np.float32)
    y = rng.integers(0, 10, (200, 1)).astype(np.float32)

    env = ag.VariableEnvironment()
    set_mode(env, mode)
    W.cnn_mnist_init(env, np.random.default_rng(0))

    def body(g):
        taps = {}
        loss, logits = W.cnn_mnist_loss(ag, g, train=True, taps=taps)
        params, grads = ag.optimizers.grad_helper([loss], g.default_namespace())
        masks = [ag.nth_tensor(taps["drop%d" % i], 1) for i in range(3)]
        r = g.evaluator().push(loss).push(logits).extend(grads).extend(masks).feed("x", x).feed("y", y).run()
        return [np.asarray(v.unwrap()) for v in r]
    got = env.run(body)
    env.close()
    masks = got[-3:]
    for m in masks:
        assert set(np.unique(m)) <= {0.0, 1.0} and abs(float(m.mean()) - 0.75) < 0.02

    renv = OG.VariableEnvironment()
    W.cnn_mnist_init(renv, np.random.default_rng(0))

    def rbody(g):
        loss, logits = W.cnn_mnist_loss(OG, g, train=True, masks=masks)
        params, grads = OG.optimizers.grad_helper([loss], g.default_namespace())
        return [np.asarray(v.unwrap()) for v in g.evaluator().push(loss).push(logits).extend(grads).feed("x", x).feed("y", y).run()]
    ref = renv.run(rbody)
    assert len(ref) == 8
    tol = 2e-5 if mode == 0 else 1e-2
    for k, (a, b) in enumerate(zip(got[:8], ref)):
        if mode == 0 or k < 2:
            assert rel(a, b) <= tol, (k, a.shape, rel(a, b))
        else:
            assert rel_l2(a, b) <= 5e-2, (k, a.shape, rel_l2(a, b))


# ------------------------------------------------------------------------------------------------ LSTM LM at hidden 1024 / vocab 8192 (configs[2])
@pytest.mark.parametrize("mode", [0, 1], ids=["3xtf32", "tf32"])
def test_lstm_lm_full_width_matches_oracle(ag, mode):
    """examples/lstm_lm.rs:18-125 at the benchmarked width (batch 128, hidden 1024, vocab 8192), 9 tokens = 8 unrolled steps, so the
    evaluator's row-stacked x*wx / h*w_pred products, the long-K weight-gradient GEMMs, the stacked cross-entropy and the fused cell
    programs all engage: loss and the five parameter gradients."""
    from rust_autograd_b200 import workloads as W
    D, V, S, B = 1024, 8192, 9, 128
    sents = np.random.default_rng(9).integers(0, V, (B, S)).astype(np.float32)

    def run(mod, m):
        env = mod.VariableEnvironment()
        if m is not None:
            set_mode(env, m)
        W.lstm_init(env, np.random.default_rng(0), D, V, scale=0.05)

        def body(g):
            loss, _ = W.lstm_loss(mod, g, D, S)
            vs = [g.variable(k) for k in ("wx", "wh", "b", "lookup_table", "w_pred")]
            return [np.asarray(r.unwrap()) for r in g.evaluator().push(loss).extend(mod.grad([loss], vs)).feed("sents", sents).run()]
        try:
            return env.run(body)
        finally:
            env.close()
    ref, got = run(OG, None), run(ag, mode)
    assert len(got) == len(ref) == 6
    for k, (a, b) in enumerate(zip(got, ref)):
        assert rel(a, b) <= (5e-5 if mode == 0 else 2e-2), (mode, k, a.shape, rel(a, b))
