"""Parity protocol for networks with discrete decisions (ReLU masks, max-pool argmaxes) — TEST / CHECKER INFRASTRUCTURE ONLY (used by
tests/test_fullsize_gpu.py and by bench.py's `parity` block; never imported by the package).

A filter gradient is a sum over ~10^4..10^6 positions with sqrt(N) cancellation, and the routing of every gradient value depends on
discrete forward decisions.  Two correct implementations whose forward values differ by one ulp can disagree on a decision at a near-tie,
and ONE such disagreement moves a gradient entry by ~1e-3 of its magnitude (measured here: the f32 C port of the reference,
oracle/cpu_ref.c, against the f64-accumulating numpy oracle on the VGG stack at 3x128x128, batch 8: conv0..conv2 gradients differ by
1e-3, conv3..conv6 by 4e-7 — one pool argmax differs).  So "same results as the reference" is checked in two parts:

  1. decisions: every ReLU mask and pool argmax the device produced equals the oracle's, except at near-ties (|pre-activation| resp. the
     difference of the two window candidates within `tie_tol` of the map's largest value), which must be rare;
  2. values: loss, logits and all parameter gradients equal the oracle's evaluated UNDER THE DEVICE'S DECISIONS (ref_graph.relu_forced /
     max_pool2d_forced), where the only remaining difference is floating-point rounding: <= 2e-5 (3xTF32 / fp32) or <= 1e-2 (TF32)
     of each tensor's largest magnitude.

The device is run twice: run A requests exactly what a training step requests (loss, logits, gradients) so the kernels and fusions are
the benchmarked ones; run B additionally requests the activations needed to read the decisions back (the ReLU outputs that feed a
convolution, the pooled maps and their argmax outputs).  ReLU layers that feed a max-pool are never materialised on the fused path; their
mask only matters at the argmax positions, where it equals (pooled value > 0)."""
import numpy as np

from . import ref_graph as OG


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _layer_plan(layers):
    """[(relu index, followed_by_pool index or None)] in network order"""
    plan, i, p = [], 0, 0
    for k, l in enumerate(layers):
        if l == "pool":
            p += 1
            continue
        nxt_pool = p if k + 1 < len(layers) and layers[k + 1] == "pool" else None
        plan.append((i, nxt_pool))
        i += 1
    return plan


def vgg_parity(ag, set_mode, mode, x, y, size=128, layers=None, device=0, tie_tol=None, ref_unforced=None):
    """Returns a dict: value errors under forced decisions + decision statistics.  `ag` = rust_autograd_b200.autograd,
    `set_mode(env, mode)` selects the math mode.  `ref_unforced` caches the oracle's own (unforced) run between modes."""
    from rust_autograd_b200 import workloads as W
    layers = layers or W.VGG_LAYERS
    plan = _layer_plan(layers)
    tie_tol = tie_tol if tie_tol is not None else (1e-5 if mode != 1 else 1e-2)

    def device_run(with_taps):
        env = ag.VariableEnvironment(device)
        set_mode(env, mode)
        W.vgg_init(env, np.random.default_rng(0), size=size, layers=layers)

        def body(g):
            taps = {}
            loss, logits = W.vgg_loss(ag, g, size=size, layers=layers, taps=taps)
            params, grads = ag.optimizers.grad_helper([loss], g.default_namespace())
            ev = g.evaluator().push(loss).push(logits).extend(grads)
            names = []
            if with_taps:
                for i, pool in plan:
                    if pool is None:
                        names.append("relu%d" % i)
                        ev.push(taps["relu%d" % i])
                for p in sorted(k for k in taps if k.startswith("pool")):
                    names += [p, p + "_idx"]
                    ev.push(taps[p]).push(ag.nth_tensor(taps[p], 1))
            out = [np.asarray(r.unwrap()) for r in ev.feed("x", x).feed("y", y).run()]
            n = 2 + len(grads)
            return out[:n], dict(zip(names, out[n:]))
        try:
            return env.run(body)
        finally:
            env.close()

    def oracle_run(forced):
        env = OG.VariableEnvironment()
        W.vgg_init(env, np.random.default_rng(0), size=size, layers=layers)

        def body(g):
            taps = {}
            loss, logits = W.vgg_loss(OG, g, size=size, layers=layers, taps=taps, forced=forced)
            params, grads = OG.optimizers.grad_helper([loss], g.default_namespace())
            ev = g.evaluator().push(loss).push(logits).extend(grads)
            names = sorted(taps)
            for k in names:
                ev.push(taps[k])
            pools = [k for k in names if k.startswith("pool")]
            for k in pools:
                ev.push(OG.nth_tensor(taps[k], 1))
            out = [np.asarray(r.unwrap()) for r in ev.feed("x", x).feed("y", y).run()]
            n = 2 + len(grads)
            d = dict(zip(names, out[n:n + len(names)]))
            d.update({k + "_idx": v for k, v in zip(pools, out[n + len(names):])})
            return out[:n], d
        return env.run(body)

    got, _ = device_run(False)                       # run A: the training step's own targets
    got_b, dec = device_run(True)                    # run B: + the activations that carry the decisions
    # the forward kernels are deterministic and do not depend on what else is requested
    fwd_same = bool(np.array_equal(got[0], got_b[0]) and np.array_equal(got[1], got_b[1]))
    if ref_unforced is None:
        ref_unforced = oracle_run(None)
    ref_u, dec_u = ref_unforced
    # ---- 1. decisions
    stats = {"relu_mismatch_frac": 0.0, "pool_mismatch_frac": 0.0, "mismatches_are_near_ties": True}
    forced = {}
    for i, pool in plan:
        if pool is None:
            a, b = dec["relu%d" % i], dec_u["relu%d" % i]
            ma, mb = a > 0, b > 0
            bad = ma != mb
            stats["relu_mismatch_frac"] = max(stats["relu_mismatch_frac"], float(bad.mean()))
            if bad.any() and float(np.maximum(np.abs(a[bad]), np.abs(b[bad])).max()) > tie_tol * float(np.abs(b).max()):
                stats["mismatches_are_near_ties"] = False
            forced["relu%d" % i] = ma.astype(np.float32)
        else:
            pv, pi = dec["pool%d" % pool], dec["pool%d_idx" % pool].astype(np.int64)
            m = np.ones(dec_u["relu%d" % i].shape, np.float32)
            m.ravel()[pi.ravel()] = (pv > 0).ravel().astype(np.float32)
            forced["relu%d" % i] = m
    for p in sorted(k for k in dec if k.startswith("pool") and not k.endswith("_idx")):
        iv, ir = dec[p + "_idx"].astype(np.int64), dec_u[p + "_idx"].astype(np.int64)
        bad = iv != ir
        stats["pool_mismatch_frac"] = max(stats["pool_mismatch_frac"], float(bad.mean()))
        if bad.any() and float(np.abs(dec[p][bad].astype(np.float64) - dec_u[p][bad]).max()) > tie_tol * float(np.abs(dec_u[p]).max()):
            stats["mismatches_are_near_ties"] = False
        forced[p] = iv
    # ---- 2. values under the device's decisions
    ref_f, _ = oracle_run(forced)
    res = {"loss_rel": _rel(got[0], ref_f[0]), "logits_rel": _rel(got[1], ref_f[1]),
           "grad_rel": [_rel(a, b) for a, b in zip(got[2:], ref_f[2:])],
           "grad_rel_unforced": [_rel(a, b) for a, b in zip(got[2:], ref_u[2:])],
           "forward_independent_of_targets": fwd_same, "decisions": stats}
    res["max_grad_rel"] = max(res["grad_rel"])
    res["max_grad_rel_unforced"] = max(res["grad_rel_unforced"])
    return res, ref_unforced
