"""CPU tests of the oracle's graph layer (oracle/ref_graph.py): every re-hosted reference gradient test (tests/refcases.py) is
run in float64 with the reference's own finite-difference checker (ag::test_helper::check_theoretical_grads,
src/test_helper.rs:9-148: perturb every scalar of every variable by +-eps, compare (f+ - f-)/2eps with the symbolic gradient).
This pins the oracle's Op::grad compositions before the GPU engine is compared against them."""
import numpy as np
import pytest

from oracle import ref_graph as OG
from oracle import ref_ops as R
import refcases


@pytest.fixture(autouse=True)
def f64():
    R.set_out_dtype(np.float64)
    yield
    R.set_out_dtype(np.float32)


def check_theoretical_grads(ag, env, g, objective, grads, var_ids, feeds, eps, tol):
    theo = [r.unwrap() for r in g.evaluator().extend(grads).feeds(feeds).run()]
    obj = ag.sum_all(objective)
    for vid, th in zip(var_ids, theo):
        base = env.get_array_by_id(vid)
        flat = base.ravel()
        for i in range(flat.size):
            vals = []
            for sgn in (+1, -1):
                pert = flat.copy()
                pert[i] += sgn * eps
                env.set_array_by_id(vid, pert.reshape(base.shape))
                vals.append(float(obj.eval(g, feeds)))
            env.set_array_by_id(vid, base)
            num = (vals[0] - vals[1]) / (2 * eps)
            assert abs(num - float(np.asarray(th).ravel()[i])) <= tol, (vid, i, num, float(np.asarray(th).ravel()[i]))


@pytest.mark.parametrize("case", refcases.CASES, ids=lambda c: c.__name__)
def test_reference_grad_case_on_oracle(case):
    env = OG.VariableEnvironment()
    rng = np.random.default_rng(1234)

    def body(g):
        z, grads, vids, feeds = case(OG, env, g, rng)
        check_theoretical_grads(OG, env, g, z, grads, vids, feeds, 1e-3, max(case.tol, 2e-3))
    env.run(body)


def test_optimizers_run_and_match_closed_form():
    """tests/test_optimizers.rs:10-64 only checks that update() does not panic; here the first Adam step is also checked."""
    for name in ("Adam", "AdaGrad", "MomentumSGD", "SGD"):
        env = OG.VariableEnvironment()
        rng = np.random.default_rng(0)
        w = env.slot().name("w").set(rng.standard_normal((2, 2)))
        b = env.slot().name("b").set(np.zeros((1, 2)))
        ids = env.default_namespace().current_var_ids()
        opt = OG.optimizers.SGD(0.1) if name == "SGD" else getattr(OG.optimizers, name).default("opt", ids, env)
        w0 = env.get_array_by_id(w)

        def body(g):
            x = OG.convert_to_tensor(np.ones((1, 2)), g)
            y = OG.convert_to_tensor(np.array([1.]), g)
            wt, bt = g.variable(w), g.variable(b)
            loss = OG.sparse_softmax_cross_entropy(OG.matmul(x, wt) + bt, y)
            grads = OG.grad([loss], [wt, bt])
            gw = grads[0].eval(g)
            opt.update([wt, bt], grads, g, OG.Feeder())
            return gw
        gw = env.run(body)
        w1 = env.get_array_by_id(w)
        assert not np.allclose(w1, w0)
        if name == "SGD":
            np.testing.assert_allclose(w1, w0 - 0.1 * gw, rtol=1e-6)
        if name == "Adam":
            np.testing.assert_allclose(w0 - w1, 1e-3 * np.sign(gw), rtol=1e-4)


def test_eval_semantics():
    """src/evaluation.rs:373-454 (test_eval, test_eval2, test_variable_eval, test_constant_eval, test_placeholder_eval)"""
    env = OG.VariableEnvironment()
    v = env.slot().set(np.array([[0., 1.], [2., 3.]]))

    def body(g):
        a = g.placeholder("a", [-1, 2])
        x = a + a
        r = g.evaluator().push(x).push(g.variable(v)).push(a).feed("a", np.ones((3, 2))).run()
        assert np.array_equal(r[0].unwrap(), 2 * np.ones((3, 2)))
        assert np.array_equal(r[1].unwrap(), np.array([[0., 1.], [2., 3.]]))          # variable target: cloned (:347-349)
        assert np.array_equal(r[2].unwrap(), np.ones((3, 2)))                          # placeholder target: copied (:350-352)
        bad = OG.matmul(x, OG.convert_to_tensor(np.ones((3, 3)), g))
        r = g.evaluator().push(bad + x).feed("a", np.ones((3, 2))).run()
        assert not r[0].is_ok() and r[0].err.kind == "IncompatibleShape"               # errors propagate to dependents (:202-211)
    env.run(body)


def test_oracle_hessian_vector_product_setdiff_map():
    """Oracle restatements of `_hessian_vector_product` (mod.rs:218-236), `setdiff1d` (doc example mod.rs:2044-2057) and `map`
    (higher_order_ops.rs:5-36): H v of f = sum(x^3) is 6 x v; the HVP also matches a central difference of the gradient."""
    from oracle import ref_graph as OG
    x0 = np.linspace(0.5, 2.0, 12).reshape(3, 4)
    v0 = np.linspace(-1.0, 1.0, 12).reshape(3, 4)
    env = OG.VariableEnvironment()
    vx = env.slot().set(x0)

    def body(g):
        x, v = g.variable(vx), g.placeholder("v", [3, 4])
        f = OG.sum_all(x * x * x)
        hv = OG._hessian_vector_product([f], [x], [v])[0]
        sd = OG.setdiff1d(OG.convert_to_tensor(np.array([4., 1., 5., 2., 3., 6.]), g), OG.convert_to_tensor(np.array([1., 3., 5.]), g))
        mp = OG.map(x * 2.0, lambda a: a[::-1] + 1.0)
        return [r.unwrap() for r in g.evaluator().extend([hv, sd, mp]).feed("v", v0).run()]
    hv, sd, mp = env.run(body)
    assert np.allclose(hv, 6.0 * x0 * v0, rtol=1e-5)
    eps = 1e-4
    fd = (3.0 * (x0 + eps * v0) ** 2 - 3.0 * (x0 - eps * v0) ** 2) / (2 * eps)       # d/d eps of grad f(x + eps v)
    assert np.allclose(hv, fd, rtol=1e-3)
    assert np.asarray(sd).tolist() == [2., 4., 6.]
    assert np.allclose(mp, (x0 * 2.0)[::-1] + 1.0, rtol=1e-6)


def test_parity_protocol_is_self_consistent():
    """oracle/parity.py with the oracle standing in for the device: its own decisions forced back in reproduce the unforced run exactly,
    so any difference the GPU tests see under forced decisions is the device's rounding, not the protocol."""
    import numpy as np
    from oracle import parity as P, ref_graph as OG
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3, 32, 32)).astype(np.float32)
    y = rng.integers(0, 10, (2, 1)).astype(np.float32)
    res, _ = P.vgg_parity(OG, lambda env, m: None, 0, x, y, size=32)
    assert res["max_grad_rel"] == 0.0 and res["loss_rel"] == 0.0 and res["max_grad_rel_unforced"] == 0.0
    assert res["decisions"] == {"relu_mismatch_frac": 0.0, "pool_mismatch_frac": 0.0, "mismatches_are_near_ties": True}
