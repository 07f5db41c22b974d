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
