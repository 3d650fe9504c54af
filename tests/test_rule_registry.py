"""The reference's extension point (`primitives.py:110-146`): `elemental_rules`, `defelemental`, `defelemental2`.
New elementwise primitives and overridden rules are traced once and lowered into the plan like built-in rules."""
import numpy as np
import pytest

import graphax_b200 as gx
from graphax_b200 import jnp
from graphax_b200.plan import build_plan
from graphax_b200.primitives import Primitive, defelemental, defelemental2, elemental_rules
from graphax_b200.tracer import make_jaxpr


@pytest.fixture
def clean_registry():
    before = dict(elemental_rules)
    yield
    elemental_rules.clear()
    elemental_rules.update(before)


def _softplus():
    softplus = Primitive("softplus", impl=lambda x: jnp.log1p(jnp.exp(x)))
    defelemental(softplus, lambda x: jnp.sigmoid(x))
    return softplus


def _replay64(plan, xs):
    from oracle import replay
    return replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)


def test_new_primitive_is_one_vertex_with_the_registered_partial(clean_registry):
    softplus = _softplus()

    def f(x, y):
        return softplus(x * y) + y, softplus(x)
    jaxpr = make_jaxpr(f, [(), (3,)])
    assert [e.primitive for e in jaxpr.eqns] == ["mul", "softplus", "add", "softplus"]
    plan = build_plan(jaxpr, "rev", (0, 1))
    # elemental edge of the new primitive: diagonal over the 3 components, like any elementwise rule
    edge = [d for s, t, d in plan.structure["elemental"] if t == 2][0]
    assert edge["out_dims"][0][0] == "S" and edge["val_shape"] == [3]
    rng = np.random.default_rng(0)
    x, y = rng.uniform(-1, 1, 64).astype(np.float32), rng.uniform(-1, 1, (64, 3)).astype(np.float32)
    jac, primal = _replay64(plan, [x, y])
    sig = lambda t: 1.0 / (1.0 + np.exp(-t))
    x64, y64 = x.astype(np.float64)[:, None], y.astype(np.float64)
    want = np.zeros((64, 4, 4))
    for k in range(3):
        want[:, k, 0] = sig(x64[:, 0] * y64[:, k]) * y64[:, k]
        want[:, k, 1 + k] = sig(x64[:, 0] * y64[:, k]) * x64[:, 0] + 1.0
    want[:, 3, 0] = sig(x64[:, 0])
    np.testing.assert_allclose(jac, want, rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(primal[:, 3], np.log1p(np.exp(x64[:, 0])), rtol=2e-6)
    # count_ops works on the new edges (same cost model as every diagonal edge)
    assert plan.stats["num_muls"] is not None and plan.stats["num_muls"] > 0


def test_primitive_without_rule_raises_like_the_reference(clean_registry):
    mystery = Primitive("mystery", impl=lambda x: x * 2.0)
    with pytest.raises(NotImplementedError, match="does not have registered elemental partial"):
        build_plan(make_jaxpr(lambda x: mystery(x), [()]), "rev", (0,))


def test_overriding_a_builtin_rule_and_defelemental2(clean_registry):
    f = lambda x, y: jnp.exp(x * y) + jnp.arcsinh(y)
    jaxpr = make_jaxpr(f, [(), ()])
    base = build_plan(jaxpr, "rev", (0, 1))
    defelemental2("exp", lambda out, x: out)                       # same derivative as the built-in rule
    defelemental("asinh", lambda x: 1. / jnp.sqrt(1. + x ** 2))     # the correction the golden generator applies
    same = build_plan(jaxpr, "rev", (0, 1))
    assert same.blob == base.blob                                  # identical operations -> identical plan
    defelemental("asinh", lambda x: jnp.sqrt(1. + x ** 2))          # the rule as shipped by the reference (:172)
    shipped = build_plan(jaxpr, "rev", (0, 1))
    x = np.array([0.3], np.float32)
    y = np.array([0.7], np.float32)
    j_base, _ = _replay64(base, [x, y])
    j_ship, _ = _replay64(shipped, [x, y])
    assert abs(j_base[0, 0, 1] - (0.3 * np.exp(0.21) + 1 / np.sqrt(1.49))) < 1e-6
    assert abs(j_ship[0, 0, 1] - (0.3 * np.exp(0.21) + np.sqrt(1.49))) < 1e-6
    elemental_rules["exp"] = lambda x: 2.0 * jnp.exp(x)            # direct assignment, a deliberately different rule
    doubled, _ = _replay64(build_plan(jaxpr, "rev", (0, 1)), [x, y])
    assert abs(doubled[0, 0, 0] - 2 * 0.7 * np.exp(0.21)) < 1e-6
    assert "exp" in elemental_rules and Primitive("exp2") not in elemental_rules


def test_rule_for_a_binary_primitive_with_a_literal_operand(clean_registry):
    hyp = Primitive("hypot", impl=lambda a, b: jnp.sqrt(a * a + b * b))
    defelemental2(hyp, lambda out, a, b: (a / out, b / out))
    f = lambda x, y: hyp(x, y) * hyp(x, 2.0)
    plan = build_plan(make_jaxpr(f, [(), ()]), "fwd", (0, 1))
    x, y = np.array([3.0], np.float32), np.array([4.0], np.float32)
    jac, primal = _replay64(plan, [x, y])
    h2 = np.sqrt(13.0)
    assert abs(primal[0, 0] - 5 * h2) < 1e-5
    np.testing.assert_allclose(jac[0, 0], [3 / 5 * h2 + 5 * 3 / h2, 4 / 5 * h2], rtol=1e-6)


@pytest.mark.gpu
def test_registered_primitive_runs_on_the_gpu(built_library, clean_registry):
    import torch
    softplus = _softplus()

    def f(x, y):
        return softplus(x * y) + jnp.sin(y), softplus(x) * y
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.uniform(-2, 2, 5000).astype(np.float32)).cuda()
    y = torch.from_numpy(rng.uniform(-2, 2, (5000, 3)).astype(np.float32)).cuda()
    jac = gx.vmap(gx.jacve(f, order="rev", argnums=(0, 1)))(x, y)
    xt, yt = x.double().requires_grad_(True), y.double().requires_grad_(True)
    sp = torch.nn.functional.softplus
    outs = (sp(xt[:, None] * yt) + torch.sin(yt), sp(xt)[:, None] * yt)
    for o, row in zip(outs, jac):
        for k in range(3):
            gx_, gy_ = torch.autograd.grad(o[:, k].sum(), (xt, yt), retain_graph=True)
            assert torch.allclose(row[0][:, k].double(), gx_, rtol=1e-5, atol=1e-6)
            assert torch.allclose(row[1][:, k, :].double(), gy_, rtol=1e-5, atol=1e-6)
