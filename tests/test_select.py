"""`jnp.where` / `lax.select_n` (reference rule `primitives.py:225-234`).  Deliberate deviation: the reference's rule
builds an inconsistent SparseTensor and raises (recorded in tests/golden/reference_run.json, quirk "jnp.where"), so
it cannot differentiate through a `where`; the engine passes the selected case's derivative (masks `which` and
`1 - which`, zero partial for the predicate), checked against torch.autograd."""
import json

import numpy as np
import pytest

from graphax_b200.plan import build_plan
from graphax_b200.tracer import jnp, make_jaxpr
from helpers import GOLDEN, bits, torch_eval_jaxpr

SHAPES = [(4,), (4,), ()]


def leaky(x, w, b):
    z = w * x + b
    a = jnp.where(z > 0., z, 0.1 * z)                                        # leaky ReLU
    c = jnp.where(x < w, jnp.log(a * a + 1.), jnp.sqrt(jnp.abs(x) + 1.))     # both branches finite everywhere
    return jnp.sum(a * c), jnp.maximum(a, 0.25)


def _inputs(batch, seed=0):
    rng = np.random.default_rng(seed)
    return [rng.normal(0, 1, (batch,) + s).astype(np.float32) for s in SHAPES]


def _tile(jaxpr, jac, batch):
    return np.concatenate([np.concatenate([b.reshape(batch, max(1, o.size), -1) for b in row], axis=2)
                           for o, row in zip(jaxpr.outvars, jac)], axis=1)


def test_reference_cannot_differentiate_through_where():
    assert json.loads((GOLDEN / "reference_run.json").read_text())["zoo"]["quirks"]["jnp.where"] == "IndexError"


def test_where_lowers_to_select_n():
    j = make_jaxpr(lambda x: jnp.where(x > 0., x, 0.1 * x), [(3,)])
    assert [e.primitive for e in j.eqns] == ["gt", "mul", "select_n"]
    assert [getattr(a, "eqn_index", None) for a in j.eqns[2].invars] == [0, 1, None]    # (which, on_false, on_true)


@pytest.mark.parametrize("order", ["fwd", "rev"])
def test_plan_replay_matches_oracle_and_torch(order):
    import torch
    from oracle import dense_ve, replay
    jaxpr = make_jaxpr(leaky, SHAPES)
    xs = _inputs(64)
    plan = build_plan(jaxpr, order, (0, 1, 2))
    rep, prim = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    j64, p64 = dense_ve.jacve(jaxpr, order, (0, 1, 2), [x.astype(np.float64) for x in xs], np.float64)
    t64 = _tile(jaxpr, j64, 64)
    assert np.abs(rep - t64).max() < 1e-5 * max(1.0, np.abs(t64).max())
    np.testing.assert_allclose(prim, np.concatenate([p.reshape(64, -1) for p in p64], axis=1), rtol=1e-5, atol=1e-6)
    for b in range(8):
        args = [torch.tensor(x[b], dtype=torch.float64, requires_grad=True) for x in xs]
        out = torch.cat([o.reshape(-1) for o in torch_eval_jaxpr(jaxpr, args)])
        rows = []
        for r in range(out.numel()):
            g = torch.autograd.grad(out[r], args, retain_graph=True, allow_unused=True)
            rows.append(torch.cat([(torch.zeros_like(a) if gi is None else gi).reshape(-1) for gi, a in zip(g, args)]))
        assert np.abs(torch.stack(rows).numpy() - t64[b]).max() < 1e-12


def test_select_of_unselected_nan_branch_keeps_the_value_finite():
    """The value is a real select (not a blend): log of a negative number in the branch not taken does not leak."""
    from oracle import replay
    jaxpr = make_jaxpr(lambda x: jnp.where(x > 0., jnp.log(x), x * 2.), [()])
    plan = build_plan(jaxpr, "rev", (0,))
    _, prim = replay.replay(plan.blob, [np.array([-1.5, 2.0], np.float32)], plan.n_rows, plan.n_cols)
    assert prim.reshape(-1)[0] == -3.0 and np.isclose(prim.reshape(-1)[1], np.log(2.0))


def test_ge_le_and_relu():
    import torch
    from graphax_b200.tracer import jnn
    from oracle import dense_ve, replay

    def f(x, y):
        clipped = jnp.where(x >= y, y, x)                 # min(x, y) through >=
        return jnn.relu(clipped - 0.2) * jnp.where(x <= 0.5, x * x, x)
    jaxpr = make_jaxpr(f, [(5,), (5,)])
    assert "ge" in [e.primitive for e in jaxpr.eqns] and "le" in [e.primitive for e in jaxpr.eqns]
    rng = np.random.default_rng(4)
    xs = [rng.uniform(-1, 2, (32, 5)).astype(np.float32) for _ in range(2)]
    plan = build_plan(jaxpr, "rev", (0, 1))
    rep, _ = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    j64, _ = dense_ve.jacve(jaxpr, "rev", (0, 1), [x.astype(np.float64) for x in xs], np.float64)
    t64 = _tile(jaxpr, j64, 32)
    assert np.abs(rep - t64).max() < 1e-5
    for b in range(4):
        args = [torch.tensor(x[b], dtype=torch.float64, requires_grad=True) for x in xs]
        out = torch_eval_jaxpr(jaxpr, args)[0]
        rows = [torch.cat([g.reshape(-1) for g in torch.autograd.grad(out[r], args, retain_graph=True)]) for r in range(5)]
        assert np.abs(torch.stack(rows).numpy() - t64[b]).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["jit", "interpreter"])
@pytest.mark.parametrize("order", ["fwd", "rev"])
def test_gpu_kernels_match_plan_replay(built_library, kernel, order):
    from helpers import run_engine
    from oracle import replay
    jaxpr = make_jaxpr(leaky, SHAPES)
    plan = build_plan(jaxpr, order, (0, 1, 2))
    xs = _inputs(1000 + 13, seed=3)
    jac, primal, used = run_engine(plan, xs, kernel)
    assert used == kernel
    ref_j, ref_p = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    scale = np.abs(ref_j).max(axis=(1, 2), keepdims=True) + 1e-30
    assert (np.abs(jac - ref_j) / scale).max() <= 2e-6           # logf differs by an ulp between CUDA and glibc
    np.testing.assert_allclose(primal, ref_p, rtol=2e-6, atol=1e-6)
