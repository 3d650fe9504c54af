"""Nested jacve (reference tests/examples/example_test.py:351-363, examples/economics.py:21-22): the Jacobian of
a function that itself returns a vertex-elimination Jacobian, i.e. the Hessian of the Black-Scholes price."""
import numpy as np
import pytest

import graphax_b200 as gx
from graphax_b200.examples.economics import BlackScholes, BlackScholes_Jacobian
from graphax_b200.plan import build_plan
from graphax_b200.tracer import make_jaxpr
from helpers import torch_eval_jaxpr

ARGS5 = [()] * 5


def _points(batch, seed=3):
    rng = np.random.default_rng(seed)
    return [rng.uniform(0.8, 1.25, batch) for _ in range(5)]       # around the reference's test point (1,1,1,1,1)


def _torch_hessian(xs, b):
    import torch
    inner = make_jaxpr(BlackScholes, ARGS5)
    f = lambda v: torch_eval_jaxpr(inner, [v[i] for i in range(5)])[0]
    return torch.autograd.functional.hessian(f, torch.tensor([x[b] for x in xs], dtype=torch.float64)).numpy()


def test_inner_plan_becomes_outer_equations():
    jaxpr = make_jaxpr(BlackScholes_Jacobian, ARGS5)
    assert len(jaxpr.outvars) == 5 and all(o.shape == () for o in jaxpr.outvars)
    assert len(jaxpr.eqns) > 2.5 * len(make_jaxpr(BlackScholes, ARGS5).eqns)   # primal + partials + products
    prims = {e.primitive for e in jaxpr.eqns}
    assert {"erf", "exp", "log", "sqrt", "div", "mul"} <= prims


@pytest.mark.parametrize("order", ["fwd", "rev"])
def test_hessian_matches_torch(order):
    from oracle import replay, ve_numpy
    jaxpr = make_jaxpr(BlackScholes_Jacobian, ARGS5)
    xs = _points(6)
    jac, _, _ = ve_numpy.jacve(jaxpr, order, range(5), xs, np.float64)
    H = np.array([[np.asarray(jac[o][a]).reshape(6) for a in range(5)] for o in range(5)]).transpose(2, 0, 1)
    for b in range(6):
        ref = _torch_hessian(xs, b)
        # the inner plan folds its constants (1/sqrt(2), 1/2, ...) in float32, torch keeps them in float64
        assert np.abs(H[b] - ref).max() <= 2e-7 * max(1.0, np.abs(ref).max())
        assert np.abs(H[b] - H[b].T).max() <= 2e-7 * np.abs(ref).max()        # second derivatives commute
    plan = build_plan(jaxpr, order, range(5))
    got, _ = replay.replay(plan.blob, [x.astype(np.float32) for x in xs], 5, 5)
    assert np.abs(got - H).max() <= 2e-5 * np.abs(H).max()


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["fwd", "rev"])
def test_gpu_second_order_greeks(built_library, order):
    """jacve(BlackScholes_Jacobian) at the reference's point (1,1,1,1,1) and vmapped over a batch."""
    import torch
    from oracle import replay
    jf = gx.jacve(BlackScholes_Jacobian, order=order, argnums=list(range(5)))
    one = jf(*[1.0] * 5)
    assert len(one) == 5 and len(one[0]) == 5
    H1 = np.array([[float(one[o][a]) for a in range(5)] for o in range(5)])
    ref = _torch_hessian([np.ones(1)] * 5, 0)
    assert np.abs(H1 - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-5          # the reference's tree_allclose tolerances
    xs = [x.astype(np.float32) for x in _points(500)]
    out = gx.vmap(jf)(*[torch.from_numpy(x).cuda() for x in xs])
    H = np.stack([np.stack([out[o][a].cpu().numpy() for a in range(5)], axis=1) for o in range(5)], axis=1)
    plan = jf.lower(ARGS5)
    want, _ = replay.replay(plan.blob, xs, 5, 5)
    assert np.abs(H - want).max() <= 2e-6 * np.abs(want).max()               # libm erf/exp/log: a few ulp host vs device
