"""Vertices of rank > 1 in the per-sample engine: the cases of the reference's tests/core/primitive_test.py
and its 4-8-4 NeuralNetwork test (tests/examples/example_test.py:168-196), restated against the jnp shim.

Chain of evidence: torch.autograd (independent autodiff) pins the dense vertex-elimination oracle in fp64;
the oracle pins the engine's plan (C replay, fp32) for forward and reverse orders; on the GPU the kernels
must reproduce the plan replay bit for bit."""
import numpy as np
import pytest

import graphax_b200 as gx
from graphax_b200.plan import build_plan
from graphax_b200.tracer import jnp, lax, make_jaxpr
from helpers import bits, torch_eval_jaxpr


def broadcast_add(x, y): return jnp.tanh(x + y)
def broadcast_sub(x, y): return jnp.tanh(x - y)
def broadcast_mul(x, y): return jnp.sin(x * jnp.exp(y))
def broadcast_div(x, y): return jnp.exp(x) / jnp.sum(jnp.exp(y), axis=1, keepdims=True) + y / x[0]
def outer_product(x, y): return jnp.sin(x * y)
def transpose(x, y): return jnp.cos(x).T * y
def matmul(x, y): return jnp.sin(x @ y)
def sums(x, y): return jnp.sin(jnp.sum(x @ y, axis=0))
def maxs(x, y): return jnp.sin(jnp.max(x @ y, axis=0))
def slicing(x, y): return jnp.sin((x @ y)[:, 0:1])
def squeezing(x, y): return jnp.squeeze(x @ y).sum()
def concatenate_1(x, y, z): return jnp.sin(x @ jnp.concatenate([y, z], axis=0))
def concatenate_2(x, y, z): return jnp.sin(jnp.concatenate([jnp.sin(x), jnp.cos(y), jnp.tanh(z)], axis=0))
def reshape(x, y): return jnp.sin(jnp.reshape(x, (2, 3)) @ y)
def large_matmul(x, y): return lax.dot_general(x, y, (([2], [0]), ([0], [1])))
def outer(x, y): return jnp.outer(x, y) + x.reshape(3, 1)
def indexing(x, y): return x[1] * y[:, 2] + x[0, 1:3].sum()


def array_split(x, y):
    z, u, v = jnp.split(x, [2, 4], axis=0)
    return jnp.sin(z) @ y, u + jnp.log(v)


def NeuralNetwork(x, W1, b1, W2, b2, y):
    a1 = jnp.tanh(W1 @ x + b1)
    a2 = jnp.tanh(W2 @ a1 + b2)
    d = a2 - y
    return .5 * jnp.sum(d ** 2)


CASES = {
    "broadcast_add": (broadcast_add, [(2, 3), (1, 3)]), "broadcast_sub": (broadcast_sub, [(2, 3), (1, 3)]),
    "broadcast_mul": (broadcast_mul, [(2, 3), (3,)]), "broadcast_div": (broadcast_div, [(2, 3), (2, 3)]), "outer_product": (outer_product, [(4, 1), (1, 3)]),
    "transpose": (transpose, [(2, 3), (3, 2)]), "matmul": (matmul, [(2, 3), (3,)]),
    "reduce_sum": (sums, [(2, 3), (3, 4)]), "reduce_max": (maxs, [(2, 3), (3, 4)]),
    "slicing": (slicing, [(2, 3), (3, 4)]), "squeezing": (squeezing, [(2, 3), (3, 1)]),
    "concatenate_1": (concatenate_1, [(2, 3), (2, 4), (1, 4)]), "concatenate_2": (concatenate_2, [(4,), (2,), (3,)]),
    "reshape": (reshape, [(6,), (3,)]), "large_matmul": (large_matmul, [(3, 1, 4), (4, 3, 2)]),
    "outer": (outer, [(3,), (4,)]), "array_split": (array_split, [(6, 3), (3, 2)]), "indexing": (indexing, [(3, 4), (4, 3)]),
    "NeuralNetwork": (NeuralNetwork, [(4,), (8, 4), (8,), (4, 8), (4,), (4,)]),
}


def _inputs(shapes, batch, seed=5):
    rng = np.random.default_rng(seed)
    return [rng.uniform(0.2, 1.2, (batch,) + s) for s in shapes]


def _pack(jac, batch):
    rows = [np.concatenate([b.reshape(batch, -1, int(np.prod(b.shape[1:])) // max(1, int(np.prod(o_shape)))) for b in row], axis=2)
            for row, o_shape in jac]
    return np.concatenate(rows, axis=1)


def _oracle_tile(jaxpr, order, argnums, xs, dtype):
    from oracle import dense_ve
    jac, primal = dense_ve.jacve(jaxpr, order, argnums, xs, dtype)
    batch = xs[0].shape[0]
    rows = []
    for o, row in zip(jaxpr.outvars, jac):
        rows.append(np.concatenate([b.reshape(batch, max(1, o.size), -1) for b in row], axis=2))
    return np.concatenate(rows, axis=1), np.concatenate([p.reshape(batch, -1) for p in primal], axis=1)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_torch_autograd(name):
    import torch
    fn, shapes = CASES[name]
    jaxpr = make_jaxpr(fn, shapes)
    argnums = tuple(range(len(shapes)))
    xs = _inputs(shapes, 2)
    for order in ("fwd", "rev"):
        tile, _ = _oracle_tile(jaxpr, order, argnums, xs, np.float64)
        for b in range(2):
            flat = torch.tensor(np.concatenate([x[b].reshape(-1) for x in xs]), dtype=torch.float64)

            def f(v):
                args, off = [], 0
                for s in shapes:
                    n = int(np.prod(s))
                    args.append(v[off:off + n].reshape(s))
                    off += n
                return torch.cat([o.reshape(-1) for o in torch_eval_jaxpr(jaxpr, args)])
            ref = torch.autograd.functional.jacobian(f, flat).numpy()
            assert np.abs(tile[b] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (name, order)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("order", ["fwd", "rev"])
def test_plan_replay_matches_oracle(name, order):
    from oracle import replay
    fn, shapes = CASES[name]
    jaxpr = make_jaxpr(fn, shapes)
    argnums = tuple(range(len(shapes)))
    plan = build_plan(jaxpr, order, argnums)
    xs = [x.astype(np.float32) for x in _inputs(shapes, 64)]
    jac, primal = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    t64, p64 = _oracle_tile(jaxpr, order, argnums, [x.astype(np.float64) for x in xs], np.float64)
    scale = np.abs(t64).max(axis=(1, 2), keepdims=True) + 1e-30
    assert (np.abs(jac - t64) / scale).max() < 2e-5
    np.testing.assert_allclose(primal, p64, rtol=2e-5, atol=1e-6)


def test_fwd_and_rev_plans_differ_but_agree():
    from oracle import replay
    fn, shapes = CASES["NeuralNetwork"]
    jaxpr = make_jaxpr(fn, shapes)
    fwd, rev = build_plan(jaxpr, "fwd", (1, 2, 3, 4)), build_plan(jaxpr, "rev", (1, 2, 3, 4))
    assert len(fwd.ssa) > 2 * len(rev.ssa)                 # reverse mode is the cheap order for a scalar loss
    xs = [x.astype(np.float32) for x in _inputs(shapes, 32)]
    a = replay.replay(fwd.blob, xs, 1, 76)[0]
    b = replay.replay(rev.blob, xs, 1, 76)[0]
    assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_kernels_match_plan_replay(built_library, name):
    import torch
    from oracle import replay
    fn, shapes = CASES[name]
    argnums = tuple(range(len(shapes)))
    xs = [x.astype(np.float32) for x in _inputs(shapes, 300)]
    exact = name in ("matmul_none",)   # every case contains libm calls (sin/tanh/...): compare within a few ulp
    for order in ("fwd", "rev"):
        jf = gx.jacve(fn, order=order, argnums=argnums, has_aux=True)
        primal, jac = gx.vmap(jf)(*[torch.from_numpy(x).cuda() for x in xs])
        plan = jf.lower(shapes)
        ref_j, ref_p = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
        jaxpr = plan.jaxpr
        outs = jac if len(jaxpr.outvars) > 1 or len(argnums) == 1 else (jac,)
        outs = [r if isinstance(r, tuple) else (r,) for r in outs]
        got = np.concatenate([np.concatenate([b.reshape(300, max(1, o.size), -1).cpu().numpy() for b in row], axis=2)
                              for o, row in zip(jaxpr.outvars, outs)], axis=1)
        scale = np.abs(ref_j).max() + 1e-30
        assert np.abs(got - ref_j).max() <= 2e-6 * scale, (name, order)
        # blocks come back shaped out.shape + in.shape, like jax.jacrev
        for o, row in zip(jaxpr.outvars, outs):
            for a, blk in zip(argnums, row):
                assert tuple(blk.shape) == (300,) + tuple(o.shape) + tuple(shapes[a])


@pytest.mark.gpu
def test_vmap_with_shared_weights(built_library):
    """Reference test_vmap_NeuralNetwork (example_test.py:200-238): x and y batched, weights shared
    (in_axes=(0, None, None, None, None, 0)); the per-sample weight Jacobians summed over the batch are the
    gradient of the summed loss, checked against torch.autograd."""
    import torch
    fn, shapes = CASES["NeuralNetwork"]
    rng = np.random.default_rng(8)
    B = 16
    x, y = rng.normal(size=(B, 4)).astype(np.float32), rng.normal(size=(B, 4)).astype(np.float32)
    W1, b1 = rng.normal(size=(8, 4)).astype(np.float32), rng.normal(size=(8,)).astype(np.float32)
    W2, b2 = rng.normal(size=(4, 8)).astype(np.float32), rng.normal(size=(4,)).astype(np.float32)
    for order in ("fwd", "rev"):
        jf = gx.jacve(fn, order=order, argnums=(1, 2, 3, 4))
        cuda = lambda a: torch.from_numpy(a).cuda()
        jac = gx.vmap(jf, in_axes=(0, None, None, None, None, 0))(cuda(x), cuda(W1), cuda(b1), cuda(W2), cuda(b2), cuda(y))
        assert [tuple(j.shape) for j in jac] == [(B, 8, 4), (B, 8), (B, 4, 8), (B, 4)]
        t = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (W1, b1, W2, b2)]
        xt, yt = torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
        a1 = torch.tanh(xt @ t[0].T + t[1])
        a2 = torch.tanh(a1 @ t[2].T + t[3])
        loss = (0.5 * (a2 - yt) ** 2).sum()
        grads = torch.autograd.grad(loss, t)
        for j, g in zip(jac, grads):
            assert torch.allclose(j.sum(0).cpu().double(), g, rtol=1e-4, atol=1e-5)
