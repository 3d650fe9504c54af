"""Seeded random functions (scalars and 3-vectors; elementwise rules, reductions, slices, `where`) and random
elimination orders: the engine's plan (C replay, fp32) against torch.autograd on the traced equations (fp64).
Plus the degenerate graphs a fuzzer found first: outputs that are bare shape operations of the inputs, where the
reference itself raises or returns a wrongly shaped block."""
import numpy as np
import pytest

from graphax_b200.plan import build_plan
from graphax_b200.tracer import jnp, make_jaxpr
from helpers import torch_eval_jaxpr

UNARY = ["sin", "cos", "tanh", "exp", "sqrt_abs", "square", "negative", "arctan", "log1p_sq", "sigmoid"]
BINARY = ["add", "subtract", "multiply", "divide_safe", "maximum", "minimum", "arctan2"]


def _random_function(seed: int, n_ops: int):
    def fn(*xs):
        r = np.random.default_rng(seed)
        vals = list(xs)
        for _ in range(n_ops):
            kind = r.choice(["un", "bi", "bi", "sum", "slice", "where"])
            a = vals[r.integers(len(vals))]
            b = vals[r.integers(len(vals))]
            if a.ndim and b.ndim and a.shape != b.shape:
                b = a * 0.75 + 0.1
            if kind == "un":
                u = UNARY[r.integers(len(UNARY))]
                v = jnp.sqrt(jnp.abs(a) + 0.5) if u == "sqrt_abs" else jnp.log1p(a * a) if u == "log1p_sq" else \
                    jnp.square(a) if u == "square" else getattr(jnp, u)(a * 0.5)
            elif kind == "bi":
                o = BINARY[r.integers(len(BINARY))]
                v = a / (jnp.square(b) + 1.) if o == "divide_safe" else jnp.arctan2(a, jnp.square(b) + 0.5) if o == "arctan2" \
                    else getattr(jnp, o)(a, b)
            elif kind == "sum":
                v = jnp.sum(a) if a.ndim else a * 2.
            elif kind == "slice":
                v = a[0:2] if a.ndim and a.shape[0] >= 3 else a + 1.
            else:
                v = jnp.where(a > b, a * b, a - b)
            vals.append(v)
        return tuple(vals[-3:])
    return fn


def _autograd(jaxpr, xs, b):
    import torch
    args = [torch.tensor(x[b], dtype=torch.float64, requires_grad=True) for x in xs]
    out = torch.cat([o.reshape(-1) for o in torch_eval_jaxpr(jaxpr, args)])
    rows = []
    for r in range(out.numel()):
        g = torch.autograd.grad(out[r], args, retain_graph=True, allow_unused=True) if out[r].requires_grad else [None] * len(args)
        rows.append(torch.cat([(torch.zeros_like(a) if gi is None else gi).reshape(-1) for gi, a in zip(g, args)]))
    return torch.stack(rows).numpy()


@pytest.mark.parametrize("seed", range(24))
def test_random_function_random_order(seed):
    from oracle import replay
    rng = np.random.default_rng(1000 + seed)
    shapes = [(), (3,), (3,), ()][:rng.integers(2, 5)]
    jaxpr = make_jaxpr(_random_function(seed, int(rng.integers(8, 28))), shapes)
    outs = {o.eqn_index for o in jaxpr.outvars}
    eliminable = [i + 1 for i in range(len(jaxpr.eqns)) if i not in outs]
    xs = [rng.uniform(0.3, 1.5, (8,) + s).astype(np.float32) for s in shapes]
    for order in ("fwd", "rev", [int(v) for v in rng.permutation(eliminable)]):
        try:
            plan = build_plan(jaxpr, order, tuple(range(len(shapes))))
        except ValueError:        # an output that is consumed later must stay in a custom order: not this test's subject
            assert not isinstance(order, str)
            continue
        rep, _ = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
        for b in range(3):
            ref = _autograd(jaxpr, xs, b)
            if not np.isfinite(rep[b]).all():
                continue          # |x| at exactly 0: the reference's rule is x / |x| (primitives.py:151)
            assert np.abs(rep[b] - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max()), (seed, order)


@pytest.mark.parametrize("case", ["slice", "slice_and_concat", "concat_same"])
def test_outputs_that_are_bare_shape_operations(case):
    """f(x) = x[0:2] and friends: the Jacobian block is a selection matrix with no value attached to the edge.  The
    reference raises (or returns a 3-vector for the first); the engine keeps the exact entries and only gives up the
    SparseTensor descriptor."""
    from oracle import replay
    fn, shapes = {"slice": (lambda x: x[0:2], [(3,)]),
                  "slice_and_concat": (lambda x, y: (x[0:2] * 1.0, jnp.concatenate([x, y])), [(3,), (2,)]),
                  "concat_same": (lambda x: jnp.concatenate([x, x]), [(3,)])}[case]
    jaxpr = make_jaxpr(fn, shapes)
    xs = [np.arange(1, 1 + int(np.prod(s)), dtype=np.float32).reshape((1,) + s) for s in shapes]
    for order in ("fwd", "rev"):
        plan = build_plan(jaxpr, order, tuple(range(len(shapes))))
        rep, _ = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
        assert np.array_equal(rep[0], _autograd(jaxpr, xs, 0))


# ---- matrices: matmul, transposes, axis reductions, reshapes, N-D slices / concatenations -----------------------------
def _random_matrix_function(seed, n_ops):
    def fn(*xs):
        r = np.random.default_rng(seed)
        vals = list(xs)
        for _ in range(n_ops):
            kind = r.choice(['un','bi','matmul','T','sum','max','reshape','where','slice','concat','expand','squeeze'])
            a = vals[r.integers(len(vals))]
            b = vals[r.integers(len(vals))]
            v = None
            if kind == 'un': v = [jnp.tanh, jnp.sin, lambda t: jnp.exp(t*0.3), lambda t: jnp.sqrt(jnp.square(t)+1.)][r.integers(4)](a)
            elif kind == 'bi':
                if a.shape == b.shape: v = [lambda p,q:p+q, lambda p,q:p*q, lambda p,q:p-q, lambda p,q: p/(jnp.square(q)+1.), jnp.maximum][r.integers(5)](a,b)
                elif b.ndim == 0: v = a * b
                elif a.ndim == 2 and b.ndim == 1 and a.shape[1] == b.shape[0]: v = a + b
            elif kind == 'matmul':
                if a.ndim >= 1 and b.ndim >= 1 and a.ndim <= 2 and b.ndim <= 2 and a.shape[-1] == b.shape[0]: v = a @ b
            elif kind == 'T':
                if a.ndim == 2: v = a.T
            elif kind == 'sum':
                if a.ndim >= 1: v = jnp.sum(a, axis=int(r.integers(a.ndim)))
            elif kind == 'max':
                if a.ndim >= 1: v = jnp.max(a, axis=int(r.integers(a.ndim)))
            elif kind == 'reshape':
                if a.ndim == 2: v = a.reshape(a.shape[0]*a.shape[1])
                elif a.ndim == 1 and a.shape[0] % 2 == 0: v = a.reshape(2, a.shape[0]//2)
            elif kind == 'where':
                if a.shape == b.shape: v = jnp.where(a > b * 0.9 + 0.05, a * b, a - 2. * b)
            elif kind == 'slice':
                if a.ndim == 2 and a.shape[0] >= 2: v = a[1:, :]
                elif a.ndim == 1 and a.shape[0] >= 2: v = a[:-1]
            elif kind == 'concat':
                if a.ndim == b.ndim == 1: v = jnp.concatenate([a, b])
                elif a.ndim == b.ndim == 2 and a.shape[1] == b.shape[1]: v = jnp.concatenate([a, b], axis=0)
            elif kind == 'expand':
                if a.ndim == 1: v = jnp.expand_dims(a, 0)
            elif kind == 'squeeze':
                if a.ndim == 2 and 1 in a.shape: v = jnp.squeeze(a)
            if v is not None and v.size <= 40: vals.append(v)
        outs = [v for v in vals[len(xs):]][-2:]
        return tuple(outs)
    return fn


@pytest.mark.parametrize("seed", range(16))
def test_random_matrix_function_random_order(seed):
    from oracle import replay
    rng = np.random.default_rng(5000 + seed)
    shapes = [(3, 4), (4,), (4, 3), ()][:rng.integers(2, 5)]
    jaxpr = make_jaxpr(_random_matrix_function(seed, int(rng.integers(6, 22))), shapes)
    outs = {getattr(o, "eqn_index", None) for o in jaxpr.outvars}
    if not jaxpr.outvars or None in outs or len(outs) < len(jaxpr.outvars):
        pytest.skip("degenerate draw: an input or the same value returned twice")
    eliminable = [i + 1 for i in range(len(jaxpr.eqns)) if i not in outs]
    xs = [rng.uniform(0.3, 1.5, (4,) + s).astype(np.float32) for s in shapes]
    for order in ("fwd", "rev", [int(v) for v in rng.permutation(eliminable)]):
        try:
            plan = build_plan(jaxpr, order, tuple(range(len(shapes))))
        except ValueError:
            assert not isinstance(order, str)
            continue
        rep, _ = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
        for b in range(2):
            ref = _autograd(jaxpr, xs, b)
            assert np.isfinite(rep[b]).all()
            assert np.abs(rep[b] - ref).max() <= 3e-4 * max(1.0, np.abs(ref).max()), (seed, order)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,seed", [("vector", s) for s in range(10)] + [("matrix", s) for s in range(8)])
def test_random_plans_on_the_gpu(built_library, kind, seed):
    """The same random plans through the run-time specialised and the interpreter kernels against the C replay
    (libm functions differ by a few ulp between CUDA and glibc)."""
    from helpers import run_engine
    from oracle import replay
    if kind == "vector":
        rng = np.random.default_rng(1000 + seed)
        shapes = [(), (3,), (3,), ()][:rng.integers(2, 5)]
        jaxpr = make_jaxpr(_random_function(seed, int(rng.integers(8, 28))), shapes)
    else:
        rng = np.random.default_rng(5000 + seed)
        shapes = [(3, 4), (4,), (4, 3), ()][:rng.integers(2, 5)]
        jaxpr = make_jaxpr(_random_matrix_function(seed, int(rng.integers(6, 22))), shapes)
    outs = {getattr(o, "eqn_index", None) for o in jaxpr.outvars}
    if not jaxpr.outvars or None in outs or len(outs) < len(jaxpr.outvars):
        pytest.skip("degenerate draw")
    xs = [rng.uniform(0.3, 1.5, (300,) + s).astype(np.float32) for s in shapes]
    plan = build_plan(jaxpr, "rev", tuple(range(len(shapes))))
    ref_j, ref_p = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    ok = np.isfinite(ref_j).all(axis=(1, 2))
    scale = np.abs(ref_j).max(axis=(1, 2), keepdims=True) + 1e-30
    for kernel in ("jit", "interpreter"):
        jac, primal, used = run_engine(plan, xs, kernel)
        assert used == kernel
        assert (np.abs(jac - ref_j) / scale)[ok].max() <= 1e-5, (kind, seed, kernel)
        np.testing.assert_allclose(primal[ok], ref_p[ok], rtol=1e-5, atol=1e-6)
