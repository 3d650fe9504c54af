"""The tracer must number vertices the way jax.make_jaxpr would."""
import hashlib

import pytest

from graphax_b200.configs import WORKLOADS
from graphax_b200.elimination import build_graph, checkify_order, eliminable_vertices
from graphax_b200.examples import orders as ref
from graphax_b200.tracer import jnp, make_jaxpr
from helpers import load_random_g


def test_equation_counts_and_output_vertices():
    j = WORKLOADS["roe1d"].jaxpr()
    assert len(j.eqns) == 101 and j.output_vertices == [97, 99, 101]
    j = WORKLOADS["robotarm"].jaxpr()
    assert len(j.eqns) == 114 and j.output_vertices == [85, 102, 114, 62, 66, 68]
    j = WORKLOADS["roe3d"].jaxpr()
    assert len(j.eqns) == 138 and j.output_vertices == [134, 136, 138]
    g, d = load_random_g()
    assert len(make_jaxpr(g, [()] * 15).eqns) == 105


def test_vmap_numbering_of_roeflux_3d():
    j = WORKLOADS["roe3d"].jaxpr()
    mapping, pads = j.to_vmap_numbering()
    assert pads == [24, 33, 48, 55, 59, 82, 91, 133]
    assert len(j.eqns) + len(pads) == 146
    assert [mapping[v] for v in j.output_vertices] == [142, 144, 146]


def test_reference_orders_are_exactly_the_eliminable_vertex_sets():
    """A custom order must contain every non-output vertex (core.py:304-310): the ids missing from each
    reference order are therefore the output vertices - an anchor for the numbering."""
    cases = [("roe1d", ref.ROEFLUX_1D_ALPHAGRAD), ("roe1d", ref.ROEFLUX_1D_MARKOWITZ),
             ("robotarm", ref.ROBOTARM_ALPHAGRAD), ("robotarm", ref.ROBOTARM_MARKOWITZ)]
    for name, order in cases:
        g = build_graph(WORKLOADS[name].jaxpr())
        assert sorted(order) == eliminable_vertices(g)
        assert checkify_order(order, g) == list(order)
    j = WORKLOADS["roe3d"].jaxpr()
    mapping, pads = j.to_vmap_numbering()
    for order in (ref.ROEFLUX_3D_VMAP_MIXED, ref.ROEFLUX_3D_VMAP_MIXED2):
        assert sorted(order) == [v for v in range(1, 147) if v not in (142, 144, 146)]
        translated = j.translate_vmap_order(order)
        assert sorted(translated) == eliminable_vertices(build_graph(j))


def test_mixed_order_translation_hashes():
    """sha1 prefixes recorded in SURVEY.md Appendix A.4."""
    j = WORKLOADS["roe3d"].jaxpr()
    assert hashlib.sha1(str(list(ref.ROEFLUX_3D_VMAP_MIXED)).encode()).hexdigest().startswith("61dd23bf18f4")
    assert hashlib.sha1(str(j.translate_vmap_order(ref.ROEFLUX_3D_VMAP_MIXED)).encode()).hexdigest() \
        .startswith("123e7a1b1d29")
    assert hashlib.sha1(str(j.translate_vmap_order(ref.ROEFLUX_3D_VMAP_MIXED2)).encode()).hexdigest() \
        .startswith("b3d294f72f5b")


def test_lowering_rules():
    def f(x, y):
        a = .5 * x
        b = x * .5
        c = jnp.sqrt(y) + jnp.sqrt(y)          # no CSE: two sqrt equations
        d = x ** 2
        e = -d
        return (a + b) * c / e, jnp.sum(y ** 2), jnp.dot(y, y), y[1:]
    j = make_jaxpr(f, [(), (3,)])
    prims = [e.primitive for e in j.eqns]
    assert prims == ["mul", "mul", "sqrt", "sqrt", "add", "integer_pow", "neg", "add", "mul", "div",
                     "integer_pow", "reduce_sum", "dot_general", "slice"]
    assert repr(j.eqns[0].invars[0]) == "0.5" and repr(j.eqns[1].invars[1]) == "0.5"   # literal keeps its side
    assert j.eqns[-1].params["start_indices"] == (1,) and j.eqns[-1].outvar.shape == (2,)


def test_order_validation_errors():
    g = build_graph(WORKLOADS["readme"].jaxpr())
    assert checkify_order("fwd", g) == [1, 2, 3, 4] and checkify_order("reverse", g) == [4, 3, 2, 1]
    with pytest.raises(ValueError, match="not a valid order identifier"):
        checkify_order("sideways", g)
    with pytest.raises(ValueError, match="missing vertices"):
        checkify_order([1, 2, 3], g)
    with pytest.raises(ValueError):
        checkify_order([1, 2, 3, 4, 5], g)
    with pytest.raises(ValueError):
        checkify_order([1, 2, 3, 4, 4], g)


def test_unknown_primitive_raises_not_implemented():
    from graphax_b200.plan import build_plan

    def f(x):
        return x.trace.emit("cumsum", (x.var,), x.shape)
    with pytest.raises(NotImplementedError, match="does not have registered elemental partial"):
        build_plan(make_jaxpr(f, [(3,)]), "rev", (0,))


def test_nn_helpers_lower_to_the_expected_primitives():
    """softmax (without jax.nn's custom_jvp wrapper), mean, keepdims reductions, split and stop_gradient."""
    from graphax_b200.tracer import jnn, jnp, lax, make_jaxpr
    j = make_jaxpr(lambda a: jnn.softmax(a, axis=1), [(3, 4)])
    assert [e.primitive for e in j.eqns] == ["reduce_max", "broadcast_in_dim", "stop_gradient", "sub", "exp",
                                             "reduce_sum", "broadcast_in_dim", "div"]
    assert j.outvars[0].shape == (3, 4)
    j = make_jaxpr(lambda a: jnp.mean(a, axis=-1), [(3, 4)])
    assert [e.primitive for e in j.eqns] == ["reduce_sum", "div"] and j.outvars[0].shape == (3,)
    j = make_jaxpr(lambda a: jnp.split(a, [2, 4], axis=0), [(6, 3)])
    assert [e.primitive for e in j.eqns] == ["slice"] * 3 and [o.shape for o in j.outvars] == [(2, 3)] * 3
    j = make_jaxpr(lambda a: jnp.sqrt(2.) * lax.stop_gradient(a), [(2,)])
    # omnistaging: an operation on constants inside a traced function is an equation too (`sqrt 2.0`)
    assert [e.primitive for e in j.eqns] == ["sqrt", "stop_gradient", "mul"]


def test_constant_fills_and_closed_over_arrays():
    """jnp.ones / full / zeros_like / ones_like stage one broadcast_in_dim of a literal (a vertex without in-edges, like
    jnp.zeros); NumPy operands and jnp.asarray values become constants of the trace, one per array object, numbered
    after the arguments in order of first use."""
    import numpy as np
    from graphax_b200.plan import build_plan
    from graphax_b200.tracer import jnp, make_jaxpr
    from oracle import replay
    table = np.array([1.0, -2.0, 0.5], np.float32)
    offsets = [0.25, 0.5, 0.75]

    def f(x):
        late = jnp.asarray(offsets)                       # registered when first used, not here
        y = x * jnp.full((3,), 2.5) + jnp.ones_like(x) * x * table - jnp.zeros_like(x)
        return jnp.sum(y * table) + jnp.sum(late * x), jnp.ones((2, 3)) * x

    jaxpr = make_jaxpr(f, [(3,)], "native")
    assert [e.primitive for e in jaxpr.eqns][:3] == ["broadcast_in_dim", "mul", "broadcast_in_dim"]
    assert [v.shape for v in jaxpr.constvars] == [(3,), (3,)] and [v.input_index for v in jaxpr.constvars] == [1, 2]
    assert np.array_equal(jaxpr.consts[0], table) and np.array_equal(jaxpr.consts[1], np.float32(offsets))
    x = np.array([[0.1, 0.2, 0.3], [1.0, -1.0, 2.0]], np.float32)
    for order in ("fwd", "rev"):
        plan = build_plan(jaxpr, order, (0,))
        assert plan.in_sizes == [3]                       # constants are not kernel inputs
        jac, primal = replay.replay(plan.blob, [x], plan.n_rows, plan.n_cols)
        want = np.broadcast_to((2.5 + table) * table + np.float32(offsets), (2, 3))
        np.testing.assert_allclose(jac[:, 0], want, rtol=1e-6)
        np.testing.assert_allclose(jac[:, 1:].reshape(2, 2, 3, 3), np.broadcast_to(np.eye(3, dtype=np.float32), (2, 2, 3, 3)))
    with pytest.raises(TypeError):
        make_jaxpr(lambda x: x * "three", [()], "native")
