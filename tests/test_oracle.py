"""Pin the oracle: reference known answers, independent fp64 autodiff goldens, einsum identities."""
import json

import numpy as np
import pytest

from graphax_b200.configs import WORKLOADS, get_workload
from graphax_b200.lowering import OPCODES
from graphax_b200.tracer import make_jaxpr
from helpers import GOLDEN, load_random_g, pack_blocks, parity_bound
from oracle import replay, ve_numpy as vo
from oracle.ve_numpy import DenseDimension as D, SparseDimension as S, SparseTensor as ST


# ---- literal known answers of the reference ------------------------------------------------------
def test_count_ops_known_answers_of_g():
    """src/graphax/examples/randoms.py:91  '# fwd: 632, rev: 566, mM: 451'"""
    g, d = load_random_g()
    jaxpr = make_jaxpr(g, [()] * 15)
    rng = np.random.default_rng(7)
    xs = [rng.uniform(0.5, 1.5, 4) for _ in range(15)]
    for order, expect in d["known_answers"].items():
        _, _, aux = vo.jacve(jaxpr, order, range(15), xs, np.float64)
        assert aux["num_muls"] == expect
        assert sum(c for _, c in aux["order_counts"]) == expect


# ---- independent autodiff goldens (tests/golden/make_goldens.py) ----------------------------------
GOLD = json.loads((GOLDEN / "jacobians_fp64.json").read_text())


@pytest.mark.parametrize("name", ["roe3d", "roe1d", "robotarm", "readme", "simple", "lighthouse", "hole", "helmholtz", "zoo"])
def test_oracle_fp64_equals_independent_autodiff(name):
    w = get_workload(name)
    xs, off = [], 0
    for shape, size in zip(w.arg_shapes, w.sizes):
        xs.append(np.asarray(w.base[off:off + size], dtype=np.float64).reshape((1,) + tuple(shape)))
        off += size
    ref = np.asarray(GOLD[name]["jacobian"])
    for key in w.order_names():
        jac, primal, _ = vo.jacve(w.jaxpr(), w.order(key), w.argnums, xs, np.float64)
        J = pack_blocks(jac, 1, w.sizes)[0]
        assert np.abs(J - ref).max() <= 2e-12 * max(1.0, np.abs(ref).max()), (name, key)
        p = np.concatenate([q.reshape(-1) for q in primal])
        np.testing.assert_allclose(p, GOLD[name]["primal"], rtol=1e-13, atol=1e-15)


def test_goldens_match_survey_appendix_d():
    """A few digits printed in SURVEY.md Appendix D (produced by the survey from separate torch code)."""
    r3 = np.asarray(GOLD["roe3d"]["jacobian"])
    assert abs(r3[0, 0] - -7.982679464) < 1e-8 and abs(r3[4, 6] - 10.937184335) < 1e-8
    assert abs(r3[3, 3] - 3.5) < 1e-12 and abs(r3[1, 7]) < 1e-14
    r1 = np.asarray(GOLD["roe1d"]["jacobian"])
    assert abs(r1[2, 5] - -0.017980587) < 1e-8
    ra = np.asarray(GOLD["robotarm"]["jacobian"])
    assert abs(ra[1, 0] - 1167.028649377) < 1e-6 and abs(ra[5, 5] - 1.0) < 1e-12
    assert abs(GOLD["roe3d"]["primal"][0] - 0.210355339) < 1e-8


# ---- fp32 behaviour: order spread as measured in SURVEY Appendix D --------------------------------
@pytest.mark.parametrize("order", ["fwd", "rev", "mixed", "mixed2"])
def test_fp32_oracles_vs_fp64_on_synthetic_batch(order):
    w = WORKLOADS["roe3d"]
    batch = 2000
    xs = w.inputs(batch)
    plan = w.plan(order)
    j64, _, _ = vo.jacve(w.jaxpr(), w.order(order), w.argnums, [x.astype(np.float64) for x in xs], np.float64)
    j32, _, _ = vo.jacve(w.jaxpr(), w.order(order), w.argnums, xs, np.float32)
    t64, t32 = pack_blocks(j64, batch, w.sizes), pack_blocks(j32, batch, w.sizes)
    rep, _ = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    mx = np.abs(t64).max(axis=(1, 2), keepdims=True)
    assert np.isfinite(t64).all() and 10 < mx.max() < 30
    assert (np.abs(t32 - t64) / mx).max() < 2.5e-5          # survey: up to 1.5e-5 for the mixed order
    assert (np.abs(rep - t64) / mx).max() < 2.5e-5
    # plan replay (FMA-contracted) vs reference-order unfused evaluation: the stated parity bound
    assert (np.abs(rep - t32) <= parity_bound(t64)).all()


def test_replay_opcode_table_in_sync():
    assert replay.opcode_names() == OPCODES


# ---- einsum identities of the reference's algebra tests (rank <= 1 cases) -------------------------
RNG = np.random.default_rng(42)
B = 3


def _rand(*shape):
    return RNG.standard_normal((B,) + shape)


def _dense_mul(x, y):
    X, Y = x.dense(), y.dense()
    X = X.reshape(B, *(x.shape if x.shape else ()))
    if x.out is None:
        X = X[:, None]
    if x.primal is None:
        X = X[..., None]
    if y.out is None:
        Y = Y[:, None]
    if y.primal is None:
        Y = Y[..., None]
    return np.einsum("bij,bjk->bik", X.reshape(B, X.shape[1], -1), Y.reshape(B, -1, Y.shape[-1]))


def _classes(n_out, n_mid):
    """Every structure class an edge s<-v can take for |s| = n_out, |v| = n_mid."""
    out = []
    if n_out == n_mid:
        out.append(lambda: ST([S(0, n_mid, 0, 1)], [S(1, n_mid, 0, 0)], _rand(n_mid)))          # DIAG
        out.append(lambda: ST([S(0, n_mid, None, 1)], [S(1, n_mid, None, 0)], _rand()))           # KRON * s
    out.append(lambda: ST([D(0, n_out, 0)], [D(1, n_mid, 1)], _rand(n_out, n_mid)))               # dense
    if n_mid == 1:
        out.append(lambda: ST([D(0, n_out, 0)], [D(1, 1, None)], _rand(n_out)))                   # BCAST
    return out


def test_simple_broadcast_and_matmul_like_reference_tests():
    """tests/core/sparse_mul_test.py:12-23 (diag @ diag) and dense_mul_test.py:11-22 (x @ y)."""
    x, y = _rand(4), _rand(4)
    res = vo._mul(ST([S(0, 4, 0, 1)], [S(1, 4, 0, 0)], x), ST([S(0, 4, 0, 1)], [S(1, 4, 0, 0)], y))
    want = np.einsum("bi,ij,bj->bij", x, np.eye(4), y)
    assert np.allclose(res.dense(), want)
    a, b = _rand(4, 3), _rand(3, 2)
    res = vo._mul(ST([D(0, 4, 0)], [D(1, 3, 1)], a), ST([D(0, 3, 0)], [D(1, 2, 1)], b))
    assert np.allclose(res.val, a @ b)
    assert res.descriptor()["out_dims"] == [["D", 0, 4, 0, None]]


@pytest.mark.parametrize("n_s,n_v,n_p", [(3, 3, 3), (3, 3, 1), (1, 3, 3), (3, 1, 3), (1, 1, 1), (3, 1, 1)])
def test_mul_equals_dense_matmul_for_every_class_pair(n_s, n_v, n_p):
    for mk_l in _classes(n_s, n_v):
        for mk_r in _classes(n_v, n_p):
            lhs, rhs = mk_l(), mk_r()
            got = vo._mul(lhs, rhs)
            want = _dense_mul(lhs, rhs)
            assert np.allclose(got.dense().reshape(want.shape), want), (lhs.descriptor(), rhs.descriptor())


def test_row_and_scalar_edges():
    row = ST([], [D(0, 3, 0)], _rand(3))                    # reduce_sum / dot_general edge
    col = ST([D(0, 3, 0)], [], _rand(3))                    # rank-0 operand broadcast into a vector
    sc = ST((), (), _rand())
    assert np.allclose(vo._mul(row, col).val, np.einsum("bi,bi->b", row.val, col.val))
    outer = vo._mul(col, row)                               # v is rank 0: outer product
    assert np.allclose(outer.val, np.einsum("bi,bj->bij", col.val, row.val))
    assert np.allclose(vo._mul(sc, row).val, sc.val[:, None] * row.val)
    diag = ST([S(0, 3, 0, 1)], [S(1, 3, 0, 0)], _rand(3))
    assert np.allclose(vo._mul(row, diag).val, row.val * diag.val)


@pytest.mark.parametrize("n", [1, 3])
def test_add_equals_dense_add(n):
    for mk_l in _classes(n, n):
        for mk_r in _classes(n, n):
            lhs, rhs = mk_l(), mk_r()
            got = vo._add(lhs, rhs)
            assert np.allclose(got.dense(), lhs.dense() + rhs.dense())
            both_sparse = lhs.out_dims[0].__class__ is S and rhs.out_dims[0].__class__ is S
            assert (got.out_dims[0].__class__ is S) == both_sparse      # S+S stays sparse, else dense


def test_transforms_are_gather_and_zero_pad():
    pre = ST([S(0, 3, 0, 1)], [S(1, 3, 0, 0)], _rand(3))
    t = vo.JacobianTransform("slice", 1, 2, 3)
    assert np.allclose(t.apply(pre).val, pre.dense()[:, 1:2])
    post = ST([D(0, 2, 0)], [D(1, 1, 1)], _rand(2, 1))
    inv = t.apply_inverse(post).val
    assert inv.shape == (B, 2, 3) and np.allclose(inv[:, :, 1:2], post.val) and np.allclose(inv[:, :, [0, 2]], 0)
    c = vo.JacobianTransform("concat", 1, 2, 3)
    k = ST([S(0, 1, None, 1)], [S(1, 1, None, 0)], _rand())
    padded = c.apply(k)
    assert padded.val.shape == (B, 3, 1) and np.allclose(padded.val[:, 1, 0], k.val)
    assert np.allclose(padded.val[:, [0, 2]], 0)
    d3 = ST([S(0, 3, 0, 1)], [S(1, 3, 0, 0)], _rand(3))
    assert np.allclose(c.apply_inverse(d3).val, d3.dense()[:, :, 1:2])
    with pytest.raises(NotImplementedError):              # primitives.py:1186
        c.apply_inverse(ST([S(0, 3, None, 1)], [S(1, 3, None, 0)], _rand()))
