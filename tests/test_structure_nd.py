"""Rank-N structure algebra (graphax_b200/structure_nd.py) against traces of the reference's OWN code.

`tests/golden/make_reference_trace.py` ran the unmodified reference sources (on the NumPy stand-in for JAX) over its
NeuralNetwork / Perceptron / Encoder / EncoderDecoder tests, the batched MLP and `examples/randoms.py:f`, and recorded
every elemental edge's SparseTensor structure, `count_ops` totals and per-vertex counts, every final block's structure
(`sparse_representation=True`) and, for the smaller graphs, every `_mul` / `_add` of the elimination with its operand
and result structures.  The engine must reproduce all of it exactly - ids, `val_dim` numbering and value shapes
included (reference `sparse/tensor.py:334-397, 826-1231, 1267-1493`, `primitives.py:42-107, 243-441, 606-1255`)."""
import gzip
import json

import numpy as np
import pytest

from graphax_b200 import structure_nd as nd
from graphax_b200.elimination import accumulate, build_graph
from helpers import GOLDEN, jaxpr_from_record

TRACE = json.loads(gzip.open(GOLDEN / "reference_trace_rank_n.json.gz").read())["cases"]


def _same(mine: dict, ref: dict) -> bool:
    return mine["out_dims"] == ref["out"] and mine["primal_dims"] == ref["primal"] and mine["val_shape"] == ref["val"]


def _from_descriptor(desc) -> nd.NStruct:
    l = len(desc["out"])
    val = desc["val"]

    def dim(entry, side):
        kind, _id, size, val_dim, other = entry
        ext = None if val_dim is None or val is None else val[val_dim]
        partner = None if kind == "D" else (other - l if side == "out" else other)
        return nd.NDim(kind, size, ext, partner)
    return nd.NStruct(tuple(dim(e, "out") for e in desc["out"]), tuple(dim(e, "primal") for e in desc["primal"]),
                      val is not None)


@pytest.mark.parametrize("name", list(TRACE))
def test_elemental_edges_and_counts_and_blocks_equal_the_reference(name):
    rec = TRACE[name]
    jaxpr = jaxpr_from_record(rec)
    g = build_graph(jaxpr)
    mine = {(s, t): st for s, t, st in g.elemental}
    assert len(mine) == len(rec["elemental"])
    for s, t, desc in rec["elemental"]:
        assert _same(nd.from_struct(mine[(s, t)]).descriptor(), desc), (name, s, t, jaxpr.eqns[t - 1].primitive)
    n_a = len(rec["argnums"])
    for order in ("fwd", "rev"):
        ref = rec["orders"][order]
        acc = accumulate(jaxpr, order, rec["argnums"])
        assert acc.structure_complete
        assert (acc.num_muls, acc.num_adds) == (ref["num_muls"], ref["num_adds"]), (name, order)
        assert [[s.vertex, s.num_muls] for s in acc.steps] == ref["order_counts"], (name, order)
        for i, desc in enumerate(ref["blocks"]):
            s = acc.blocks[(i // n_a, i % n_a)][0]
            if desc is None:
                assert s is None
            else:
                assert _same(nd.from_struct(s).descriptor(), desc), (name, order, i)


@pytest.mark.parametrize("name,order", [(n, o) for n, rec in TRACE.items() for o in ("fwd", "rev")
                                        if "trace" in rec["orders"][o]])
def test_every_traced_product_and_sum(name, order):
    """Each `_mul` / `_add` the reference performed: same product class decision, same result structure, same
    `get_num_muls` / `get_num_adds`."""
    trace = TRACE[name]["orders"][order]["trace"]
    n_mul = n_add = 0
    last = None
    for rec in trace:
        if rec[0] == "mul":
            lhs, rhs = _from_descriptor(rec[1]), _from_descriptor(rec[2])
            got = nd.mul(lhs, rhs)
            assert _same(got.descriptor(), rec[3]), (name, order, rec)
            last = ("mul", lhs, rhs)
            n_mul += 1
        elif rec[0] == "add":
            lhs, rhs = _from_descriptor(rec[1]), _from_descriptor(rec[2])
            got = nd.add(lhs, rhs)
            assert _same(got.descriptor(), rec[3]), (name, order, rec)
            last = ("add", got, rhs)
            n_add += 1
        elif rec[0] == "num_muls":
            assert last[0] == "mul" and nd.num_muls(last[1], last[2]) == rec[1], (name, order, rec)
        elif rec[0] == "num_adds":
            assert last[0] == "add" and nd.num_adds(last[1], last[2]) == rec[1], (name, order, rec)
    assert n_mul > 0


def test_count_ops_and_sparse_representation_through_the_plan():
    """`count_ops` / `sparse_representation` of a plan over matrix-valued vertices now carry the reference's numbers."""
    from graphax_b200.plan import build_plan
    rec = TRACE["Perceptron"]
    plan = build_plan(jaxpr_from_record(rec), "rev", rec["argnums"])
    assert (plan.stats["num_muls"], plan.stats["num_adds"]) == (312, 21)
    assert [list(c) for c in plan.stats["order_counts"]] == rec["orders"]["rev"]["order_counts"]
    blocks = plan.structure["blocks"]
    for i, desc in enumerate(rec["orders"]["rev"]["blocks"]):
        assert _same(blocks[f"{i // len(rec['argnums'])},{i % len(rec['argnums'])}"], desc)


def test_rank_one_descriptors_convert_losslessly():
    """Graphs that mix vector and matrix vertices hand rank <= 1 descriptors (structure.py) to the N-dimensional
    algebra: the conversion must keep kind, size, value axis and pending transforms."""
    from graphax_b200 import structure as st
    for s in (st.SCALAR, st.diag(3), st.kron_scalar(4), st.bcast(3, 1), st.rank0_operand(5), st.row(3)):
        assert nd.from_struct(s).descriptor()["out_dims"] == s.descriptor()["out_dims"]
        assert nd.from_struct(s).descriptor()["primal_dims"] == s.descriptor()["primal_dims"]
        assert nd.from_struct(s).descriptor()["val_shape"] == s.descriptor()["val_shape"]
    t = nd.from_struct(st.transform_edge(st.Transform("concat", 1, 4, 5)))
    assert not t.has_val and t.pre[0].kind == "concat" and (t.pre[0].lo, t.pre[0].hi, t.pre[0].out_shape) == (1, 4, (5,))
