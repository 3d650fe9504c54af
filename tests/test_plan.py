"""Plan builder (host logic, no GPU): structure parity with the oracle, count_ops, blob, scheduling."""
import re
import struct
from pathlib import Path

import os

import numpy as np
import pytest

from graphax_b200 import structure as st
from graphax_b200.configs import EXTRA_WORKLOADS, WORKLOADS, get_workload, named_plans
from graphax_b200.lowering import OPCODES, Dag, _f32_from_fraction
from graphax_b200.plan import MAGIC, build_plan, fnv1a64
from graphax_b200.tracer import make_jaxpr
from helpers import load_random_g, pack_blocks
from oracle import replay, ve_numpy as vo

ROOT = Path(__file__).resolve().parent.parent
CASES = [(w.name, k) for w in list(WORKLOADS.values()) + list(EXTRA_WORKLOADS.values()) for k in w.order_names()]


@pytest.mark.parametrize("name,order", CASES)
def test_edge_structure_and_counts_match_oracle_exactly(name, order):
    """The SparseTensor structure of every elemental edge and every final Jacobian block, and the
    reference's count_ops numbers, must be identical between the engine's symbolic elimination and
    the oracle's value-carrying restatement ("plan and edge structure bit-exact")."""
    w = get_workload(name)
    plan = w.plan(order)
    _, _, aux = vo.jacve(w.jaxpr(), w.order(order), w.argnums, w.inputs(2), np.float32)
    assert plan.structure["order"] == aux["structure"]["order"]
    assert plan.structure["elemental"] == aux["structure"]["elemental"]
    assert plan.structure["blocks"] == aux["structure"]["blocks"]
    assert plan.stats["num_muls"] == aux["num_muls"] and plan.stats["num_adds"] == aux["num_adds"]
    assert [tuple(x) for x in plan.stats["order_counts"]] == [tuple(x) for x in aux["order_counts"]]


def test_count_ops_known_answers():
    """randoms.py:91 (632 / 566) through the engine's plan builder; scalar-graph counts of SURVEY App. C."""
    g, d = load_random_g()
    jaxpr = make_jaxpr(g, [()] * 15)
    for order, expect in d["known_answers"].items():
        assert build_plan(jaxpr, order, range(15)).stats["num_muls"] == expect
    expect = {("roe1d", "fwd"): (620, 224), ("roe1d", "rev"): (364, 135), ("roe1d", "alphagrad"): (353, 143),
              ("roe1d", "markowitz"): (407, 164), ("robotarm", "fwd"): (397, 79), ("robotarm", "rev"): (301, 84),
              ("robotarm", "alphagrad"): (254, 64), ("robotarm", "markowitz"): (288, 64)}
    for (name, order), (muls, adds) in expect.items():
        s = WORKLOADS[name].plan(order).stats
        assert (s["num_muls"], s["num_adds"]) == (muls, adds)


def test_elemental_edge_classes_of_roeflux_3d():
    """the structure class each kind of equation produces."""
    plan = WORKLOADS["roe3d"].plan("rev")
    by_edge = {(s, d): desc for s, d, desc in plan.structure["elemental"]}
    ul0, ul = -1, -2
    assert by_edge[(ul, 7)]["out_dims"] == [["S", 0, 3, None, 1]] and by_edge[(ul, 7)]["val_shape"] == []   # KRON*s
    assert by_edge[(ul0, 7)]["out_dims"] == [["D", 0, 3, 0, None]]                                          # BCAST
    assert by_edge[(ul0, 7)]["primal_dims"] == [["D", 1, 1, None, None]]
    assert by_edge[(ul0, 9)]["out_dims"] == [["S", 0, 1, 0, 1]] and by_edge[(ul0, 9)]["val_shape"] == [1]   # DIAG (unary)
    assert by_edge[(21, 22)] == {"out_dims": [], "primal_dims": [["D", 0, 3, 0, None]], "val_shape": [3],
                                 "pre_transforms": []}                                                     # ROW
    assert by_edge[(23, 24)]["out_dims"] == [["D", 0, 1, 0, None]] and by_edge[(23, 24)]["primal_dims"] == []  # RANK0
    assert by_edge[(2, 3)]["val_shape"] is None and by_edge[(2, 3)]["pre_transforms"] == [["slice", 0, 1, 3]]
    assert by_edge[(26, 77)]["pre_transforms"] == [["concat", 0, 1, 3]]
    assert (76, 77) not in by_edge or True      # zeros literal: the concatenate operand edge exists, the zeros vertex has no in-edge
    assert not any(d == 76 for _, d, _ in plan.structure["elemental"])


def test_structure_algebra_rules():
    d3, k3 = st.diag(3), st.kron_scalar(3)
    dense = st.Struct(st.Dim("D", 3, 0), st.Dim("D", 3, 1), (3, 3))
    assert st.mul(d3, d3) == d3 and st.mul(k3, k3) == k3 and st.mul(k3, d3) == d3
    assert st.mul(d3, dense) == dense and st.mul(dense, d3) == dense                   # S.D / D.S -> D
    assert st.num_muls(d3, d3) == 3 and st.num_muls(k3, k3) == 1 and st.num_muls(dense, dense) == 27
    assert st.num_muls(d3, dense) == 9 and st.num_muls(st.SCALAR, st.SCALAR) == 1
    ssum = st.add(k3, k3)
    assert ssum.out.kind == "S" and ssum.out.val_dim == 0 and ssum.val_shape == (1,)    # S+S stays sparse
    assert st.add(d3, dense) == dense and st.num_adds(dense, d3) == 9 and st.num_adds(ssum, k3) == 1
    row, col = st.row(3), st.rank0_operand(3)
    assert st.mul(row, col) == st.SCALAR and st.mul_class(row, col) == "dot"
    assert st.mul(col, row).val_shape == (3, 3) and st.mul_class(col, st.SCALAR) == "dot"
    t = st.Transform("concat", 0, 1, 3)
    with pytest.raises(NotImplementedError, match="Finish the implementation"):
        st.apply_inverse_transform(t, k3)


def test_blob_layout_and_determinism():
    w = WORKLOADS["roe3d"]
    a, b = build_plan(w.jaxpr(), "rev", w.argnums), build_plan(make_jaxpr(w.fun, w.arg_shapes), "rev", w.argnums)
    assert a.blob == b.blob and a.key == b.key == f"{fnv1a64(a.blob):016x}"
    h = struct.unpack("<16I", a.blob[:64])
    assert h[0] == MAGIC and h[1] == 1 and h[2] == len(a.blob)
    assert (h[3], h[4], h[5], h[6], h[7]) == (6, 3, 10, 10, 5)
    assert h[8] == len(a.ssa) and h[9] == a.n_slots and a.bytes_per_jacobian == 240
    assert WORKLOADS["roe1d"].plan("rev").bytes_per_jacobian == 96
    assert WORKLOADS["robotarm"].plan("rev").bytes_per_jacobian == 168
    stores = sorted(op.dst for op in a.ssa if op.op == "store_jac")
    assert stores == list(range(50))                     # every tile entry written exactly once
    assert len({p.key for _, p in named_plans()}) >= 24   # plans without divisions coincide across modes


def test_opcode_tables_in_sync():
    header = (ROOT / "graphax_b200" / "csrc" / "gx_ops.h").read_text()
    names = re.findall(r"X\(\w+, (\w+)\)", header.split("#define GX_OPCODE_LIST(X)")[1].split("enum gx_opcode")[0])
    assert names == OPCODES == replay.opcode_names()


@pytest.mark.parametrize("name,order", [("roe3d", "rev"), ("roe3d", "mixed"), ("roe1d", "alphagrad")])
def test_schedule_and_division_variants_agree(name, order):
    """Any topological order of the plan computes the same bits; the reciprocal lowering stays within
    a few ulp of the verbatim (three-division) formulas."""
    w = WORKLOADS[name]
    xs = w.inputs(512)
    out = {}
    for division in ("reciprocal", "ieee"):
        for schedule in ("pressure", "elimination"):
            p = build_plan(w.jaxpr(), w.order(order), w.argnums, division=division, schedule=schedule)
            out[(division, schedule)] = replay.replay(p.blob, xs, p.n_rows, p.n_cols)[0]
    for division in ("reciprocal", "ieee"):
        assert np.array_equal(out[(division, "pressure")].view(np.uint32), out[(division, "elimination")].view(np.uint32))
    a, b = out[("reciprocal", "pressure")], out[("ieee", "pressure")]
    mx = np.abs(b).max(axis=(1, 2), keepdims=True)
    assert (np.abs(a - b) / mx).max() < 1e-5     # fp32 conditioning of the flux amplifies single-ulp differences
    p_press = build_plan(w.jaxpr(), w.order(order), w.argnums, schedule="pressure")
    p_elim = build_plan(w.jaxpr(), w.order(order), w.argnums, schedule="elimination")
    assert p_press.n_slots <= p_elim.n_slots


def test_jittered_schedules_compute_the_same_bits():
    """schedule='minpressure:<seed>:<sigma>': another statement order of about the same pressure, the same results."""
    w = WORKLOADS["roe3d"]
    xs = w.inputs(256)
    base = build_plan(w.jaxpr(), w.order("mixed"), w.argnums, rounding="reference", schedule="minpressure")
    ref = replay.replay(base.blob, xs, base.n_rows, base.n_cols)
    keys = {base.key}
    for schedule in ("minpressure:1:8", "minpressure:2:30"):
        p = build_plan(w.jaxpr(), w.order("mixed"), w.argnums, rounding="reference", schedule=schedule)
        got = replay.replay(p.blob, xs, p.n_rows, p.n_cols)
        assert np.array_equal(got[0].view(np.uint32), ref[0].view(np.uint32))
        assert np.array_equal(got[1].view(np.uint32), ref[1].view(np.uint32))
        assert sorted(o.op for o in p.ssa) == sorted(o.op for o in base.ssa)
        assert abs(p.n_slots - base.n_slots) <= 12
        keys.add(p.key)
    assert len(keys) == 3                                    # different orders, different blobs
    with pytest.raises(ValueError):
        build_plan(w.jaxpr(), "rev", w.argnums, schedule="random")


def test_reference_helper_functions():
    """get_valid_vertices / get_output_vertices (reference utils.py:8-36) and sparse_tensor_zeros_like
    (sparse/tensor.py:190-201, exported at graphax/__init__.py:4)."""
    import graphax_b200 as gx
    from graphax_b200.examples import RoeFlux_1d, RoeFlux_3d
    w1, w3 = WORKLOADS["roe1d"], WORKLOADS["roe3d"]
    valid = gx.get_valid_vertices(RoeFlux_1d, w1.arg_shapes)
    assert sorted(set(range(1, 102)) - set(valid)) == [97, 99, 101]          # the three pure outputs
    assert valid == sorted(valid) and gx.get_output_vertices(RoeFlux_1d, w1.arg_shapes) == set()
    assert len(gx.get_valid_vertices(RoeFlux_3d, w3.arg_shapes)) == 135
    # an output that feeds another vertex is an output vertex AND eliminable (core.py:404-409)
    def chained(x):
        y = gx.jnp.sin(x)
        return y, gx.jnp.cos(y)
    assert gx.get_output_vertices(chained, [()]) == {1}
    assert gx.get_valid_vertices(chained, [()]) == [1]
    st = gx.SparseJacobian([{"kind": "dense", "size": 3}], [{"kind": "dense", "size": 2}], [3, 2], np.ones((3, 2), np.float32))
    z = gx.sparse_tensor_zeros_like(st)
    assert z.out_dims == st.out_dims and z.primal_dims == st.primal_dims and z.val_shape == [3, 2]
    assert z.dense.shape == (3, 2) and not z.dense.any() and st.dense.all()


def test_input_staging_is_chosen_per_plan(monkeypatch):
    """tuning.json `input_staged: 0` / `prefetch_at` (the RobotArm_6DOF plans: measured faster without the
    shared-memory input stage, the next tile's loads issued early in the body) reach the generated source of the
    ahead-of-time kernels; GX_INPUT_STAGED / GX_PREFETCH_AT override."""
    from graphax_b200 import codegen
    monkeypatch.delenv("GX_INPUT_STAGED", raising=False)
    monkeypatch.delenv("GX_PREFETCH_AT", raising=False)
    direct = WORKLOADS["robotarm"].plan("markowitz", rounding="reference")
    staged = WORKLOADS["roe3d"].plan("mixed", rounding="reference")
    assert codegen.input_stages(direct, 128) == 0 and codegen.input_stages(staged, 128) == 2
    assert "#define GX_IN_STAGES 0" in codegen.emit_source(direct, nvrtc=False)
    assert "#define GX_IN_STAGES 2" in codegen.emit_source(staged, nvrtc=False)
    assert codegen.input_stages(direct, 128, samples_per_thread=2) == 2      # the packed kernels always stage
    assert codegen.prefetch_position(direct) == int(0.2 * len(direct.ssa))   # tuning.json prefetch_at
    assert 0.45 * len(staged.ssa) <= codegen.prefetch_position(staged) <= 0.75 * len(staged.ssa)   # the automatic rule
    monkeypatch.setenv("GX_PREFETCH_AT", "0.5")
    assert codegen.prefetch_position(direct) == int(0.5 * len(direct.ssa))
    monkeypatch.setenv("GX_INPUT_STAGED", "1")
    assert codegen.input_stages(direct, 128) == 2
    monkeypatch.setenv("GX_INPUT_STAGED", "0")
    assert codegen.input_stages(staged, 128) == 0


def test_argnums_subset_prunes_dead_work():
    w = WORKLOADS["roe3d"]
    full = build_plan(w.jaxpr(), "rev", w.argnums)
    sub = build_plan(w.jaxpr(), "rev", (1,))
    assert sub.n_cols == 3 and len(sub.ssa) < len(full.ssa)
    assert sub.stats["num_muls"] == full.stats["num_muls"]      # count_ops is unpruned, like the reference (core.py:524)
    xs = w.inputs(64)
    jf = replay.replay(full.blob, xs, 5, 10)[0]
    js = replay.replay(sub.blob, xs, 5, 3)[0]
    assert np.array_equal(js.view(np.uint32), jf[:, :, 1:4].view(np.uint32))


def test_constant_folding_is_exactly_rounded():
    from fractions import Fraction
    dag = Dag()
    a, b, c = dag.const(0.1), dag.const(3.0), dag.const(-0.3)
    n = dag.fma(a, b, c)
    exact = Fraction(float(np.float32(0.1))) * 3 + Fraction(float(np.float32(-0.3)))
    assert dag.op[n] == "const" and dag.cval[n] == _f32_from_fraction(exact)
    x = dag.input(0)
    assert dag.mul(x, dag.const(1.0)) == x and dag.op[dag.mul(x, dag.const(-1.0))] == "neg"
    assert dag.fma(x, dag.const(1.0), a) == dag.add(x, a)
    assert dag.mul(x, a) == dag.mul(a, x)                       # hash-consing of commutative ops


def test_sum_contraction_is_opt_in_and_stays_close():
    """plan._contract_sums: fewer operations, same values up to the association of the edge sums.  Against
    fp64 the contracted plan must be at least as accurate (rms) as the reference-order plan; against the
    reference-order plan it may differ by up to 1.5x the parity bound on the ill-conditioned mixed order
    (two independent fp32 roundings), which is why it is not the default."""
    from helpers import pack_blocks, parity_bound
    from oracle import ve_numpy
    w = WORKLOADS["roe3d"]
    xs = w.inputs(2000)
    for order in w.order_names():
        strict = build_plan(w.jaxpr(), w.order(order), w.argnums)
        fused = build_plan(w.jaxpr(), w.order(order), w.argnums, contract=True)
        assert strict.key == w.plan(order).key                                   # default = reference association
        n = lambda p: sum(1 for o in p.ssa if o.op not in ("neg", "store_jac", "store_primal"))
        assert n(fused) < 0.95 * n(strict)
        assert sum(o.op in ("add", "sub") for o in fused.ssa) < 0.6 * sum(o.op in ("add", "sub") for o in strict.ssa)
        a = replay.replay(strict.blob, xs, 5, 10)[0]
        b = replay.replay(fused.blob, xs, 5, 10)[0]
        j64, _, _ = ve_numpy.jacve(w.jaxpr(), w.order(order), w.argnums, [x.astype(np.float64) for x in xs], np.float64)
        t64 = pack_blocks(j64, 2000, w.sizes)
        assert (np.abs(a - b) <= 1.5 * parity_bound(t64)).all()
        mx = np.abs(t64).max(axis=(1, 2), keepdims=True)
        rms = lambda e: float(np.sqrt(np.mean((e / mx) ** 2)))
        assert rms(b - t64) <= 1.02 * rms(a - t64)


def test_launch_shape_fits_shared_memory_and_fast_math_heuristic():
    """Wide tiles get fewer warps per CTA so that the per-warp staging fits in 227 KB, if need be with a single
    input stage; the checked fast reciprocal / sqrt path is
    only generated for plans with enough of them."""
    from graphax_b200.codegen import MAX_DYNAMIC_SMEM, checked_fast_math, launch_shape
    from graphax_b200.examples.deep_learning import Encoder, EncoderDecoder
    M = (4, 4)
    enc = build_plan(make_jaxpr(Encoder, [M, (2, 4)] + [M] * 6 + [M, (2, 4), (4,), (2, 1)] + [()] * 4), "rev", range(2, 16))
    threads, _, stages, spt = launch_shape(enc)
    per_warp = 4 * (2 * enc.n_in_scalars * 32 + (enc.n_rows * enc.n_cols + enc.n_rows) * 32 + 4)
    assert threads < 128 and (threads // 32) * per_warp <= MAX_DYNAMIC_SMEM
    encdec = build_plan(make_jaxpr(EncoderDecoder, [M, (2, 4)] + [M] * 9 + [M, (2, 4), (4,), (2, 1)] + [()] * 6), "rev", range(2, 21))
    from graphax_b200.codegen import input_stages
    t2 = launch_shape(encdec)[0]                               # [8 x 180] tile: one warp, and only one input stage fits
    assert t2 == 32 and input_stages(encdec, t2) == 1 and input_stages(enc, threads) == 2
    os.environ["GX_INPUT_STAGED"] = "0"                        # the measured alternative: inputs prefetched into registers
    try:
        assert input_stages(enc, threads) == 0 and input_stages(enc, threads, 2) == 2 and launch_shape(enc)[0] == 128
    finally:
        del os.environ["GX_INPUT_STAGED"]
    assert checked_fast_math(WORKLOADS["roe3d"].plan("rev")) and checked_fast_math(WORKLOADS["roe1d"].plan("rev"))
    assert not checked_fast_math(WORKLOADS["readme"].plan("readme"))


def test_generated_source_shares_sincos_and_uses_checked_reciprocals():
    from graphax_b200.codegen import emit_source
    src = emit_source(WORKLOADS["robotarm"].plan("rev"), nvrtc=True)
    assert src.count("sincosf(") == 6                       # six joint angles, each needs sin and cos
    src = emit_source(WORKLOADS["roe3d"].plan("mixed", rounding="fma"), nvrtc=True)
    body = src.split("template <bool kPrecise, class Prefetch>\n__device__ __forceinline__ unsigned gx_body")[1].split("return (kPrecise")[0]
    assert body.count("gx_rcp<kPrecise>(") == 9 and body.count("gx_sqrt<kPrecise>(") == 3
    assert "__frcp_rn(v" not in body and "__fdiv_rn(" not in body
    # every operand of a checked operation is folded into the running range exactly once, two per instruction
    # (squares of an already checked denominator need no check of their own)
    assert 8 <= 2 * body.count("gx_chk2(") + body.count("gx_chk1(") <= 12
    # the reference rounding: no FMA outside the division / square-root sequences, three divisions per x/y,
    # one shared reciprocal per denominator
    src = emit_source(WORKLOADS["roe3d"].plan("mixed", rounding="reference"), nvrtc=True)
    body = src.split("template <bool kPrecise, class Prefetch>\n__device__ __forceinline__ unsigned gx_body")[1].split("return (kPrecise")[0]
    assert "__fmaf_rn(" not in body and "__fdiv_rn(" not in body
    assert body.count("gx_div<kPrecise>(") == 39 and body.count("gx_rcp<kPrecise>(") == 16
