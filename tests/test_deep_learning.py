"""Perceptron, Encoder, EncoderDecoder (reference tests/examples/example_test.py:240-349): attention blocks,
layer norm, SiLU and softmax cross-entropy on matrix-valued vertices, forward and reverse orders.  The custom
93-vertex order of test_Encoder is written against the jaxpr of the reference's JAX version (which wraps
jax.nn.softmax differently) and cannot be mapped without JAX, so it is not restated here."""
from functools import lru_cache

import numpy as np
import pytest

import graphax_b200 as gx
from graphax_b200.examples.deep_learning import Encoder, EncoderDecoder, Perceptron
from graphax_b200.plan import build_plan
from graphax_b200.tracer import make_jaxpr
from helpers import torch_eval_jaxpr

M = (4, 4)
CASES = {
    # name: (function, per-sample shapes, argnums) as in the reference's tests
    "Perceptron": (Perceptron, [(4,), (4,), (8, 4), (8,), (4, 8), (4,), (), ()], tuple(range(8))),
    "Encoder": (Encoder, [M, (2, 4)] + [M] * 6 + [M, (2, 4), (4,), (2, 1)] + [()] * 4, tuple(range(2, 16))),
    "EncoderDecoder": (EncoderDecoder, [M, (2, 4)] + [M] * 9 + [M, (2, 4), (4,), (2, 1)] + [()] * 6, tuple(range(2, 21))),
}


def _inputs(name, batch, seed=11):
    """Reference test points: ones for x, normal weights, (gamma, beta) = (0, 1) pairs - jittered so that
    every sample differs; gamma = 0 would zero most of the Jacobian, so the scalars are drawn too."""
    _, shapes, _ = CASES[name]
    rng = np.random.default_rng(seed)
    return [rng.normal(0.0, 0.6, (batch,) + s) if s else rng.uniform(0.5, 1.5, (batch,)) for s in shapes]


@lru_cache(maxsize=None)
def _plan(name, order):
    fn, shapes, argnums = CASES[name]
    return build_plan(make_jaxpr(fn, shapes), order, argnums)


def _tile(jaxpr, jac, batch):
    return np.concatenate([np.concatenate([b.reshape(batch, max(1, o.size), -1) for b in row], axis=2)
                           for o, row in zip(jaxpr.outvars, jac)], axis=1)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_torch_autograd(name):
    import torch
    from oracle import dense_ve
    fn, shapes, argnums = CASES[name]
    jaxpr = make_jaxpr(fn, shapes)
    xs = _inputs(name, 2)
    for order in ("fwd", "rev"):
        jac, _ = dense_ve.jacve(jaxpr, order, argnums, xs, np.float64)
        tile = _tile(jaxpr, jac, 2)
        for b in range(2):
            args = [torch.tensor(x[b], dtype=torch.float64, requires_grad=(i in argnums)) for i, x in enumerate(xs)]
            out = torch.cat([o.reshape(-1) for o in torch_eval_jaxpr(jaxpr, args)])
            rows = []
            for r in range(out.numel()):
                g = torch.autograd.grad(out[r], [args[i] for i in argnums], retain_graph=True, allow_unused=True)
                rows.append(torch.cat([(torch.zeros_like(args[i]) if gi is None else gi).reshape(-1)
                                       for i, gi in zip(argnums, g)]))
            ref = torch.stack(rows).numpy()
            assert np.abs(tile[b] - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), (name, order)


@pytest.mark.parametrize("name,order", [("Perceptron", "fwd"), ("Perceptron", "rev"), ("Encoder", "rev"),
                                        ("EncoderDecoder", "rev")])
def test_plan_replay_matches_oracle(name, order):
    from oracle import dense_ve, replay
    fn, shapes, argnums = CASES[name]
    plan = _plan(name, order)
    xs = [x.astype(np.float32) for x in _inputs(name, 48)]
    jac, primal = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    j64, p64 = dense_ve.jacve(make_jaxpr(fn, shapes), order, argnums, [x.astype(np.float64) for x in xs], np.float64)
    t64 = _tile(make_jaxpr(fn, shapes), j64, 48)
    scale = np.abs(t64).max(axis=(1, 2), keepdims=True) + 1e-30
    assert (np.abs(jac - t64) / scale).max() < 1e-4          # fp32 through exp/log/softmax/layer-norm chains
    np.testing.assert_allclose(primal, np.concatenate([p.reshape(48, -1) for p in p64], axis=1), rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name,order", [("Perceptron", "fwd"), ("Perceptron", "rev"), ("Encoder", "rev"),
                                        ("EncoderDecoder", "rev")])
def test_gpu_matches_plan_replay(built_library, name, order, monkeypatch):
    """Through the public API.  The 23k-operation EncoderDecoder plan would take ~50 s to compile at run time
    (tools/encdec_bench.py does that); here the threshold is lowered so that it runs on the interpreter."""
    import torch
    monkeypatch.setenv("GX_JIT_MAX_OPS", "12000")
    from oracle import replay
    fn, shapes, argnums = CASES[name]
    xs = [x.astype(np.float32) for x in _inputs(name, 200)]
    jf = gx.jacve(fn, order=order, argnums=argnums, has_aux=True)
    primal, jac = gx.vmap(jf)(*[torch.from_numpy(x).cuda() for x in xs])
    plan = jf.lower(shapes)
    ref_j, ref_p = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    outs = jac if len(plan.jaxpr.outvars) > 1 else (jac,)
    got = np.concatenate([np.concatenate([b.reshape(200, max(1, o.size), -1).cpu().numpy() for b in row], axis=2)
                          for o, row in zip(plan.jaxpr.outvars, outs)], axis=1)
    scale = np.abs(ref_j).max(axis=(1, 2), keepdims=True) + 1e-30
    assert (np.abs(got - ref_j) / scale).max() <= 2e-5        # libm (exp/log/tanh) differs by a few ulp host vs device
    for a, blk in zip(argnums, outs[0]):
        assert tuple(blk.shape) == (200,) + tuple(plan.jaxpr.outvars[0].shape) + tuple(shapes[a])


@pytest.mark.gpu
def test_large_plan_interpreter(built_library):
    """Forward mode over the Encoder: 89k operations and 4.9k live values do not fit in shared memory, so the
    plan runs on the global-memory interpreter; it must agree with the replay and with the reverse order."""
    import torch
    from oracle import replay
    from graphax_b200 import runtime
    fn, shapes, argnums = CASES["Encoder"]
    plan = _plan("Encoder", "fwd")
    assert plan.stats["n_ops"] > 50000
    xs = [x.astype(np.float32) for x in _inputs("Encoder", 100)]
    dp = runtime.DevicePlan(plan, 0)
    assert dp.info().kernel_kind == runtime.KERNEL_INTERPRETER and dp.info().smem_bytes == 0
    ins = [torch.from_numpy(x).cuda() for x in xs]
    out = torch.full((100, plan.n_rows, plan.n_cols), float("nan"), device="cuda")
    prim = torch.empty((100, plan.n_rows), device="cuda")
    dp.launch(100, [t.data_ptr() for t in ins], out.data_ptr(), prim.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref_j, ref_p = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    scale = np.abs(ref_j).max(axis=(1, 2), keepdims=True) + 1e-30
    assert (np.abs(out.cpu().numpy() - ref_j) / scale).max() <= 2e-5      # libm (exp / log / tanh): host vs device
    rev = _plan("Encoder", "rev")
    rev_j, _ = replay.replay(rev.blob, xs, rev.n_rows, rev.n_cols)
    assert (np.abs(ref_j - rev_j) / scale).max() <= 1e-4
    dp.close()


@pytest.mark.gpu
def test_large_plan_interpreter_through_the_chunked_host_path(built_library):
    """The global-memory interpreter keeps the live values of CTA b in one per-plan scratch area.  The host pipeline
    (`gx_jacve_host`) launches consecutive chunks on two streams, and callers may launch one plan from several streams:
    launches of such a plan are chained through an event (gx_abi.cu), so more than two chunks - and two user streams -
    must still reproduce the single-launch result bit for bit."""
    import ctypes
    import torch
    from graphax_b200 import runtime
    plan = _plan("Encoder", "fwd")
    batch = 3 * 1024 + 77                                           # > 2 chunks of the (small) host chunk below
    xs = [np.ascontiguousarray(x.astype(np.float32)) for x in _inputs("Encoder", batch)]
    dp = runtime.DevicePlan(plan, 0)
    assert dp.info().kernel_kind == runtime.KERNEL_INTERPRETER and dp.info().smem_bytes == 0
    ins = [torch.from_numpy(x).cuda() for x in xs]
    want = torch.empty((batch, plan.n_rows, plan.n_cols), device="cuda")
    dp.launch(batch, [t.data_ptr() for t in ins], want.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    # (1) two user streams, launched back to back on halves of the batch
    half = batch // 2
    got = torch.full_like(want, float("nan"))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for rep in range(3):
        dp.launch(half, [t[:half].contiguous().data_ptr() for t in ins], got[:half].data_ptr(), None, s1.cuda_stream)
        dp.launch(batch - half, [t[half:].contiguous().data_ptr() for t in ins], got[half:].data_ptr(), None, s2.cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))
    # (2) the chunked host pipeline (pinned buffers: chunks go straight from / to the caller's memory)
    import os
    os.environ["GX_HOST_CHUNK"] = "1024"
    try:
        dp2 = runtime.DevicePlan(plan, 0)
        hin = [torch.from_numpy(x).pin_memory() for x in xs]
        hj = torch.empty((batch, plan.n_rows, plan.n_cols)).pin_memory()
        dp2.run_host(batch, [t.data_ptr() for t in hin], hj.data_ptr(), None, True)
        assert np.array_equal(hj.numpy().view(np.uint32), want.cpu().numpy().view(np.uint32))
        dp2.close()
    finally:
        del os.environ["GX_HOST_CHUNK"]
    dp.close()


@pytest.mark.gpu
def test_count_ops_and_sparse_representation_above_rank_one_through_the_api(built_library):
    """`jacve(..., count_ops=True)` and `sparse_representation=True` on matrix-valued vertices: the reference's own
    numbers (Perceptron: 7762 / 1384 forward, 312 / 21 reverse, tests/golden/reference_trace_rank_n.json.gz) and the
    SparseTensor structure of every Jacobian block (core.py:555-581), with values equal to the dense result."""
    import gzip
    import json
    import torch
    from helpers import GOLDEN
    trace = json.loads(gzip.open(GOLDEN / "reference_trace_rank_n.json.gz").read())["cases"]["Perceptron"]
    fn, shapes, argnums = CASES["Perceptron"]
    xs = [torch.as_tensor(np.asarray(x[0], dtype=np.float32)).cuda() for x in _inputs("Perceptron", 1)]
    for order in ("fwd", "rev"):
        jac, aux = gx.jacve(fn, order=order, argnums=argnums, count_ops=True)(*xs)
        ref = trace["orders"][order]
        assert (aux["num_muls"], aux["num_adds"]) == (ref["num_muls"], ref["num_adds"]) and "model" not in aux
        assert [list(c) for c in aux["order_counts"]] == ref["order_counts"]
        sparse = gx.jacve(fn, order=order, argnums=argnums, sparse_representation=True)(*xs)
        blocks = sparse if isinstance(sparse[0], gx.SparseJacobian) or sparse[0] is None else sparse[0]
        dense = jac if not isinstance(jac[0], tuple) else jac[0]
        for i, (b, d) in enumerate(zip(blocks, ref["blocks"])):
            if d is None:
                assert b is None
                continue
            assert (b.out_dims, b.primal_dims, b.val_shape) == (d["out"], d["primal"], d["val"]), (order, i)
            assert torch.equal(b.dense, dense[i])
