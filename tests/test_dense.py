"""Tensor-valued path (scope row (f)1): tcgen05 contraction kernel and reverse-order elimination of an MLP."""
import numpy as np
import pytest

from graphax_b200.examples import MLP_loss
from graphax_b200.tracer import make_jaxpr


def _mlp_args(B, I, H, O, seed=0):
    rng = np.random.default_rng(seed)
    return [(rng.standard_normal(s) * sc).astype(np.float32) for s, sc in
            [((B, I), 1.0), ((H, I), I ** -0.5), ((H,), 0.1), ((O, H), H ** -0.5), ((O,), 0.1), ((B, O), 1.0)]]


def test_dense_oracle_matches_independent_autodiff():
    """Pin the tensor-level oracle with torch.autograd in fp64 (the role jax.jacrev plays in the reference's
    tests/examples/example_test.py:168-238)."""
    import torch
    from oracle import dense_numpy as dn
    args = [a.astype(np.float64) for a in _mlp_args(64, 16, 32, 8)]
    jaxpr = make_jaxpr(MLP_loss, [a.shape for a in args])
    grads, loss = dn.jacve_rev(jaxpr, (1, 2, 3, 4), args)
    t = [torch.tensor(a, requires_grad=True) for a in args]
    d = torch.tanh(torch.tanh(t[0] @ t[1].T + t[2]) @ t[3].T + t[4]) - t[5]
    L = .5 * (d ** 2).sum()
    L.backward()
    assert abs(float(loss) - float(L.detach())) < 1e-10
    for g, ti in zip(grads, (t[1], t[2], t[3], t[4])):
        assert np.abs(g - ti.grad.numpy()).max() < 1e-12


def _reference_mlp():
    """Weight Jacobians computed by the reference's own code for the batched MLP of its test_vmap_NeuralNetwork
    (example_test.py:204-215) at non-square widths 6 -> 10 -> 4, batch 12 (tests/golden/make_reference_goldens.py)."""
    import json
    from helpers import GOLDEN
    arr = np.load(GOLDEN / "reference_run.npz")
    rec = json.loads((GOLDEN / "reference_run.json").read_text())["mlp_batched"]
    args = [arr[f"mlp_batched/in{i}"] for i in range(6)]
    ref = {o: ([arr[f"mlp_batched/grad{a}/{o}"] for a in (1, 2, 3, 4)], float(arr[f"mlp_batched/loss/{o}"]))
           for o in ("fwd", "rev")}
    return rec, args, ref


def test_dense_oracle_matches_reference_run():
    """The tensor-level oracle against the reference run: same loss bit for bit, weight Jacobians within float32
    summation-order noise (3e-7 of max|J|) of BOTH the reference's forward and reverse results, which the fp64
    evaluation of the oracle confirms to the same level."""
    from oracle import dense_numpy as dn
    rec, args, ref = _reference_mlp()
    assert rec["primitives"].count("dot_general") == 2 and rec["argnums"] == [1, 2, 3, 4]
    jaxpr = make_jaxpr(MLP_loss, [a.shape for a in args])
    g32, l32 = dn.jacve_rev(jaxpr, (1, 2, 3, 4), args, np.float32)
    g64, _ = dn.jacve_rev(jaxpr, (1, 2, 3, 4), [a.astype(np.float64) for a in args], np.float64)
    for order in ("fwd", "rev"):
        grads, loss = ref[order]
        assert np.float32(l32) == np.float32(loss)
        for r, g, gg in zip(grads, g32, g64):
            scale = np.abs(r).max()
            assert np.abs(r - g.reshape(r.shape)).max() <= 5e-7 * scale
            assert np.abs(r - gg.reshape(r.shape)).max() <= 5e-7 * scale


def test_dense_plan_structure():
    from graphax_b200.dense import build_dense_plan
    jaxpr = make_jaxpr(MLP_loss, [(256, 128), (256, 128), (256,), (128, 256), (128,), (256, 128)])
    assert [e.primitive for e in jaxpr.eqns] == ["dot_general", "broadcast_in_dim", "add", "tanh", "dot_general",
                                                 "broadcast_in_dim", "add", "tanh", "sub", "integer_pow",
                                                 "reduce_sum", "mul"]
    plan = build_dense_plan(jaxpr, "rev", (1, 2, 3, 4))
    kinds = [n.kind for n in plan.dag.nodes]
    assert kinds.count("gemm") == 6 and kinds.count("colsum") == 2 and kinds.count("sumall") == 1
    assert [plan.dag.nodes[g].shape for g in plan.grads] == [(256, 128), (1, 256), (128, 256), (1, 128)]
    with pytest.raises(NotImplementedError):
        build_dense_plan(jaxpr, "fwd", (1,))


def test_launch_groups_of_the_mlp_step(built_library, monkeypatch):
    """Host-side planning of config 5 (no GPU): three contractions with fused epilogues, then the two weight-Jacobian
    contractions - independent plain split-K launches - as one pair; a pair never contains a dependent contraction."""
    from graphax_b200.dense import DenseExecutor, build_dense_plan
    B, I, H, O = 4096, 512, 1024, 512
    jaxpr = make_jaxpr(MLP_loss, [(B, I), (H, I), (H,), (O, H), (O,), (B, O)])
    plan = build_dense_plan(jaxpr, "rev", (1, 2, 3, 4))
    ex = DenseExecutor(plan, 0)
    groups = ex.launch_groups(host_zero=True)
    kinds = [tuple(plan.dag.nodes[n].kind for n in g) for g in groups]
    assert kinds == [("gemm",), ("gemm",), ("gemm",), ("gemm", "gemm")]
    assert all(g[0] in ex.gemm_fused for g in groups[:3])
    a, b = groups[3]
    assert a not in plan.dag.nodes[b].args and {plan.dag.nodes[a].shape, plan.dag.nodes[b].shape} == {(O, H), (H, I)}
    monkeypatch.setenv("GX_DENSE_DUAL", "0")
    assert [len(g) for g in ex.launch_groups(host_zero=True)] == [1] * 5


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,k,splits", [(128, 128, 64, 1), (256, 384, 1024, 1), (512, 256, 2048, 4),
                                          (1024, 512, 4096, 8)])
def test_tcgen05_gemm(built_library, m, n, k, splits):
    """C = A B^T on the tensor cores vs an fp64 product of the same bf16-rounded operands."""
    import torch
    from graphax_b200 import runtime
    torch.manual_seed(1)
    a = (torch.randn(m, k) * 0.5).bfloat16()
    b = (torch.randn(n, k) * 0.5).bfloat16()
    ref = (a.double() @ b.double().T).numpy()
    c = runtime.gemm_bf16_tn(a.cuda(), b.cuda(), k_splits=splits).cpu().numpy()
    assert np.abs(c - ref).max() <= 2e-6 * k * 0.25 + 1e-5          # fp32 accumulation of k products of size ~0.25
    if splits == 1:
        cb = runtime.gemm_bf16_tn(a.cuda(), b.cuda(), out_dtype=torch.bfloat16).float().cpu().numpy()
        assert np.abs(cb - ref).max() <= 2 ** -8 * np.abs(ref).max()
    x = torch.randn(384, 640).bfloat16().cuda()
    assert torch.equal(runtime.transpose_bf16(x), x.T.contiguous())


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["bf16", "f32_as_bf16", "fp32"])
@pytest.mark.parametrize("sizes", [(512, 128, 256, 128), (2048, 256, 1024, 768), (500, 100, 300, 10), (3000, 784, 1000, 10)])
def test_mlp_weight_jacobians(built_library, mode, sizes, monkeypatch):
    """jacve(MLP_loss, 'rev', argnums=(W1, b1, W2, b2)) on the GPU vs the oracle.

    * ``bf16``: bf16 matrix arguments (config 5), ``f32_as_bf16``: float32 arguments with ``precision="bf16"``
      requested explicitly - operands of the contractions are rounded to bf16: tight against the oracle run with
      the same rounding points, 3e-2 of max|J| against exact arithmetic.  The small cases run split-K contractions
      and separate elementwise kernels, the large ones (>= 96 output tiles per contraction) the run-time compiled
      contractions with fused epilogues; both must also agree with the unfused execution.
    * ``fp32``: float32 arguments, default precision: three-term bf16 split contractions; within the parity bound
      |dJ| <= 1e-5 max|J| + 1e-5 |J| of the float64 oracle, i.e. float32 accuracy."""
    import torch
    from graphax_b200.dense import dense_jacve
    from oracle import dense_numpy as dn
    B, I, H, O = sizes
    args = _mlp_args(B, I, H, O)
    tens = [torch.from_numpy(a).cuda() for a in args]
    if mode == "bf16":
        tens = [t.bfloat16() if t.ndim == 2 else t for t in tens]
        args = [t.float().cpu().numpy() for t in tens]              # the oracle sees the same rounded inputs
    kw = {"precision": "bf16"} if mode == "f32_as_bf16" else {}
    jaxpr = make_jaxpr(MLP_loss, [a.shape for a in args])
    loss, grads = dense_jacve(MLP_loss, "rev", (1, 2, 3, 4), *tens, has_aux=True, **kw)
    torch.cuda.synchronize()
    ref64, loss64 = dn.jacve_rev(jaxpr, (1, 2, 3, 4), [a.astype(np.float64) for a in args], np.float64)
    ex = [v for k, v in dense_jacve.__kwdefaults__["_cache"].items() if k[3] == tuple(a.shape for a in args)][-1]
    if mode == "fp32":
        assert ex.precision == "fp32" and not ex.gemm_fused
        assert abs(float(loss) - float(loss64)) <= 2e-5 * abs(float(loss64))
        for g, r in zip(grads, ref64):
            g = g.cpu().numpy()
            assert g.shape == r.shape
            bound = 1e-5 * np.abs(r).max() + 1e-5 * np.abs(r)
            assert (np.abs(g - r) <= bound).all(), float((np.abs(g - r) / bound).max())
        return
    assert ex.precision == "bf16"
    emu, loss_emu = dn.jacve_rev(jaxpr, (1, 2, 3, 4), args, np.float32, bf16_operands=True)
    assert abs(float(loss) - float(loss_emu)) <= 2e-3 * abs(float(loss_emu))
    for g, e, r in zip(grads, emu, ref64):
        g = g.cpu().numpy()
        assert g.shape == r.shape
        scale = np.abs(r).max()
        assert np.abs(g - e).max() <= 2e-3 * scale       # same rounding points as the engine: tight
        assert np.abs(g - r).max() <= 3e-2 * scale       # vs exact arithmetic: bf16 operand rounding
    assert ex.gemm_fused or B < 2048                     # the large cases really took the fused path
    if ex.gemm_fused:
        monkeypatch.setenv("GX_DENSE_FUSE", "0")
        dense_jacve.__kwdefaults__["_cache"].clear()
        _, plain = dense_jacve(MLP_loss, "rev", (1, 2, 3, 4), *tens, has_aux=True, **kw)
        dense_jacve.__kwdefaults__["_cache"].clear()
        for g, p in zip(grads, plain):
            # same rounding points; only the summation order of column sums / split-K partials differs
            assert float((g - p).abs().max()) <= 1e-4 * float(p.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,k,splits,a_mn,b_mn", [(256, 128, 64, 1, False, False), (256, 384, 1024, 1, False, True),
                                                    (512, 256, 2048, 4, True, True), (1024, 512, 4096, 8, True, False),
                                                    (4096, 1024, 512, 1, False, False), (384, 256, 512, 1, False, False)])
def test_tcgen05_gemm_cta_pairs(built_library, m, n, k, splits, a_mn, b_mn):
    """The cta_group::2 variant (clusters of two CTAs on one 256 x 128 tile, each loading half of the B tile; the
    leader issues the MMAs of both, its commits release the stages of both) gives the results of the one-CTA kernel:
    K-major and MN-major operands, split-K, shallow and deep rings; an odd number of tile rows (384) falls back."""
    import torch
    from graphax_b200 import runtime
    torch.manual_seed(2)
    a = (torch.randn(m, k) * 0.5).bfloat16()
    b = (torch.randn(n, k) * 0.5).bfloat16()
    ref = (a.double() @ b.double().T).numpy()
    da = (a.T.contiguous() if a_mn else a).cuda()
    db = (b.T.contiguous() if b_mn else b).cuda()
    lib = runtime.load_library()
    single = runtime.gemm_bf16_tn(da, db, k_splits=splits, a_mn=a_mn, b_mn=b_mn).cpu().numpy()
    before = lib.gx_gemm_set_pair(1)
    try:
        paired = runtime.gemm_bf16_tn(da, db, k_splits=splits, a_mn=a_mn, b_mn=b_mn).cpu().numpy()
        if splits == 1:
            paired16 = runtime.gemm_bf16_tn(da, db, out_dtype=torch.bfloat16, a_mn=a_mn, b_mn=b_mn).float().cpu().numpy()
    finally:
        lib.gx_gemm_set_pair(before)
    assert np.abs(paired - ref).max() <= 2e-6 * k * 0.25 + 1e-5
    if splits == 1:
        assert np.array_equal(paired, single)            # same products, same accumulation order per element
        assert np.abs(paired16 - ref).max() <= 2 ** -8 * np.abs(ref).max()
    else:
        assert np.abs(paired - single).max() <= 1e-5 * np.abs(ref).max()    # the order of the split-K atomics differs


@pytest.mark.gpu
def test_mlp_weight_jacobians_cta_pairs(built_library, monkeypatch):
    """The run-time compiled contractions with fused epilogues in their CTA-pair form (GX_GEMM_PAIR=1)."""
    import torch
    from graphax_b200.dense import dense_jacve
    B, I, H, O = 2048, 256, 1024, 768
    tens = [torch.from_numpy(a).cuda() for a in _mlp_args(B, I, H, O)]
    tens = [t.bfloat16() if t.ndim == 2 else t for t in tens]
    cache = dense_jacve.__kwdefaults__["_cache"]
    cache.clear()
    _, single = dense_jacve(MLP_loss, "rev", (1, 2, 3, 4), *tens, has_aux=True)
    ex = list(cache.values())[-1]
    assert ex.gemm_fused and all(meta[3] in (3, 6) for meta in ex._fused_meta.values())
    cache.clear()
    monkeypatch.setenv("GX_GEMM_PAIR", "1")
    lib = ex.lib
    before = lib.gx_gemm_set_pair(1)
    try:
        _, paired = dense_jacve(MLP_loss, "rev", (1, 2, 3, 4), *tens, has_aux=True)
        torch.cuda.synchronize()
        ex = list(cache.values())[-1]
        assert ex.gemm_fused and all(meta[3] in (4, 8) for meta in ex._fused_meta.values())
    finally:
        lib.gx_gemm_set_pair(before)
        cache.clear()
    for g, p in zip(single, paired):
        assert float((g - p).abs().max()) <= 1e-4 * float(g.abs().max())


@pytest.mark.gpu
def test_arena_cleared_by_the_first_contraction(built_library, monkeypatch):
    """The step's accumulator arena (split-K outputs, bias column sums) is cleared by the epilogue warps of the first
    contraction instead of a memset launch: repeated steps of one executor give the same Jacobians (a missed clear
    would double them), equal to the memset variant up to the order of the atomic additions."""
    import torch
    from graphax_b200.dense import dense_jacve
    B, I, H, O = 2048, 256, 1024, 768
    tens = [torch.from_numpy(a).cuda() for a in _mlp_args(B, I, H, O)]
    tens = [t.bfloat16() if t.ndim == 2 else t for t in tens]
    cache = dense_jacve.__kwdefaults__["_cache"]
    cache.clear()
    runs = []
    for _ in range(3):
        _, grads = dense_jacve(MLP_loss, "rev", (1, 2, 3, 4), *tens, has_aux=True)
        runs.append([g.clone() for g in grads])
    def launches_per_step():
        ex = list(cache.values())[-1]
        l0 = ex.launches
        ex.run(tens)
        return ex.launches - l0
    fused_clear = launches_per_step()
    cache.clear()
    monkeypatch.setenv("GX_DENSE_MEMSET", "1")
    _, with_memset = dense_jacve(MLP_loss, "rev", (1, 2, 3, 4), *tens, has_aux=True)
    assert launches_per_step() == fused_clear + 1      # the memset is the only launch that disappears
    cache.clear()
    for a, b, c, m in zip(*runs, with_memset):
        scale = float(m.abs().max())
        assert float((a - b).abs().max()) <= 1e-5 * scale and float((a - c).abs().max()) <= 1e-5 * scale
        assert float((a - m).abs().max()) <= 1e-5 * scale


@pytest.mark.gpu
def test_dual_contraction_launch(built_library, monkeypatch):
    """Two independent contractions in one launch (gx_gemm_bf16_dual_ld) give the results of two launches: different
    shapes, operand layouts and split counts side by side; the MLP step uses it for its two weight Jacobians."""
    import ctypes
    import torch
    from graphax_b200 import runtime
    from graphax_b200.dense import dense_jacve
    lib = runtime.load_library()
    torch.manual_seed(3)
    probs = [(512, 1024, 4096, 4, True, True), (1024, 512, 2048, 2, False, True)]
    ops, refs, outs = [], [], []
    for (m, n, k, splits, a_mn, b_mn) in probs:
        a = (torch.randn(m, k) * 0.5).bfloat16()
        b = (torch.randn(n, k) * 0.5).bfloat16()
        refs.append((a.double() @ b.double().T).numpy())
        ops.append(((a.T.contiguous() if a_mn else a).cuda(), (b.T.contiguous() if b_mn else b).cuda()))
        outs.append(torch.zeros((m, n), dtype=torch.float32, device="cuda"))
    I2, P2 = ctypes.c_int * 2, ctypes.c_void_p * 2
    col = lambda i: I2(*[p[i] for p in probs])
    rc = lib.gx_gemm_bf16_dual_ld(col(0), col(1), col(2), P2(*[o[0].data_ptr() for o in ops]),
                                  I2(*[o[0].shape[1] for o in ops]), I2(*[int(p[4]) for p in probs]),
                                  P2(*[o[1].data_ptr() for o in ops]), I2(*[o[1].shape[1] for o in ops]),
                                  I2(*[int(p[5]) for p in probs]), P2(*[o.data_ptr() for o in outs]), I2(0, 0), I2(0, 0),
                                  col(3), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.gx_gemm_last_error().decode()
    torch.cuda.synchronize()
    for (m, n, k, *_), out, ref in zip(probs, outs, refs):
        assert np.abs(out.cpu().numpy() - ref).max() <= 2e-6 * k * 0.25 + 1e-5
    # and inside the MLP step: with and without the pairing of the two weight-Jacobian contractions
    B, I, H, O = 4096, 512, 1024, 512
    tens = [torch.from_numpy(a).cuda() for a in _mlp_args(B, I, H, O)]
    tens = [t.bfloat16() if t.ndim == 2 else t for t in tens]
    cache = dense_jacve.__kwdefaults__["_cache"]
    cache.clear()
    _, paired = dense_jacve(MLP_loss, "rev", (1, 2, 3, 4), *tens, has_aux=True)
    ex = list(cache.values())[-1]
    l0 = ex.launches
    ex.run(tens)
    n_paired = ex.launches - l0
    cache.clear()
    monkeypatch.setenv("GX_DENSE_DUAL", "0")
    _, single = dense_jacve(MLP_loss, "rev", (1, 2, 3, 4), *tens, has_aux=True)
    ex = list(cache.values())[-1]
    l0 = ex.launches
    ex.run(tens)
    assert ex.launches - l0 == n_paired + 1
    cache.clear()
    for g, p in zip(single, paired):
        assert float((g - p).abs().max()) <= 1e-5 * float(g.abs().max())


@pytest.mark.gpu
def test_tensor_path_matches_the_reference_run(built_library):
    """The tcgen05 path against weight Jacobians computed by the reference's own code (batched MLP of the
    reference's test_vmap_NeuralNetwork at widths 6 -> 10 -> 4, batch 12): float32 arguments take the three-term
    split contractions and must sit within |dJ| <= 1e-5 max|J| + 1e-5 |J| of the reference's forward AND reverse
    results; with ``precision="bf16"`` the same call is only good to bf16 operand rounding."""
    import torch
    from graphax_b200.dense import dense_jacve
    rec, args, ref = _reference_mlp()
    tens = [torch.from_numpy(a).cuda() for a in args]
    loss, grads = dense_jacve(MLP_loss, "rev", (1, 2, 3, 4), *tens, has_aux=True)
    torch.cuda.synchronize()
    for order in ("fwd", "rev"):
        r_grads, r_loss = ref[order]
        assert abs(float(loss) - r_loss) <= 1e-5 * abs(r_loss)
        for g, r in zip(grads, r_grads):
            g, r = g.cpu().numpy().astype(np.float64), r.astype(np.float64)
            assert g.shape == r.shape
            bound = 1e-5 * np.abs(r).max() + 1e-5 * np.abs(r)
            assert (np.abs(g - r) <= bound).all(), float((np.abs(g - r) / bound).max())
    _, coarse = dense_jacve(MLP_loss, "rev", (1, 2, 3, 4), *tens, has_aux=True, precision="bf16")
    worst = max(float(np.abs(g.cpu().numpy() - r).max() / np.abs(r).max()) for g, r in zip(coarse, ref["rev"][0]))
    assert 1e-5 < worst < 3e-2


@pytest.mark.gpu
def test_split3_contraction(built_library):
    """gx_split3_bf16 + gx_gemm_bf16x3_ld: C = A B^T of float32 matrices to float32 accuracy on the bf16 tensor cores."""
    import ctypes
    import torch
    from graphax_b200 import runtime
    lib = runtime.load_library()
    torch.manual_seed(5)
    m, n, k = 256, 384, 1000
    a = torch.randn(m, k, device="cuda")
    b = torch.randn(n, k, device="cuda") * 3.0
    stream = torch.cuda.current_stream().cuda_stream
    parts = []
    for t in (a, b):
        p3 = [torch.empty_like(t, dtype=torch.bfloat16) for _ in range(3)]
        assert lib.gx_split3_bf16(t.data_ptr(), p3[0].data_ptr(), p3[1].data_ptr(), p3[2].data_ptr(), t.numel(), stream) == 0
        torch.cuda.synchronize()
        resid = (t.double() - sum(p.double() for p in p3)).abs().max()
        assert float(resid) <= 2.0 ** -23 * float(t.abs().max())
        parts.append(p3)
    c = torch.zeros(m, n, device="cuda")
    vp = ctypes.c_void_p
    rc = lib.gx_gemm_bf16x3_ld(m, n, k, (vp * 3)(*[p.data_ptr() for p in parts[0]]), k, 0,
                               (vp * 3)(*[p.data_ptr() for p in parts[1]]), k, 0, c.data_ptr(), n, 1, stream)
    assert rc == 0, lib.gx_gemm_last_error()
    torch.cuda.synchronize()
    ref = a.double() @ b.double().T
    err = float((c.double() - ref).abs().max())
    assert err <= 2e-6 * float(ref.abs().max()), err          # float32 accumulation of 1000 products
    one = torch.zeros(m, n, device="cuda")
    assert lib.gx_gemm_bf16_ld(m, n, k, parts[0][0].data_ptr(), k, 0, parts[1][0].data_ptr(), k, 0, one.data_ptr(), n, 0, 1, stream) == 0
    torch.cuda.synchronize()
    assert float((one.double() - ref).abs().max()) > 100 * err  # a single bf16 product is two orders worse


@pytest.mark.gpu
def test_jacve_api_dispatches_matrix_arguments(built_library):
    """graphax_b200.jacve(f, 'rev', argnums)(x, W1, b1, W2, b2, y): matrix arguments take the tensor-valued path
    and come back like jax.jacrev would return them - one array per argnum, shaped like the argument."""
    import torch
    import graphax_b200 as gx
    args = _mlp_args(256, 128, 128, 128, seed=3)
    tens = [torch.from_numpy(a).cuda() for a in args]
    loss, grads = gx.jacve(MLP_loss, order="rev", argnums=(1, 2, 3, 4), has_aux=True, precision="bf16")(*tens)
    t = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in args]
    L = .5 * ((torch.tanh(torch.tanh(t[0] @ t[1].T + t[2]) @ t[3].T + t[4]) - t[5]) ** 2).sum()
    L.backward()
    assert [tuple(g.shape) for g in grads] == [(128, 128), (128,), (128, 128), (128,)]
    assert abs(float(loss) - float(L.detach())) <= 5e-3 * abs(float(L.detach()))
    for g, ti in zip(grads, (t[1], t[2], t[3], t[4])):
        assert float((g.double().cpu() - ti.grad).abs().max()) <= 3e-2 * float(ti.grad.abs().max())
    with pytest.raises(NotImplementedError):
        gx.jacve(MLP_loss, order="fwd", argnums=(1,))(*tens)          # 65k input scalars: refused, not attempted


@pytest.mark.gpu
def test_dense_jacve_other_orders_on_small_problems(built_library):
    """`dense_jacve(order="fwd")`: tensor-level elimination is reverse-only, but small problems (the reference's own
    test sizes) run on the per-sample engine in any order - here against the reference run of the batched MLP,
    whose forward and reverse results both exist."""
    import torch
    from graphax_b200.dense import dense_jacve
    rec, args, ref = _reference_mlp()
    tens = [torch.from_numpy(a).cuda() for a in args]
    loss, grads = dense_jacve(MLP_loss, "fwd", (1, 2, 3, 4), *tens, has_aux=True)
    r_grads, r_loss = ref["fwd"]
    assert abs(float(loss) - r_loss) <= 1e-5 * abs(r_loss)
    for g, r in zip(grads, r_grads):
        g = g.cpu().numpy().astype(np.float64).reshape(r.shape)
        bound = 1e-5 * np.abs(r).max() + 1e-5 * np.abs(r)
        assert (np.abs(g - r) <= bound).all()


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,k", [(4096, 4096, 512), (8192, 2048, 576), (4096, 3968, 512), (2560, 3840, 1024)])
def test_persistent_gemm_variants(built_library, m, n, k):
    """Grids of several waves run on the persistent kernels: 128 x 256 tiles with a double-buffered TMEM
    accumulator (4096 x 4096: 512 tiles = 3 rounds of 148 CTAs + 68 tiles cut into 136 half-width items),
    128 x 128 tiles when N is not a multiple of 256; every operand layout, fp32 and bf16 results."""
    import torch
    from graphax_b200 import runtime
    torch.manual_seed(1)
    A = (torch.randn(m, k, device="cuda") * 0.1).bfloat16()
    B = (torch.randn(n, k, device="cuda") * 0.1).bfloat16()
    ref = A.float() @ B.float().T
    for a_mn in (False, True):
        for b_mn in (False, True):
            a = A.T.contiguous() if a_mn else A
            b = B.T.contiguous() if b_mn else B
            c32 = runtime.gemm_bf16_tn(a, b, out_dtype=torch.float32, a_mn=a_mn, b_mn=b_mn)
            c16 = runtime.gemm_bf16_tn(a, b, out_dtype=torch.bfloat16, a_mn=a_mn, b_mn=b_mn)
            assert float((c32 - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
            assert float((c16.float() - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max())
