"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded inputs."""
import numpy as np
import pytest

from graphax_b200.configs import WORKLOADS
from helpers import bits, pack_blocks, parity_bound, run_engine

pytestmark = pytest.mark.gpu

EXACT = {"roe1d", "roe3d"}           # plans made only of correctly rounded ops: bit-exact vs the replay
CASES = [(w.name, k) for w in WORKLOADS.values() for k in w.order_names()]
# plan variants: the reference's own rounding (default), and the contracted ("fma") rounding with x/y lowered through a
# reciprocal or through the reference's three IEEE divisions (plan.build_plan, lowering.lower_eqn)
DIVISIONS = ["reference", "reciprocal", "ieee"]


def _plan(w, order, mode):
    return w.plan(order, rounding="reference") if mode == "reference" else w.plan(order, division=mode)


def _tile64(w, order, xs):
    from oracle import ve_numpy
    jac, primal, _ = ve_numpy.jacve(w.jaxpr(), w.order(order), w.argnums, [x.astype(np.float64) for x in xs],
                                    np.float64)
    return pack_blocks(jac, len(xs[0]), w.sizes), np.concatenate([p.reshape(len(xs[0]), -1) for p in primal], 1)


@pytest.mark.parametrize("name,order", CASES)
@pytest.mark.parametrize("division", DIVISIONS)
@pytest.mark.parametrize("kernel", ["aot", "interpreter"])
def test_kernel_matches_plan_replay(built_library, name, order, kernel, division):
    from oracle import replay
    w = WORKLOADS[name]
    plan = _plan(w, order, division)
    batch = 4096 + 37                                  # full tiles + ragged tail
    xs = w.inputs(batch)
    jac, primal, used = run_engine(plan, xs, kernel)
    assert used == kernel
    ref_j, ref_p = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    assert not np.isnan(jac).any() and not np.isnan(primal).any()
    if name in EXACT:
        assert np.array_equal(bits(jac), bits(ref_j))
        assert np.array_equal(bits(primal), bits(ref_p))
    else:   # sin / cos / atan2 differ between CUDA libm and glibc by a few ulp
        t64, _ = _tile64(w, order, xs)
        assert (np.abs(jac - ref_j) <= parity_bound(t64)).all()
        np.testing.assert_allclose(primal, ref_p, rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("name,order", [("roe3d", "rev"), ("roe3d", "mixed"), ("roe1d", "rev"), ("robotarm", "markowitz")])
@pytest.mark.parametrize("spt", [1, 2])
def test_jit_kernel_matches_plan_replay(built_library, name, order, spt):
    """NVRTC-specialised kernels, scalar and packed (two samples per thread through FFMA2/FMUL2/FADD2)."""
    from oracle import replay
    w = WORKLOADS[name]
    plan = w.plan(order)
    xs = w.inputs(1000)
    jac, primal, used = run_engine(plan, xs, "jit", samples_per_thread=spt)
    assert used == "jit"
    ref_j, ref_p = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    if name in EXACT:
        assert np.array_equal(bits(jac), bits(ref_j)) and np.array_equal(bits(primal), bits(ref_p))
    else:
        t64, _ = _tile64(w, order, xs)
        assert (np.abs(jac - ref_j) <= parity_bound(t64)).all()


@pytest.mark.parametrize("name,order", CASES)
@pytest.mark.parametrize("division", DIVISIONS)
def test_engine_matches_reference_order_oracle(built_library, name, order, division):
    """Engine vs the NumPy restatement of the reference algorithm (fp32, unfused, tensor-level order,
    the reference's division formulas) within |dJ| <= 1e-5 max|J| + 1e-5 |J|; both also against fp64."""
    from oracle import ve_numpy
    w = WORKLOADS[name]
    plan = _plan(w, order, division)
    batch = 2048
    xs = w.inputs(batch)
    jac, primal, _ = run_engine(plan, xs, "aot")
    j32, p32, aux = ve_numpy.jacve(w.jaxpr(), w.order(order), w.argnums, xs, np.float32)
    t32 = pack_blocks(j32, batch, w.sizes)
    t64, p64 = _tile64(w, order, xs)
    bound = parity_bound(t64)
    assert (np.abs(jac - t32) <= bound).all(), float((np.abs(jac - t32) / bound).max())
    if division == "reference" and name in EXACT:
        assert np.array_equal(jac, t32) and np.array_equal(primal, np.concatenate([p.reshape(batch, -1) for p in p32], 1))
    mx = np.abs(t64).max(axis=(1, 2), keepdims=True)
    err_engine, err_oracle = (np.abs(jac - t64) / mx).max(), (np.abs(t32 - t64) / mx).max()
    assert err_engine < 3e-5 and err_oracle < 3e-5, (err_engine, err_oracle)
    np.testing.assert_allclose(primal, p64, rtol=1e-5, atol=1e-6)
    assert aux["num_muls"] == plan.stats["num_muls"] and aux["num_adds"] == plan.stats["num_adds"]


def _reference_run():
    import json
    from helpers import GOLDEN
    run = json.loads((GOLDEN / "reference_run.json").read_text())
    return run, np.load(GOLDEN / "reference_run.npz")


REF_CASES = [(n, o) for n in ("roe3d", "roe1d", "robotarm") for o in WORKLOADS[n].order_names()]


@pytest.mark.parametrize("name,order", REF_CASES)
@pytest.mark.parametrize("division", DIVISIONS)
@pytest.mark.parametrize("kernel", ["aot", "interpreter"])
def test_kernels_match_reference_run(built_library, name, order, division, kernel):
    """CUDA kernels against float32 Jacobians computed by the reference's own code (unmodified graphax sources on
    the NumPy stand-in for JAX, tests/golden/make_reference_goldens.py) at the committed points: north-star
    tolerance rtol 1e-5 in the form |dJ| <= 1e-5 max|J_sample| + 1e-5 |J| (the kernels contract mul+add into FMA
    and, in `reciprocal` mode, replace x/y by x*(1/y) with one correction step)."""
    run, arr = _reference_run()
    w = WORKLOADS[name]
    flat = arr[f"{name}/inputs"]
    xs, off = [], 0
    for shape, size in zip(w.arg_shapes, w.sizes):
        xs.append(np.ascontiguousarray(flat[:, off:off + size].reshape((len(flat),) + tuple(shape))))
        off += size
    plan = _plan(w, order, division)
    assert plan.structure["order"] == run["workloads"][name]["orders"][order]["eliminated"]
    jac, primal, used = run_engine(plan, xs, kernel)
    assert used == kernel
    ref = arr[f"{name}/jac/{order}"].astype(np.float64)
    bound = parity_bound(ref)
    assert np.isfinite(jac).all()
    assert (np.abs(jac - ref) <= bound).all(), float((np.abs(jac - ref) / bound).max())
    np.testing.assert_allclose(primal, arr[f"{name}/primal"], rtol=2e-6, atol=1e-6)


def _example_names():
    import json
    from helpers import GOLDEN
    return list(json.loads((GOLDEN / "reference_run.json").read_text())["examples"])


@pytest.mark.parametrize("name", _example_names())
@pytest.mark.parametrize("order", ["fwd", "rev"])
def test_further_reference_examples_on_the_gpu(built_library, name, order):
    """KerrSenn_metric, FreeEnergy, CloudSchemes_step, PropaneCombustion, HumanHeartDipole, SteadyStateCombustion and
    BlackScholes, each rebuilt from the jaxpr the reference traced, through run-time specialised kernels against
    the Jacobians the reference computed (tests/golden/make_reference_goldens.py)."""
    from graphax_b200.plan import build_plan
    from helpers import jaxpr_from_record
    run, arr = _reference_run()
    rec = run["examples"][name]
    jaxpr = jaxpr_from_record(rec)
    flat = arr[f"examples/{name}/inputs"]
    xs, off = [], 0
    for shape in rec["in_shapes"]:
        size = int(np.prod(shape, dtype=np.int64)) if shape else 1
        xs.append(np.ascontiguousarray(flat[:, off:off + size].reshape((len(flat),) + tuple(shape))))
        off += size
    plan = build_plan(jaxpr, order, tuple(range(len(xs))))
    jac, _, used = run_engine(plan, xs, "jit")
    assert used == "jit"
    ref = arr[f"examples/{name}/jac/{order}"].astype(np.float64)
    bound = parity_bound(ref)
    assert np.isfinite(jac).all()
    assert (np.abs(jac - ref) <= bound).all(), float((np.abs(jac - ref) / bound).max())


def _random_cases():
    import json
    from helpers import GOLDEN
    run = json.loads((GOLDEN / "reference_run.json").read_text())
    return [(n, k) for n, rec in run["random_orders"].items() for k in rec]


@pytest.mark.parametrize("name,key", _random_cases())
def test_random_orders_on_the_gpu(built_library, name, key):
    """Arbitrary caller-given orders: seeded random permutations, run-time specialised kernels against the Jacobians
    the reference computed for the same order (bound widened by the order's own fp32 noise, see the CPU test)."""
    from graphax_b200.plan import build_plan
    from oracle import ve_numpy
    run, arr = _reference_run()
    w, ro = WORKLOADS[name], run["random_orders"][name][key]
    flat = arr[f"random_orders/{name}/inputs"]
    xs, off = [], 0
    for shape, size in zip(w.arg_shapes, w.sizes):
        xs.append(np.ascontiguousarray(flat[:, off:off + size].reshape((len(flat),) + tuple(shape))))
        off += size
    plan = build_plan(w.jaxpr(), ro["order"], w.argnums)
    jac, _, used = run_engine(plan, xs, "jit")
    assert used == "jit"
    ref = arr[f"random_orders/{name}/jac/{key}"].astype(np.float64)
    j64, _, _ = ve_numpy.jacve(w.jaxpr(), ro["order"], w.argnums, [x.astype(np.float64) for x in xs], np.float64)
    noise = np.abs(ref - pack_blocks(j64, len(flat), w.sizes)).max(axis=(1, 2), keepdims=True)
    assert np.isfinite(jac).all()
    assert (np.abs(jac - ref) <= parity_bound(ref) + 2 * noise).all()


@pytest.mark.parametrize("batch", [0, 1, 127, 128, 129, 300])
@pytest.mark.parametrize("kernel", ["aot", "interpreter"])
def test_ragged_batches(built_library, batch, kernel):
    from oracle import replay
    w = WORKLOADS["roe3d"]
    plan = w.plan("rev")
    xs = w.inputs(max(batch, 1))
    xs = [x[:batch] for x in xs]
    if batch == 0:
        import torch
        from graphax_b200 import runtime
        dp = runtime.DevicePlan(plan, 0, jit=False)
        dp.launch(0, [0] * 6, 0, None, torch.cuda.current_stream().cuda_stream)   # no-op, must not fault
        dp.close()
        return
    jac, primal, _ = run_engine(plan, xs, kernel)
    ref_j, ref_p = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    assert np.array_equal(bits(jac), bits(ref_j)) and np.array_equal(bits(primal), bits(ref_p))


def test_misaligned_buffers_take_the_guarded_path(built_library):
    """Caller buffers that are not 16-byte aligned cannot use bulk copies; results must not change."""
    import torch
    from graphax_b200 import runtime
    from oracle import replay
    w = WORKLOADS["roe3d"]
    plan = w.plan("rev")
    batch = 515
    xs = w.inputs(batch)
    dp = runtime.DevicePlan(plan, 0, jit=False)
    dev = []
    for x in xs:
        flat = torch.zeros(x.size + 1, device="cuda")
        flat[1:] = torch.from_numpy(x.reshape(-1)).cuda()
        dev.append(flat[1:])                                   # data_ptr() is 4 bytes off alignment
    jac_flat = torch.zeros(batch * 50 + 1, device="cuda")
    dp.launch(batch, [t.data_ptr() for t in dev], jac_flat[1:].data_ptr(), None,
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref_j, _ = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    assert np.array_equal(bits(jac_flat[1:].cpu().numpy().reshape(batch, 5, 10)), bits(ref_j))
    dp.close()


def test_full_size_properties(built_library):
    """Headline size (2^20): batch independence, determinism, primal consistency, NaN-freeness."""
    import torch
    from graphax_b200 import runtime
    from oracle import replay
    w = WORKLOADS["roe3d"]
    batch = 1 << 20
    xs = w.inputs(batch)
    dev = [torch.from_numpy(x).cuda() for x in xs]
    stream = torch.cuda.current_stream().cuda_stream
    for order in ("rev", "fwd", "mixed"):
        plan = w.plan(order)
        dp = runtime.DevicePlan(plan, 0, jit=False)
        assert dp.kernel == "aot"
        jac = torch.empty((batch, 5, 10), device="cuda")
        jac2 = torch.empty_like(jac)
        primal = torch.empty((batch, 5), device="cuda")
        dp.launch(batch, [t.data_ptr() for t in dev], jac.data_ptr(), primal.data_ptr(), stream)
        dp.launch(batch, [t.data_ptr() for t in dev], jac2.data_ptr(), None, stream)
        torch.cuda.synchronize()
        assert torch.equal(jac, jac2)                              # deterministic
        assert bool(torch.isfinite(jac).all())
        # any window of the big batch equals a small run on that window (samples are independent)
        for lo in (0, 123_457, batch - 4099):
            sl = slice(lo, lo + 4099)
            ref_j, ref_p = replay.replay(plan.blob, [x[sl] for x in xs], plan.n_rows, plan.n_cols)
            assert np.array_equal(bits(jac[sl].cpu().numpy()), bits(ref_j))
            assert np.array_equal(bits(primal[sl].cpu().numpy()), bits(ref_p))
        dp.close()
    # different elimination orders accumulate the same Jacobian (within the fp32 order spread)
    j_rev = run_engine(w.plan("rev"), [x[:8192] for x in xs], "aot")[0]
    j_mix = run_engine(w.plan("mixed"), [x[:8192] for x in xs], "aot")[0]
    mx = np.abs(j_rev).max(axis=(1, 2), keepdims=True)
    assert (np.abs(j_rev - j_mix) / mx).max() < 3e-5


def test_batch_with_more_than_2_31_jacobian_entries(built_library):
    """Maximum sizes: 43 M samples = 2.15e9 Jacobian entries (8.6 GB), past the range of 32-bit element offsets; ragged.
    The batch is a 2^20-sample block repeated, so every repetition must equal the first (all of the 8.6 GB checked
    on the device) and windows of the first, a middle and the last repetition equal the plan replay."""
    import torch
    from graphax_b200 import runtime
    from oracle import replay
    w = WORKLOADS["roe3d"]
    block, reps, tail = 1 << 20, 41, 19
    batch = block * reps + tail
    assert batch * 50 > 2 ** 31
    if torch.cuda.mem_get_info()[0] < 20 * 2 ** 30:
        pytest.skip("needs 20 GB of free device memory")
    xs = w.inputs(block)
    dev = [torch.cat([torch.from_numpy(x).cuda().repeat((reps,) + (1,) * (x.ndim - 1)), torch.from_numpy(x[:tail]).cuda()])
           for x in xs]
    plan = w.plan("rev", rounding="reference")
    dp = runtime.DevicePlan(plan, 0, jit=False)
    assert dp.kernel == "aot"
    jac = torch.empty((batch, 5, 10), device="cuda")
    primal = torch.empty((batch, 5), device="cuda")
    dp.launch(batch, [t.data_ptr() for t in dev], jac.data_ptr(), primal.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    first_j, first_p = jac[:block], primal[:block]
    for r in range(1, reps):
        assert torch.equal(jac[r * block:(r + 1) * block], first_j), f"repetition {r} differs"
        assert torch.equal(primal[r * block:(r + 1) * block], first_p)
    assert torch.equal(jac[reps * block:], first_j[:tail]) and torch.equal(primal[reps * block:], first_p[:tail])
    for lo in (0, 20 * block + 54_321, (reps - 1) * block + block - 4099):
        sl = slice(lo % block, lo % block + 4099)
        ref_j, ref_p = replay.replay(plan.blob, [x[sl] for x in xs], plan.n_rows, plan.n_cols)
        assert np.array_equal(bits(jac[lo:lo + 4099].cpu().numpy()), bits(ref_j))
        assert np.array_equal(bits(primal[lo:lo + 4099].cpu().numpy()), bits(ref_p))
    dp.close()


def test_host_pipeline(built_library):
    """gx_jacve_host (H2D -> kernel -> D2H, chunked) equals the device-resident launch."""
    import torch
    from graphax_b200 import runtime
    from oracle import replay
    w = WORKLOADS["roe3d"]
    plan = w.plan("rev")
    batch = (1 << 17) + 333                                     # > 2 chunks, ragged
    xs = w.inputs(batch)
    ref_j, ref_p = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    dp = runtime.DevicePlan(plan, 0, jit=False)
    for pinned in (False, True):
        if pinned:
            hin = [torch.from_numpy(x).pin_memory() for x in xs]
            hj = torch.empty((batch, 5, 10)).pin_memory()
            hp = torch.empty((batch, 5)).pin_memory()
        else:
            hin = [torch.from_numpy(x.copy()) for x in xs]
            hj, hp = torch.empty((batch, 5, 10)), torch.empty((batch, 5))
        dp.run_host(batch, [t.data_ptr() for t in hin], hj.data_ptr(), hp.data_ptr(), pinned)
        assert np.array_equal(bits(hj.numpy()), bits(ref_j)) and np.array_equal(bits(hp.numpy()), bits(ref_p))
    dp.close()


@pytest.mark.parametrize("kernel", ["aot", "jit", "interpreter"])
@pytest.mark.parametrize("batch", [1, 95, 4096 + 33])
def test_no_out_of_bounds_writes(built_library, kernel, batch):
    """compute-sanitizer is closed on this pool, so bounds are checked the manual way: canary regions on
    both sides of every output buffer must survive launches that end in a ragged tile."""
    import torch
    from graphax_b200 import runtime
    from oracle import replay
    w = WORKLOADS["roe3d"]
    plan = w.plan("mixed")
    xs = w.inputs(batch)
    dp = runtime.DevicePlan(plan, 0, jit=False)
    if kernel == "jit":
        dp.specialize()
    dp.select_kernel({"interpreter": 0, "aot": 1, "jit": 2}[kernel])
    pad = 4096                                             # floats of canary on each side (16 KiB, keeps alignment)
    dev = [torch.from_numpy(x).cuda() for x in xs]
    jbuf = torch.full((2 * pad + batch * 50,), 12345.0, device="cuda")
    pbuf = torch.full((2 * pad + batch * 5,), 12345.0, device="cuda")
    dp.launch(batch, [t.data_ptr() for t in dev], jbuf[pad:].data_ptr(), pbuf[pad:].data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for buf, n in ((jbuf, batch * 50), (pbuf, batch * 5)):
        assert bool((buf[:pad] == 12345.0).all()) and bool((buf[pad + n:] == 12345.0).all())
    ref_j, ref_p = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    assert np.array_equal(bits(jbuf[pad:pad + batch * 50].cpu().numpy().reshape(batch, 5, 10)), bits(ref_j))
    assert np.array_equal(bits(pbuf[pad:pad + batch * 5].cpu().numpy().reshape(batch, 5)), bits(ref_p))
    dp.close()


@pytest.mark.parametrize("order", ["fwd", "rev", "mixed", "mixed2"])
def test_contracted_plans_run_bit_exact(built_library, order):
    """Opt-in sum contraction (plan._contract_sums): the kernel compiled at run time must reproduce the C
    replay of the contracted plan bit for bit, like every other Roe plan."""
    from graphax_b200.plan import build_plan
    from oracle import replay
    w = WORKLOADS["roe3d"]
    plan = build_plan(w.jaxpr(), w.order(order), w.argnums, contract=True)
    xs = w.inputs(1000)
    jac, primal, used = run_engine(plan, xs, "jit")
    ref_j, ref_p = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    assert used == "jit"
    assert np.array_equal(bits(jac), bits(ref_j)) and np.array_equal(bits(primal), bits(ref_p))


@pytest.mark.gpu
def test_checked_fast_math_falls_back_exactly(built_library):
    """The kernels evaluate 1/y and sqrt(x) through the branch-free fast paths of the IEEE intrinsics and
    re-evaluate a warp tile with the intrinsics themselves when an operand leaves the range where the two agree
    (tools/rcp_sqrt_exhaustive.py: all 2^32 operands).  Zeros, denormals, huge values, infinities, NaNs and
    negative radicands must therefore still reproduce the C replay (IEEE division / sqrtf) bit for bit."""
    from graphax_b200.plan import build_plan
    from graphax_b200.tracer import jnp, make_jaxpr
    from oracle import replay

    def f(x, y):
        return jnp.sqrt(x) / y, x / jnp.sqrt(y)
    plan = build_plan(make_jaxpr(f, [(), ()]), "rev", (0, 1))
    specials = np.array([0.0, -0.0, 1e-45, 1e-40, 1.1754944e-38, 3e-38, 1e-30, 0.5, 1.0, 4.0, 1e30, 3e38,
                         np.inf, -np.inf, np.nan, -1.0, -1e-40], dtype=np.float32)
    xs, ys = np.meshgrid(specials, specials, indexing="ij")
    rng = np.random.default_rng(0)
    # several warp tiles: some all regular (fast path only), some with special lanes mixed in (fallback)
    x = np.concatenate([rng.uniform(0.5, 2.0, 128).astype(np.float32), xs.reshape(-1), rng.uniform(0.5, 2.0, 77).astype(np.float32)])
    y = np.concatenate([rng.uniform(0.5, 2.0, 128).astype(np.float32), ys.reshape(-1), rng.uniform(0.5, 2.0, 77).astype(np.float32)])
    with np.errstate(all="ignore"):
        ref_j, ref_p = replay.replay(plan.blob, [x, y], plan.n_rows, plan.n_cols)
    for kernel in ("jit", "interpreter"):
        jac, primal, used = run_engine(plan, [x, y], kernel)
        assert used == kernel
        for got, ref in ((jac, ref_j), (primal, ref_p)):
            same = (bits(got) == bits(ref)) | (np.isnan(got) & np.isnan(ref))
            assert same.all(), (kernel, np.argwhere(~same)[:5])


@pytest.mark.gpu
def test_wide_input_single_stage_kernel(built_library):
    """400 inputs and a [3 x 400] tile: two input stages per warp no longer fit beside the output tile, so the
    kernel runs with one (gx_skeleton.cuh GX_IN_STAGES) and requests the next tile as soon as the current one is
    in registers.  Many tiles per warp, ragged tail."""
    from graphax_b200.codegen import input_stages, launch_shape
    from graphax_b200.plan import build_plan
    from graphax_b200.tracer import jnp, make_jaxpr
    from oracle import replay

    def f(x):
        return jnp.sum(x * x), jnp.sum(x), x[0] * x[399] + x[7]
    plan = build_plan(make_jaxpr(f, [(400,)]), "rev", (0,))
    assert input_stages(plan, launch_shape(plan)[0]) == 1
    x = np.random.default_rng(3).uniform(-1, 1, (148 * 32 * 3 + 77, 400)).astype(np.float32)
    jac, primal, used = run_engine(plan, [x], "jit")
    assert used == "jit"
    ref_j, ref_p = replay.replay(plan.blob, [x], plan.n_rows, plan.n_cols)
    assert np.array_equal(bits(jac), bits(ref_j)) and np.array_equal(bits(primal), bits(ref_p))
