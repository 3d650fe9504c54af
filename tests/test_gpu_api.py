"""GPU tests of the reference-facing operator API (jacve / vmap), modelled on tests/examples/example_test.py."""
import numpy as np
import pytest

import graphax_b200 as gx
from graphax_b200.configs import EXTRA_WORKLOADS, WORKLOADS
from graphax_b200.examples import RobotArm_6DOF, RoeFlux_1d, RoeFlux_3d, readme_f
from helpers import GOLDEN, pack_blocks

pytestmark = pytest.mark.gpu


def _golden(name):
    import json
    return json.loads((GOLDEN / "jacobians_fp64.json").read_text())[name]


def test_readme_example(built_library):
    """README.rst:41-42: jacve(f, order=[1,3,2,4], argnums=(0,1,2,3))(1., 1., 2., 7.)"""
    jac = gx.jacve(readme_f, order=[1, 3, 2, 4], argnums=(0, 1, 2, 3))(1., 1., 2., 7.)
    fwd = gx.jacve(readme_f, order="fwd", argnums=(0, 1, 2, 3))(1., 1., 2., 7.)
    rev = gx.jacve(readme_f, order="rev", argnums=(0, 1, 2, 3))(1., 1., 2., 7.)
    assert isinstance(jac, tuple) and len(jac) == 4 and all(j.shape == () for j in jac)
    assert gx.tree_allclose(jac, fwd) and gx.tree_allclose(jac, rev)
    x, y, z, w = 1., 1., 2., 7.
    a = x * y
    expect = [(np.cos(a) * y + z * y) * w, (np.cos(a) * x + z * x) * w, a * w, np.sin(a) + a * z]
    assert gx.tree_allclose(jac, tuple(np.float32(e) for e in expect))


@pytest.mark.parametrize("order", ["fwd", "rev", "alphagrad", "markowitz"])
def test_RobotArm_6DOF(built_library, order):
    w = WORKLOADS["robotarm"]
    g = _golden("robotarm")
    out, aux = gx.jacve(RobotArm_6DOF, order=w.order(order), argnums=range(6), count_ops=True)(*[.02] * 6)
    assert len(out) == 6 and len(out[0]) == 6
    J = np.array([[float(b) for b in row] for row in out])
    ref = np.array(g["jacobian"])
    assert np.allclose(J, ref, atol=1e-5 * np.abs(ref).max(), rtol=1e-4)
    assert aux["num_muls"] == {"fwd": 397, "rev": 301, "alphagrad": 254, "markowitz": 288}[order]


@pytest.mark.parametrize("order", ["fwd", "rev", "alphagrad", "markowitz"])
def test_RoeFlux_1d(built_library, order):
    w = WORKLOADS["roe1d"]
    g = _golden("roe1d")
    primal, jac = gx.jacve(RoeFlux_1d, order=w.order(order), argnums=(0, 1, 2, 3, 4, 5), has_aux=True)(*w.base)
    J = np.array([[float(b) for b in row] for row in jac])
    ref = np.array(g["jacobian"])
    assert np.allclose(J, ref, atol=1e-5 * np.abs(ref).max(), rtol=1e-4)
    assert np.allclose([float(p) for p in primal], g["primal"], rtol=1e-5)


@pytest.mark.parametrize("order", ["fwd", "rev", "mixed", "mixed2"])
def test_RoeFlux_3d(built_library, order):
    import torch
    w = WORKLOADS["roe3d"]
    g = _golden("roe3d")
    args = [torch.tensor(v) for v in ([.1], [.1, .2, .3], [.5], [.2], [.2, .2, .4], [.6])]
    jac = gx.jacve(RoeFlux_3d, order=w.order(order), argnums=list(range(6)))(*args)
    assert [tuple(b.shape) for b in jac[1]] == [(3, 1), (3, 3), (3, 1), (3, 1), (3, 3), (3, 1)]
    J = pack_blocks([[b[None] for b in row] for row in jac], 1, w.sizes)[0]
    ref = np.array(g["jacobian"])
    assert np.allclose(J, ref, atol=1e-5 * np.abs(ref).max(), rtol=1e-4)
    # the reference's own order list, written against jacve(vmap(f)) numbering, is accepted as is
    if order == "mixed":
        from graphax_b200.examples.orders import ROEFLUX_3D_VMAP_MIXED
        jac2 = gx.jacve(RoeFlux_3d, order=ROEFLUX_3D_VMAP_MIXED, argnums=list(range(6)),
                        order_numbering="vmap")(*args)
        assert all(torch.equal(a, b) for ra, rb in zip(jac, jac2) for a, b in zip(ra, rb))


def test_vmap_batch_and_argnums_subset(built_library):
    import torch
    w = WORKLOADS["roe3d"]
    xs = [torch.from_numpy(x).cuda() for x in w.inputs(777)]
    full = gx.vmap(gx.jacve(RoeFlux_3d, order="rev", argnums=list(range(6))))(*xs)
    sub = gx.vmap(gx.jacve(RoeFlux_3d, order="rev", argnums=(4, 1)))(*xs)
    assert tuple(full[1][4].shape) == (777, 3, 3)
    assert torch.equal(sub[1][0], full[1][4]) and torch.equal(sub[2][1], full[2][1])
    single = gx.vmap(gx.jacve(RoeFlux_3d, order="rev", argnums=(1,)))(*xs)
    assert torch.equal(single[0], full[0][1])


def test_bad_orders_raise(built_library):
    with pytest.raises(ValueError):
        gx.jacve(readme_f, order="sideways", argnums=(0,))(1., 1., 2., 7.)
    with pytest.raises(ValueError):
        gx.jacve(readme_f, order=[1, 2, 3], argnums=(0,))(1., 1., 2., 7.)      # vertex 4 missing
    with pytest.raises(ValueError):
        gx.jacve(readme_f, order=[1, 2, 3, 4, 5], argnums=(0,))(1., 1., 2., 7.)  # 5 is the output


@pytest.mark.parametrize("name", list(EXTRA_WORKLOADS))
def test_further_reference_examples_via_runtime_specialisation(built_library, name):
    """Simple / Lighthouse / Hole / Helmholtz (tests/examples/example_test.py:43-75): no ahead-of-time kernel
    exists for them, so the first call compiles the plan with NVRTC; fwd, rev and custom orders must agree
    with each other and with the independent fp64 autodiff."""
    import torch
    w = EXTRA_WORKLOADS[name]
    g = _golden(name)
    args, off = [], 0
    for shape, size in zip(w.arg_shapes, w.sizes):
        args.append(torch.tensor(w.base[off:off + size], dtype=torch.float32).reshape(shape))
        off += size
    ref = np.array(g["jacobian"])
    results = []
    for key in w.order_names():
        jf = gx.jacve(w.fun, order=w.order(key), argnums=w.argnums, count_ops=True)
        out, aux = jf(*args)
        dp = next(iter(jf._device_plans.values()))
        assert dp.kernel == "jit"
        rows = out if len(w.jaxpr().outvars) > 1 or len(w.argnums) == 1 else (out,)
        rows = [r if isinstance(r, tuple) else (r,) for r in rows]
        J = pack_blocks([[b[None] for b in r] for r in rows], 1, w.sizes)[0]
        assert np.allclose(J, ref, atol=1e-5 * np.abs(ref).max(), rtol=1e-4), (name, key)
        assert aux["num_muls"] == w.plan(key).stats["num_muls"]
        results.append(J)
    for J in results[1:]:
        assert np.allclose(J, results[0], atol=1e-5 * np.abs(ref).max(), rtol=1e-4)
    if name == "helmholtz":      # one input, one output: a 1-tuple holding the [4, 4] block (core.py:76-92)
        out, _ = gx.jacve(w.fun, order=[2, 5, 4, 3, 1], argnums=(0,), count_ops=True)(*args)
        assert isinstance(out, tuple) and len(out) == 1 and tuple(out[0].shape) == (4, 4)


@pytest.mark.parametrize("order", ["fwd", "rev"])
def test_function_closing_over_arrays_on_the_gpu(built_library, order):
    """GaussianDataFitting closes over its 65 data points and sample times (reference `examples/minpack.py`; constants
    of the trace, core.py:374): single calls and a vmapped batch on the GPU equal the Jacobians of the reference run
    (tests/golden/make_reference_goldens.py::run_closures) within the parity bound, with the reference's count_ops."""
    import json
    import torch
    from graphax_b200.examples import GaussianDataFitting
    from helpers import GOLDEN, parity_bound
    rec = json.loads((GOLDEN / "reference_run.json").read_text())["closures"]["GaussianDataFitting"]
    arr = np.load(GOLDEN / "reference_run.npz")
    pts, ref = arr["closures/GaussianDataFitting/inputs"], arr["closures/GaussianDataFitting/jac/" + order].astype(np.float64)
    jf = gx.jacve(GaussianDataFitting, order=order, argnums=tuple(range(11)), count_ops=True)
    out, aux = jf(*[float(v) for v in pts[0]])
    assert (aux["num_muls"], aux["num_adds"]) == (rec["orders"][order]["num_muls"], rec["orders"][order]["num_adds"])
    J = np.stack([b.cpu().numpy().reshape(65) for b in out], axis=1)             # one output: a tuple over the arguments
    assert (np.abs(J - ref[0]) <= parity_bound(ref[:1])[0]).all()
    batched, _ = gx.vmap(jf)(*[torch.from_numpy(np.ascontiguousarray(pts[:, k])).cuda() for k in range(11)])
    JB = np.stack([b.cpu().numpy().reshape(len(pts), 65) for b in batched], axis=2)
    assert (np.abs(JB - ref) <= parity_bound(ref)).all()


@pytest.mark.parametrize("order", ["fwd", "rev"])
def test_elementwise_function_of_matrices(built_library, order):
    """Reference test_Simple with 50x50 arguments (example_test.py:49-54): every edge is diagonal, the blocks
    come back dense as out.shape + in.shape = [50,50,50,50] (or as their diagonal values with
    sparse_representation=True)."""
    import torch
    from graphax_b200.examples import Simple
    rng = np.random.default_rng(4)
    x, y = rng.uniform(0.5, 1.5, (50, 50)).astype(np.float32), rng.uniform(0.5, 1.5, (50, 50)).astype(np.float32)
    jac = gx.jacve(Simple, order=order, argnums=(0, 1))(x, y)
    assert len(jac) == 2 and len(jac[0]) == 2 and tuple(jac[0][0].shape) == (50, 50, 50, 50)
    xt, yt = torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
    z = xt * yt
    want = {(0, 0): (torch.cos(z) + 1) * yt, (0, 1): (torch.cos(z) + 1) * xt,
            (1, 0): torch.cos(z) / torch.sin(z) * yt, (1, 1): torch.cos(z) / torch.sin(z) * xt}
    for (o, a), d in want.items():
        blk = jac[o][a].cpu().double().reshape(2500, 2500)
        assert torch.allclose(blk.diagonal(), d.reshape(-1), rtol=1e-5, atol=1e-6)
        assert float((blk - torch.diag(blk.diagonal())).abs().max()) == 0.0
    sp = gx.jacve(Simple, order=order, argnums=(0, 1), sparse_representation=True)(x, y)
    assert sp[1][0].val_shape == [50, 50] and sp[1][0].out_dims[0][0] == "sparse"
    assert torch.allclose(sp[1][0].dense.cpu().double(), want[(1, 0)], rtol=1e-5, atol=1e-6)
    _, aux = gx.jacve(Simple, order=order, argnums=(0, 1), count_ops=True)(x, y)
    _, aux1 = gx.jacve(Simple, order=order, argnums=(0, 1), count_ops=True)(5., 7.)
    assert aux["num_muls"] == 2500 * aux1["num_muls"] and aux["num_adds"] == 2500 * aux1["num_adds"]


@pytest.mark.gpu
def test_reference_timing_script_form(built_library):
    """performance_measurements.py:105-127 times jacve(RoeFlux_1d, order, argnums)(*six (600,) arrays): an
    elementwise function of vectors, whose [600, 600] blocks are diagonal.  The diagonal must equal the batched
    per-sample Jacobians."""
    import torch
    from graphax_b200.examples import RoeFlux_1d
    w = WORKLOADS["roe1d"]
    xs = [torch.from_numpy(x).cuda() for x in w.inputs(600)]
    jac = gx.jacve(RoeFlux_1d, order="rev", argnums=list(range(6)))(*xs)
    ref = gx.vmap(gx.jacve(RoeFlux_1d, order="rev", argnums=list(range(6))))(*xs)
    assert len(jac) == 3 and len(jac[0]) == 6 and tuple(jac[0][0].shape) == (600, 600)
    for o in range(3):
        for a in range(6):
            assert torch.equal(jac[o][a].diagonal(), ref[o][a])
            assert float((jac[o][a] - torch.diag(jac[o][a].diagonal())).abs().max()) == 0.0
