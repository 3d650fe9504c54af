"""Full-size parity of the headline config: every RoeFlux_3d elimination order over the whole 2^20 batch on the GPU
against the float32 NumPy restatement of the reference algorithm (oracle/ve_numpy.py, itself bit-exact with the run
of the reference's own sources, tests/test_reference_goldens.py).

* ``rounding="reference"`` plans must reproduce the oracle exactly: 0 of 52 428 800 Jacobian entries and 0 of
  5 242 880 primal outputs may differ in value (exact zeros may differ in sign: the plan does not add structural
  zeros, ``x + 0.0`` turns -0 into +0 in the reference).
* ``rounding="fma"`` plans are measured against the stated tolerance ``|dJ| <= 1e-5 max|J_sample| + 1e-5 |J|``;
  the counts are recorded (profiles/parity_full_r02.json) and bounded: the ill-conditioned ``mixed`` order exceeds
  the bound on a handful of samples out of 2^20 (two fp32 roundings of the same computation differ that much
  there), which is why the reference rounding exists.
"""
import json
import os
from pathlib import Path

import numpy as np
import pytest

from graphax_b200.configs import WORKLOADS
from helpers import bits, pack_blocks

pytestmark = pytest.mark.gpu

BATCH = 1 << 20
CHUNK = 1 << 16
ORDERS = ["rev", "fwd", "mixed", "mixed2"]


def _gpu_jacobians(plan, xs):
    import torch
    from graphax_b200 import runtime
    dp = runtime.DevicePlan(plan, 0, jit=False)
    assert dp.kernel == "aot", dp.kernel
    dev = [torch.from_numpy(x).cuda() for x in xs]
    jac = torch.empty((BATCH, plan.n_rows, plan.n_cols), device="cuda")
    primal = torch.empty((BATCH, plan.n_rows), device="cuda")
    dp.launch(BATCH, [t.data_ptr() for t in dev], jac.data_ptr(), primal.data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = jac.cpu().numpy(), primal.cpu().numpy()
    dp.close()
    return out


def _oracle_chunks(w, order, xs):
    from oracle import ve_numpy
    for lo in range(0, BATCH, CHUNK):
        part = [x[lo:lo + CHUNK] for x in xs]
        j32, p32, _ = ve_numpy.jacve(w.jaxpr(), w.order(order), w.argnums, part, np.float32)
        yield lo, pack_blocks(j32, CHUNK, w.sizes), np.concatenate([p.reshape(CHUNK, -1) for p in p32], axis=1)


def _record(key, value):
    """Merge one result into gpurun_out/parity_full.json (copied to profiles/ by the builder)."""
    path = Path(os.environ.get("GRAFT_REPO_ROOT", Path(__file__).resolve().parent.parent)) / "gpurun_out" / "parity_full.json"
    try:
        path.parent.mkdir(parents=True, exist_ok=True)
        data = json.loads(path.read_text()) if path.exists() else {}
        data[key] = value
        path.write_text(json.dumps(data, indent=1, sort_keys=True))
    except OSError:
        pass


@pytest.mark.parametrize("order", ORDERS)
def test_reference_rounding_is_exact_over_the_full_batch(built_library, order):
    w = WORKLOADS["roe3d"]
    xs = w.inputs(BATCH)
    plan = w.plan(order, rounding="reference")
    jac, primal = _gpu_jacobians(plan, xs)
    assert np.isfinite(jac).all() and np.isfinite(primal).all()
    value_mismatch = bit_mismatch = primal_mismatch = 0
    for lo, t32, p32 in _oracle_chunks(w, order, xs):
        got = jac[lo:lo + CHUNK]
        value_mismatch += int((got != t32).sum())
        bit_mismatch += int((bits(got) != bits(t32)).sum())
        primal_mismatch += int((primal[lo:lo + CHUNK] != p32).sum())
    _record(f"roe3d:{order}:reference", {"samples": BATCH, "entries": int(jac.size), "value_mismatches": value_mismatch,
                                         "zero_sign_differences": bit_mismatch - value_mismatch,
                                         "primal_mismatches": primal_mismatch})
    assert value_mismatch == 0 and primal_mismatch == 0, (value_mismatch, primal_mismatch)
    assert bit_mismatch <= 1000          # only exact zeros of opposite sign


@pytest.mark.parametrize("order", ORDERS)
@pytest.mark.parametrize("division", ["reciprocal", "ieee"])
def test_fma_rounding_against_the_stated_tolerance_over_the_full_batch(built_library, order, division):
    w = WORKLOADS["roe3d"]
    xs = w.inputs(BATCH)
    plan = w.plan(order, division=division, rounding="fma")
    jac, primal = _gpu_jacobians(plan, xs)
    assert np.isfinite(jac).all()
    worst, over, near = 0.0, 0, 0
    for lo, t32, p32 in _oracle_chunks(w, order, xs):
        ref = t32.astype(np.float64)
        bound = 1e-5 * np.abs(ref).max(axis=(1, 2), keepdims=True) + 1e-5 * np.abs(ref)
        ratio = (np.abs(jac[lo:lo + CHUNK] - ref) / bound).max(axis=(1, 2))
        worst = max(worst, float(ratio.max()))
        over += int((ratio > 1.0).sum())
        near += int((ratio > 0.8).sum())
        np.testing.assert_allclose(primal[lo:lo + CHUNK], p32, rtol=2e-6, atol=1e-7)
    _record(f"roe3d:{order}:fma:{division}", {"samples": BATCH, "worst_ratio_to_bound": worst,
                                              "samples_over_bound": over, "samples_over_0.8_bound": near})
    if order == "mixed":
        # not a pass in the sense of the north-star tolerance: documented, bounded, and the reason the
        # reference rounding is the default
        assert worst < 2.0 and over <= 16, (worst, over)
    else:
        assert over == 0 and worst < 0.5, (worst, over)


@pytest.mark.parametrize("order", ["rev", "fwd", "alphagrad", "markowitz"])
def test_roeflux_1d_reference_rounding_is_exact_over_its_full_batch(built_library, order):
    """Config 2 (RoeFlux_1d, batch 2^16): every order of the reference's tests, the whole batch, equal values."""
    import torch
    from graphax_b200 import runtime
    from oracle import ve_numpy
    w = WORKLOADS["roe1d"]
    batch = w.batch
    xs = w.inputs(batch)
    plan = w.plan(order, rounding="reference")
    dp = runtime.DevicePlan(plan, 0, jit=False)
    assert dp.kernel == "aot"
    dev = [torch.from_numpy(x).cuda() for x in xs]
    jac = torch.empty((batch, plan.n_rows, plan.n_cols), device="cuda")
    primal = torch.empty((batch, plan.n_rows), device="cuda")
    dp.launch(batch, [t.data_ptr() for t in dev], jac.data_ptr(), primal.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    j32, p32, _ = ve_numpy.jacve(w.jaxpr(), w.order(order), w.argnums, xs, np.float32)
    t32 = pack_blocks(j32, batch, w.sizes)
    got = jac.cpu().numpy()
    assert int((got != t32).sum()) == 0
    assert int((primal.cpu().numpy() != np.concatenate([p.reshape(batch, -1) for p in p32], axis=1)).sum()) == 0
    _record(f"roe1d:{order}:reference", {"samples": batch, "entries": int(got.size), "value_mismatches": 0,
                                         "zero_sign_differences": int((bits(got) != bits(t32)).sum())})
    dp.close()
