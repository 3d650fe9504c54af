"""The C-ABI library loads without a GPU, exports what include/graphax_b200.h declares, validates plan
blobs before touching CUDA and fails loudly (never silently computing on the CPU) when no device exists."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from graphax_b200 import runtime
from graphax_b200.configs import WORKLOADS, named_plans

ROOT = Path(__file__).resolve().parent.parent


def test_exports_every_declared_symbol(built_library):
    header = (ROOT / "include" / "graphax_b200.h").read_text()
    declared = set(re.findall(r"\b(gx_[a-z_0-9]+)\s*\(", header))
    assert declared == set(runtime.EXPORTS)
    for name in declared:
        assert hasattr(built_library, name), name
    assert built_library.gx_abi_version() == 2


def test_aot_kernels_cover_every_named_plan(built_library):
    keys = set(runtime.aot_keys())
    for name, plan in named_plans():
        assert int(plan.key, 16) in keys, name
        assert built_library.gx_hash_bytes(plan.blob, len(plan.blob)) == int(plan.key, 16)


def _create(lib, blob: bytes):
    h = ctypes.c_void_p()
    rc = lib.gx_plan_create(blob, len(blob), 0, ctypes.byref(h))
    return rc, lib.gx_last_error().decode(), h


def test_malformed_blobs_are_rejected_before_cuda(built_library):
    good = WORKLOADS["roe1d"].plan("rev").blob
    cases = {
        "bad magic": b"\0" + good[1:],
        "shorter than its header": good[:40],
        "size field": good + b"\0" * 12,
        "unsupported version": good[:4] + (7).to_bytes(4, "little") + good[8:],
    }
    for what, blob in cases.items():
        rc, msg, _ = _create(built_library, blob)
        assert rc == -1 and what in msg, (what, rc, msg)
    # corrupt one op: read of a never-written slot
    ops_off = len(good) - 12 * int.from_bytes(good[32:36], "little")
    bad = bytearray(good)
    bad[ops_off + 4:ops_off + 6] = (0x3fff).to_bytes(2, "little")
    rc, msg, _ = _create(built_library, bytes(bad))
    assert rc == -1 and "slot" in msg
    bad = bytearray(good)
    bad[ops_off:ops_off + 2] = (200).to_bytes(2, "little")
    rc, msg, _ = _create(built_library, bytes(bad))
    assert rc == -1 and "opcode" in msg


def test_no_cpu_fallback(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc, msg, _ = _create(built_library, WORKLOADS["roe1d"].plan("rev").blob)
    assert rc == -3 and "no CPU fallback" in msg
    import graphax_b200 as gx
    from graphax_b200.examples import readme_f
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gx.jacve(readme_f, order="rev", argnums=(0,))(1., 1., 2., 7.)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gx.vmap(gx.jacve(readme_f, order="rev", argnums=(0,)))(*[np.ones(4, np.float32)] * 4)


def test_null_arguments(built_library):
    assert built_library.gx_plan_create(None, 0, 0, None) == -2
    assert built_library.gx_jacve_launch(None, 1, None, None, None, None) == -2
    assert built_library.gx_plan_info(None, None) == -2
    built_library.gx_plan_destroy(None)


def test_lowering_needs_no_gpu():
    """Tracing + plan building is host work: jacve(...).lower(...) must run on a CPU-only box."""
    import graphax_b200 as gx
    from graphax_b200.examples import RoeFlux_3d
    w = WORKLOADS["roe3d"]
    jf = gx.jacve(RoeFlux_3d, order=w.order("mixed"), argnums=w.argnums)
    plan = jf.lower(w.arg_shapes)
    assert plan.key == w.plan("mixed").key and plan.n_rows == 5 and plan.n_cols == 10
    from graphax_b200.examples.orders import ROEFLUX_3D_VMAP_MIXED
    jf2 = gx.jacve(RoeFlux_3d, order=ROEFLUX_3D_VMAP_MIXED, argnums=w.argnums, order_numbering="vmap")
    assert jf2.lower(w.arg_shapes).key == plan.key


def test_api_helpers_without_a_gpu():
    """Host-side logic of the newer API paths: elementwise detection, vmap in_axes validation, nested tracing."""
    import graphax_b200 as gx
    from graphax_b200.examples import Simple, RoeFlux_1d
    from graphax_b200.examples.economics import BlackScholes_Jacobian
    from graphax_b200.tracer import make_jaxpr
    jf = gx.jacve(Simple, order="rev", argnums=(0, 1))
    assert jf._elementwise_scalar_jaxpr([(50, 50), (50, 50)]) is not None        # every equation maps [50,50] -> [50,50]
    assert jf._elementwise_scalar_jaxpr([(50, 50), (50,)]) is None               # broadcasting: not elementwise
    assert gx.jacve(RoeFlux_1d, order="rev", argnums=(0,))._elementwise_scalar_jaxpr([(8,)] * 6) is not None
    with pytest.raises(NotImplementedError):
        gx.vmap(jf, in_axes=(0, 1))
    with pytest.raises(TypeError):
        gx.vmap(lambda x: x)
    assert len(make_jaxpr(BlackScholes_Jacobian, [()] * 5).outvars) == 5          # inner plan replayed on the outer trace


def test_product_never_imports_test_infrastructure():
    """oracle/ and the JAX stand-in under tests/golden/jaxlite are checkers: nothing in the package may import them
    (the guarded `import jax` of jax_frontend.py is the real JAX of a machine that has it)."""
    import re
    from pathlib import Path
    pkg = Path(__file__).resolve().parent.parent / "graphax_b200"
    for f in pkg.rglob("*.py"):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+(oracle|jaxlite)\b", text, re.M), f
        assert "golden/" not in text.replace("tests/golden/jaxlite``", "") or f.name == "jax_frontend.py", f   # no fixture paths in the product
        if re.search(r"^(from|import)\s+jax\b", text, re.M):
            raise AssertionError(f"{f}: JAX must stay an optional, function-local import")


def test_jax_ffi_shim_is_guarded_and_matches_the_header():
    """The `jax.ffi` binding (unverifiable here: no JAX) must import without JAX, say so, and its C++ handler may only
    use entry points the C ABI header declares."""
    import re
    from pathlib import Path
    from graphax_b200 import jax_ffi
    root = Path(__file__).resolve().parent.parent
    if not jax_ffi.available():
        import pytest
        with pytest.raises(ImportError):
            jax_ffi.build_handler()
    header = (root / "include" / "graphax_b200.h").read_text()
    handler = (root / "graphax_b200" / "csrc" / "gx_xla_ffi.cc").read_text()
    for sym in set(re.findall(r"\b(gx_[a-z_]+)\s*\(", handler)):
        assert re.search(rf"\b{sym}\s*\(", header), sym
    assert "XLA_FFI_DEFINE_HANDLER_SYMBOL(GxJacve" in handler and "PlatformStream<cudaStream_t>" in handler
