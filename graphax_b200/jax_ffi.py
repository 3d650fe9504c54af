"""`jax.ffi` binding of the jacve launch - the boundary BASELINE.json names.  EXPERIMENTAL: never compiled or run (it
needs JAX, jaxlib's XLA FFI headers and a GPU at the same time; this image has no JAX); nothing here is imported by
the rest of the package, and jit / vmap composition is the design intent, not a tested property.

With JAX installed::

    from graphax_b200 import jax_ffi
    jac = jax.jit(jax_ffi.jacve(f, order="rev", argnums=(0, 1)))(x, y)     # x, y: jax arrays [batch, ...]

`f` is traced with `jax.make_jaxpr` (jax_frontend.py), the elimination plan is built once per shape signature and
uploaded (`gx_plan_create`), and every call is one `jax.ffi.ffi_call` of the handler in `csrc/gx_xla_ffi.cc`, which
forwards XLA's CUDA stream to `gx_jacve_launch`.  `vmap_method="broadcast_all"` folds nested vmaps into the leading
batch axis, so `jax.vmap` of the returned function composes; `jit` composes by construction.  Differentiating THROUGH
the call (second order) would need a `custom_jvp` whose rule is another plan and is not provided.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path
from typing import Callable, Dict, Sequence, Tuple

import numpy as np

_PKG = Path(__file__).resolve().parent
_LIB = _PKG / "libgraphax_b200_xla.so"
_registered = False


def available() -> bool:
    try:
        import jax
        return hasattr(jax, "ffi") and hasattr(jax.ffi, "ffi_call")
    except ImportError:
        return False


def build_handler(force: bool = False) -> Path:
    """Compile csrc/gx_xla_ffi.cc against jaxlib's FFI headers, in-tree."""
    import jax
    if _LIB.exists() and not force:
        return _LIB
    cuda = Path("/usr/local/cuda/include")
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", f"-I{jax.ffi.include_dir()}", f"-I{cuda}",
           f"-I{_PKG.parent / 'include'}", str(_PKG / "csrc" / "gx_xla_ffi.cc"), f"-L{_PKG}", "-lgraphax_b200",
           f"-Wl,-rpath,{_PKG}", "-o", str(_LIB)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(f"building the XLA FFI handler failed:\n{r.stderr}")
    return _LIB


def register() -> None:
    global _registered
    if _registered:
        return
    import jax
    lib = ctypes.CDLL(str(build_handler()))
    jax.ffi.register_ffi_target("gx_jacve", jax.ffi.pycapsule(lib.GxJacve), platform="CUDA")
    _registered = True


def jacve(fun: Callable, order, argnums: Sequence[int] = (0,), device: int = None) -> Callable:
    """Batched Jacobian of `fun` (arguments carry a leading batch axis) as a JAX function built on one
    `jax.ffi.ffi_call`; result pytree as `graphax.jacve` (tuple per output of tuple per argnum), blocks
    `[batch, *out, *in]`.  `device` is the CUDA device the plan is uploaded to (default: JAX's first GPU); it is
    fixed HERE, outside any trace - under `jax.jit` the arguments are tracers and carry no device."""
    from .api import JacveFunction
    host = JacveFunction(fun, order, argnums, False, False, False, "per_sample", None, True, False, "jax")
    plans: Dict[Tuple, object] = {}
    if device is None:
        import jax
        device = jax.devices("gpu")[0].id

    def batched(*xs):
        import jax
        import jax.numpy as jnp
        register()
        shapes = [tuple(x.shape[1:]) for x in xs]
        plan = host.lower(shapes)
        dp = plans.setdefault((plan.key, device), host._device_plan(plan, device))
        batch = xs[0].shape[0]
        flat = [jnp.reshape(x.astype(jnp.float32), (batch, -1)) for x in xs]
        call = jax.ffi.ffi_call("gx_jacve", jax.ShapeDtypeStruct((batch, plan.n_rows, plan.n_cols), jnp.float32),
                                vmap_method="broadcast_all")
        tile = call(*flat, plan=np.int64(dp._h.value))
        return host._package(plan, tile, None, shapes, batch, True)
    return batched
