"""The scalable form the engine implements - `jax.vmap(graphax.jacve(f, order, argnums))` - run with
the reference's OWN sources on the NumPy stand-in for JAX, in a subprocess (the stand-in must not shadow anything in
this process).  Its Jacobians must EQUAL the NumPy oracle's, which the CUDA kernels equal in turn
(tests/test_gpu_full_parity.py).  This is also exactly what `bench.py --impl reference` times."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
JAXLITE = ROOT / "tests" / "golden" / "jaxlite"
SOURCES = [p for p in (Path("/root/reference/src"), ROOT / "baseline" / "_ref") if (p / "graphax" / "core.py").exists()]

SCRIPT = r"""
import sys, json
import numpy as np
sys.path[:0] = [{jaxlite!r}, {src!r}]
sys.path.insert(0, {root!r})
import jax, jax.numpy as jnp, graphax
from graphax import examples as gex
from graphax_b200.configs import WORKLOADS
assert "jaxlite" in jax.__file__ and {src!r} in graphax.__file__
out = {{}}
for name, fun, orders in (("roe3d", gex.RoeFlux_3d, ("rev", "mixed")), ("roe1d", gex.RoeFlux_1d, ("fwd", "markowitz"))):
    w = WORKLOADS[name]
    xs = w.inputs(257)
    for o in orders:
        order = w.order(o)
        order = order if isinstance(order, str) else [int(v) for v in order]
        jac = jax.vmap(graphax.jacve(fun, order=order, argnums=tuple(w.argnums)))(*[jnp.array(x) for x in xs])
        tile = np.concatenate([np.concatenate([np.asarray(b, np.float32).reshape(257, -1, s) for b, s in zip(row, w.sizes)], axis=2)
                               for row in jac], axis=1)
        np.save({tmp!r} + f"/{{name}}_{{o}}.npy", tile)
print("ok")
"""


@pytest.mark.skipif(not SOURCES, reason="neither /root/reference nor baseline/_ref holds the reference's sources")
def test_reference_vmap_form_equals_the_oracle(tmp_path):
    from graphax_b200.configs import WORKLOADS
    from helpers import pack_blocks
    from oracle import ve_numpy
    code = SCRIPT.format(jaxlite=str(JAXLITE), src=str(SOURCES[0]), root=str(ROOT), tmp=str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
    for name, orders in (("roe3d", ("rev", "mixed")), ("roe1d", ("fwd", "markowitz"))):
        w = WORKLOADS[name]
        xs = w.inputs(257)
        for o in orders:
            ref = np.load(tmp_path / f"{name}_{o}.npy")
            j32, _, _ = ve_numpy.jacve(w.jaxpr(), w.order(o), w.argnums, xs, np.float32)
            assert np.array_equal(ref, pack_blocks(j32, 257, w.sizes)), (name, o)
