"""`frontend="jax"`: functions written against `jax.numpy` are traced by `jax.make_jaxpr` and converted
(graphax_b200/jax_frontend.py).  JAX is not installed in this image, so the conversion runs against
tests/golden/jaxlite - the NumPy stand-in for the JAX API that also carries the reference run - in a child process
(so that the stand-in's `jax` never shadows anything in the test session)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

CHILD = r'''
import json, sys
from pathlib import Path
root = Path(sys.argv[1])
sys.path[:0] = [str(root), str(root / "tests" / "golden" / "jaxlite")]
import jax, jax.numpy as jnp            # jaxlite
import graphax_b200 as gx
from graphax_b200 import jax_frontend
from graphax_b200.configs import WORKLOADS, EXTRA_WORKLOADS
from graphax_b200.plan import build_plan
from graphax_b200.tracer import Literal, make_jaxpr

assert jax.__version__.endswith("jaxlite") and jax_frontend.available()
sources = {"roe3d": ("roe_flux.py", "RoeFlux_3d"), "roe1d": ("roe_flux.py", "RoeFlux_1d"),
           "robotarm": ("robot_arm.py", "RobotArm_6DOF"), "simple": ("simple.py", "Simple"),
           "lighthouse": ("simple.py", "Lighthouse"), "helmholtz": ("simple.py", "Helmholtz")}
out = {}
for name, (file, fn) in sources.items():
    w = WORKLOADS.get(name) or EXTRA_WORKLOADS[name]
    text = (root / "graphax_b200" / "examples" / file).read_text().replace("from ..tracer import jnp", "import jax.numpy as jnp")
    scope = {}
    exec(text, scope)
    fun = scope[fn]                       # the same source, now a function of jax.numpy
    via_jax = make_jaxpr(fun, w.arg_shapes, "jax")
    auto = make_jaxpr(fun, w.arg_shapes, "auto")      # the native tracers are rejected -> falls back to JAX
    native = w.jaxpr()
    key = lambda v: ("lit", v.val) if isinstance(v, Literal) else ("in", v.input_index) if v.input_index is not None else ("v", v.eqn_index)
    same = len(via_jax.eqns) == len(native.eqns) == len(auto.eqns)
    for a, b in zip(via_jax.eqns, native.eqns):
        same &= a.primitive == b.primitive and [key(x) for x in a.invars] == [key(x) for x in b.invars] \
            and a.outvar.shape == b.outvar.shape
        same &= all(tuple(a.params[k]) == tuple(b.params[k]) if isinstance(b.params[k], (tuple, list)) else a.params[k] == b.params[k]
                    for k in a.params if k in b.params and b.params[k] is not None)
    same &= via_jax.vmap_pads == native.vmap_pads and [key(v) for v in via_jax.outvars] == [key(v) for v in native.outvars]
    blobs = all(build_plan(via_jax, w.order(o), w.argnums).blob == w.plan(o, with_primal=True).blob for o in w.order_names())
    out[name] = {"same_jaxpr": bool(same), "same_plans": bool(blobs), "n_eqns": len(via_jax.eqns)}

# the stand-in's vmap reproduces jax.vmap's identity pads for RoeFlux_3d (SURVEY.md A.1 rule 5)
text = (root / "graphax_b200" / "examples" / "roe_flux.py").read_text().replace("from ..tracer import jnp", "import jax.numpy as jnp")
scope = {}
exec(text, scope)
w = WORKLOADS["roe3d"]
batched = [jnp.array(x) for x in w.inputs(3)]
vj = jax.make_jaxpr(jax.vmap(scope["RoeFlux_3d"]))(*batched).jaxpr
out["vmap_eqns"] = len(vj.eqns)
out["vmap_pads"] = [n for n, e in enumerate(vj.eqns, 1)
                    if e.primitive.name == "broadcast_in_dim" and not isinstance(e.invars[0], jax.core.Literal)]

# unknown primitives are refused like in the reference (core.py:396-398)
import jax.lax as lax
dn = lax.ScatterDimensionNumbers((0,), (), (0,))
try:
    build_plan(make_jaxpr(lambda x, i: lax.scatter(x, i, x * 2., dn), [(3,), (1,)], "jax"), "rev", (0,))
    out["refuses_untraceable"] = False
except NotImplementedError:
    out["refuses_untraceable"] = True
# the public API takes the jax.numpy function unchanged (lowering only: no GPU in this process)
jf = gx.jacve(scope_fun := (lambda x, y: jnp.sin(x) * y), order="rev", argnums=(0, 1))
plan = jf.lower([(3,), (3,)])
out["api_lowers"] = [plan.n_rows, plan.n_cols, len(plan.jaxpr.eqns)]
# arrays the function closes over arrive as jaxpr.constvars (one per object): same plan as the native tracer
import numpy as np
Xn = (np.arange(15, dtype=np.float32).reshape(5, 3) % 7 - 3) / 4
yn = np.array([0.5, -1.0, 0.25, 2.0, -0.75], np.float32)
Xj, yj = jnp.array(Xn), jnp.array(yn)
def closing_jax(w):
    r = Xj @ w - yj
    return jnp.sum(r * (yj - r)), r * yj
def closing_native(w):
    r = Xn @ w - yn
    return gx.jnp.sum(r * (yn - r)), r * yn
cj, cn = make_jaxpr(closing_jax, [(3,)], "jax"), make_jaxpr(closing_native, [(3,)], "native")
out["closure"] = {"constvars": [list(v.shape) for v in cj.constvars], "input_index": [v.input_index for v in cj.constvars],
                  "same_consts": all(np.array_equal(a, b) for a, b in zip(cj.consts, cn.consts)),
                  "same_plans": all(build_plan(cj, o, (0,)).blob == build_plan(cn, o, (0,)).blob for o in ("fwd", "rev"))}
print(json.dumps(out))
'''


def test_jax_frontend_reproduces_the_native_tracer_on_a_jax_numpy_function():
    r = subprocess.run([sys.executable, "-c", CHILD, str(ROOT)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out.pop("refuses_untraceable") is True
    assert out.pop("api_lowers") == [3, 6, 2]
    assert out.pop("closure") == {"constvars": [[5, 3], [5]], "input_index": [1, 2], "same_consts": True, "same_plans": True}
    assert out.pop("vmap_eqns") == 146 and out.pop("vmap_pads") == [24, 33, 48, 55, 59, 82, 91, 133]
    assert out["roe3d"]["n_eqns"] == 138 and out["roe1d"]["n_eqns"] == 101 and out["robotarm"]["n_eqns"] == 114
    for name, rec in out.items():
        assert rec["same_jaxpr"] and rec["same_plans"], (name, rec)


def test_jax_frontend_needs_jax():
    """Without JAX the explicit frontend fails loudly, and functions of the native jnp keep working."""
    import pytest
    from graphax_b200 import jax_frontend
    from graphax_b200.tracer import jnp, make_jaxpr
    if jax_frontend.available():
        pytest.skip("JAX (or a stand-in) is importable in this session")
    with pytest.raises(ImportError):
        make_jaxpr(lambda x: x * x, [()], "jax")
    assert len(make_jaxpr(lambda x: jnp.sin(x) * x, [()], "auto").eqns) == 2
    with pytest.raises(ValueError):
        make_jaxpr(lambda x: x, [()], "xla")


GPU_CHILD = r'''
import json, sys
from pathlib import Path
root = Path(sys.argv[1])
sys.path[:0] = [str(root), str(root / "tests"), str(root / "tests" / "golden" / "jaxlite")]
import numpy as np, torch
import jax, jax.numpy as jnp            # jaxlite
import graphax_b200 as gx
from graphax_b200.configs import WORKLOADS

text = (root / "graphax_b200" / "examples" / "roe_flux.py").read_text().replace("from ..tracer import jnp", "import jax.numpy as jnp")
scope = {}
exec(text, scope)
w = WORKLOADS["roe3d"]
xs = w.inputs(4096)
jf = gx.jacve(scope["RoeFlux_3d"], order=list(w.vmap_orders["mixed"]), argnums=tuple(range(6)), order_numbering="vmap")
jac = gx.vmap(jf)(*[jnp.array(x) for x in xs])          # stand-in arrays in, torch CUDA tensors out
got = np.concatenate([np.concatenate([b.reshape(4096, b.shape[1], -1).cpu().numpy() for b in row], axis=2) for row in jac], axis=1)
native = gx.jacve(w.fun, order=list(w.vmap_orders["mixed"]), argnums=tuple(range(6)), order_numbering="vmap")
ref = gx.vmap(native)(*[torch.from_numpy(x).cuda() for x in xs])
ref = np.concatenate([np.concatenate([b.reshape(4096, b.shape[1], -1).cpu().numpy() for b in row], axis=2) for row in ref], axis=1)
print(json.dumps({"bit_equal": bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32))), "shape": list(got.shape),
                  "kernel": next(iter(jf._device_plans.values())).kernel}))
'''


def test_gpu_call_with_a_jax_numpy_function():
    """End to end on the GPU: RoeFlux_3d written against `jax.numpy` (stand-in), the reference's mixed order in its
    own 146-numbering, arrays of the other library as inputs - same plan, same ahead-of-time kernel, same bits as the
    native path."""
    import pytest
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    r = subprocess.run([sys.executable, "-c", GPU_CHILD, str(ROOT)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out == {"bit_equal": True, "shape": [4096, 5, 10], "kernel": "aot"}


test_gpu_call_with_a_jax_numpy_function = __import__("pytest").mark.gpu(test_gpu_call_with_a_jax_numpy_function)


def test_nested_call_is_one_vertex_without_in_edges_like_the_reference():
    """`pjit` (what jnp.where / jnp.clip / jax.nn.* become under real JAX): the reference's rule evaluates the call and
    returns no partials (primitives.py:576-602), so the call is one vertex whose derivative is dropped.  Same here,
    with a warning; `iota` and `device_put` are value-only as well (primitives.py:444-455)."""
    import warnings
    import numpy as np
    from graphax_b200.plan import build_plan
    from graphax_b200.tracer import Eqn, Jaxpr, Var, jnp, make_jaxpr
    from oracle import replay
    inner = make_jaxpr(lambda a: jnp.sin(a) * 2.0, [()])
    x, y = Var((), input_index=0), Var((), input_index=1)
    v = [Var((), eqn_index=i) for i in range(4)]
    eqns = [Eqn("pjit", (x,), v[0], {"jaxpr": inner, "name": "g"}), Eqn("device_put", (y,), v[1], {}),
            Eqn("mul", (v[0], y), v[2], {}), Eqn("add", (v[2], x), v[3], {})]
    jaxpr = Jaxpr([x, y], eqns, [v[3]], [0] * 4, False)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        plan = build_plan(jaxpr, "rev", (0, 1))
    assert any("dropped" in str(c.message) for c in caught)
    jac, primal = replay.replay(plan.blob, [np.array([0.5], np.float32), np.array([3.0], np.float32)], 1, 2)
    # f = g(x) * y + x with g opaque: df/dx = 1 (not 1 + 2 cos(x) y), df/dy = g(x) = 2 sin(x)
    np.testing.assert_allclose(jac[0, 0], [1.0, 2 * np.sin(0.5)], rtol=1e-6)
    np.testing.assert_allclose(primal[0, 0], 2 * np.sin(0.5) * 3.0 + 0.5, rtol=1e-6)
    ramp, scaled = Var((3,), eqn_index=0), Var((3,), eqn_index=1)
    iota = Jaxpr([x], [Eqn("iota", (), ramp, {"shape": (3,), "dimension": 0}), Eqn("mul", (ramp, x), scaled, {})],
                 [scaled], [0, 0], False)
    p2 = build_plan(iota, "rev", (0,))
    j2, _ = replay.replay(p2.blob, [np.array([2.0], np.float32)], 3, 1)
    np.testing.assert_allclose(j2[0, :, 0], [0.0, 1.0, 2.0])
