"""Golden vectors from the reference's OWN code.

Run once in the build container (where /root/reference exists):

    python tests/golden/make_reference_goldens.py

graphax is pure Python on top of JAX and JAX is not installed here, so this script
puts `tests/golden/jaxlite` (a NumPy stand-in for the JAX API, see its `jax/_impl.py`)
and `/root/reference/src` on `sys.path` and then runs the UNMODIFIED reference:
`jax.make_jaxpr` of the reference's example functions, `graphax.core._build_graph`,
`graphax.jacve(f, order, argnums, count_ops=True)` and
`vertex_elimination_jaxpr(..., sparse_representation=True)`.

What is recorded (tests/golden/reference_run.json + reference_run.npz):
  * the jaxpr of every example function (primitive, operands, literals, parameters);
  * the SparseTensor structure of every elemental edge (`core.py:314-412`);
  * per order: the validated order, `num_muls`, `num_adds`, per-vertex `order_counts`,
    the structure of every final Jacobian block (`sparse_representation=True`);
  * float32 inputs, primal outputs and dense Jacobians at N points per (function, order).

The arithmetic behind the values is NumPy float32, one call per primitive, in the
order the reference's Python issues them; a real JAX run can differ from it only by
XLA's fusion / FMA contraction of those same operations.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REFERENCE = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src")

sys.path.insert(0, str(REPO))
from graphax_b200.configs import WORKLOADS, EXTRA_WORKLOADS   # noqa: E402  (points and orders only)

try:                                     # a box that has the real thing: same script, real JAX / XLA arithmetic
    import jax as _real_jax              # noqa: F401
    BACKEND = f"jax {_real_jax.__version__}"
    sys.path.insert(0, str(REFERENCE))
except ImportError:
    BACKEND = "jaxlite (tests/golden/jaxlite: NumPy float32, one call per primitive)"
    sys.path[:0] = [str(HERE / "jaxlite"), str(REFERENCE)]
import jax                      # noqa: E402  (jaxlite unless JAX is installed)
import jax.numpy as jnp         # noqa: E402
import graphax                  # noqa: E402  (the reference)
import graphax.core as gcore    # noqa: E402
from graphax import examples as gex             # noqa: E402
from graphax.sparse.tensor import SparseDimension   # noqa: E402

assert str(REFERENCE) in graphax.__file__, "graphax must be the reference's own sources"

N_POINTS = {"roe3d": 96, "roe1d": 96, "robotarm": 96}

# workload name (graphax_b200.configs) -> the reference's own function
FUNCTIONS = {
    "roe3d": gex.RoeFlux_3d, "roe1d": gex.RoeFlux_1d, "robotarm": gex.RobotArm_6DOF,
    "simple": gex.Simple, "lighthouse": gex.Lighthouse, "hole": gex.Hole, "helmholtz": gex.Helmholtz,
}


def dims(ds):
    return [["S" if isinstance(d, SparseDimension) else "D", d.id, d.size, d.val_dim,
             d.other_id if isinstance(d, SparseDimension) else None] for d in ds]


def describe(t):
    if t is None:
        return None
    return {"out_dims": dims(t.out_dims), "primal_dims": dims(t.primal_dims),
            "val_shape": None if t.val is None else list(np.shape(t.val)),
            "n_pre_transforms": len(t.pre_transforms), "n_post_transforms": len(t.post_transforms)}


def jsonable(v):
    if isinstance(v, (tuple, list)):
        return [jsonable(x) for x in v]
    if isinstance(v, np.dtype):
        return v.name
    if isinstance(v, np.generic):
        return v.item()
    return v


def dump_jaxpr(jaxpr):
    key = {}
    for i, v in enumerate(list(jaxpr.invars) + list(jaxpr.constvars)):     # constants continue the arguments' numbering
        key[v] = -(i + 1)
    eqns = []
    for n, eqn in enumerate(jaxpr.eqns, start=1):
        key[eqn.outvars[0]] = n
        ops = [["lit", float(a.val)] if isinstance(a, jax.core.Literal) else ["var", key[a]] for a in eqn.invars]
        eqns.append({"primitive": eqn.primitive.name, "operands": ops, "shape": list(eqn.outvars[0].aval.shape),
                     "params": {k: jsonable(v) for k, v in eqn.params.items() if v is not None}})
    rec = {"n_eqns": len(eqns), "eqns": eqns, "outvars": [key[v] for v in jaxpr.outvars],
           "in_shapes": [list(v.aval.shape) for v in jaxpr.invars]}
    if jaxpr.constvars:
        rec["const_shapes"] = [list(v.aval.shape) for v in jaxpr.constvars]
    return key, rec


def pack(out, n_rows_each, n_cols_each):
    """The reference's tuple-of-tuples of blocks -> one [rows, cols] matrix (row-major blocks)."""
    rows = []
    for o, row in zip(n_rows_each, out):
        row = row if isinstance(row, tuple) else (row,)
        rows.append(np.concatenate([np.asarray(b, dtype=np.float32).reshape(o, c) for b, c in zip(row, n_cols_each)], axis=1))
    return np.concatenate(rows, axis=0)


def run_workload(name, w, fun, n_points):
    argnums = tuple(w.argnums)
    base = [np.asarray(w.base[sum(w.sizes[:i]):sum(w.sizes[:i + 1])], np.float32).reshape(s) for i, s in enumerate(w.arg_shapes)]
    pts = w.inputs(n_points - 1)
    points = [[jnp.array(b)] + [jnp.array(p[k]) for k in range(n_points - 1)] for b, p in zip(base, pts)]
    args0 = [p[0] for p in points]

    closed = jax.make_jaxpr(fun)(*args0)
    jaxpr = closed.jaxpr
    key, rec = dump_jaxpr(jaxpr)
    rec["function"] = f"graphax.examples.{fun.__name__}"
    rec["point0"] = [np.asarray(a).reshape(-1).tolist() for a in args0]

    env, graph, tgraph, vo = gcore._build_graph(jaxpr, args0, closed.literals)
    rec["vo_vertices"] = sorted(int(v) for v in vo)
    rec["elemental"] = [[key[src], key[dst], describe(t)] for src in graph for dst, t in graph[src].items()]

    out_sizes = [int(np.prod(v.aval.shape, dtype=np.int64)) for v in jaxpr.outvars]
    in_sizes = [int(np.prod(jaxpr.invars[a].aval.shape, dtype=np.int64)) for a in argnums]
    arrays = {"inputs": np.stack([np.concatenate([np.asarray(p[k]).reshape(-1) for p in points]) for k in range(n_points)]).astype(np.float32)}
    prim = []
    for k in range(n_points):
        vals = fun(*[p[k] for p in points])
        vals = vals if isinstance(vals, (tuple, list)) else (vals,)
        prim.append(np.concatenate([np.asarray(v, np.float32).reshape(-1) for v in vals]))
    arrays["primal"] = np.stack(prim)

    rec["orders"] = {}
    for oname in w.order_names():
        order = w.order(oname)
        order = order if isinstance(order, str) else [int(v) for v in order]
        t0 = time.time()
        jacs = []
        for k in range(n_points):
            out, aux = graphax.jacve(fun, order=order, argnums=argnums, count_ops=True)(*[p[k] for p in points])
            if len(jaxpr.outvars) == 1:
                out = (out,)
            if len(argnums) == 1:
                out = tuple((b,) for b in out)
            jacs.append(pack(out, out_sizes, in_sizes))
            if k == 0:
                aux0 = aux
            else:
                assert aux["num_muls"] == aux0["num_muls"] and aux["order_counts"] == aux0["order_counts"]
        sparse = gcore.vertex_elimination_jaxpr(jaxpr, order, closed.literals, *args0, argnums=argnums,
                                                sparse_representation=True)
        if len(argnums) > 1:
            sparse = [t for row in sparse for t in row]
        n_a = len(argnums)
        rec["orders"][oname] = {
            "order": order, "eliminated": [int(v) for v, _ in aux0["order_counts"]],
            "num_muls": int(aux0["num_muls"]), "num_adds": int(aux0["num_adds"]),
            "order_counts": [[int(v), int(c)] for v, c in aux0["order_counts"]],
            "blocks": {f"{i // n_a},{i % n_a}": describe(t) for i, t in enumerate(sparse)},
        }
        arrays[f"jac/{oname}"] = np.stack(jacs)
        print(f"  {name}:{oname}: muls {aux0['num_muls']} adds {aux0['num_adds']}  "
              f"{n_points} points in {time.time() - t0:.1f} s", flush=True)
    return rec, arrays


def run_random_orders(n_orders=3, n_points=8):
    """The caller-given order is arbitrary: seeded random permutations of the eliminable vertices of the three named
    workloads, through the reference (count_ops, per-vertex counts, block structures, Jacobians)."""
    out, arrays = {}, {}
    for name in ("roe3d", "roe1d", "robotarm"):
        w, fun = WORKLOADS[name], FUNCTIONS[name]
        base = [np.asarray(w.base[sum(w.sizes[:i]):sum(w.sizes[:i + 1])], np.float32).reshape(s) for i, s in enumerate(w.arg_shapes)]
        pts = w.inputs(n_points, start=1000)
        jaxpr = jax.make_jaxpr(fun)(*[jnp.array(b) for b in base]).jaxpr
        outs = {id(v) for v in jaxpr.outvars}
        eliminable = [i for i, e in enumerate(jaxpr.eqns, 1) if id(e.outvars[0]) not in outs]
        argnums = tuple(w.argnums)
        out_sizes = [int(np.prod(v.aval.shape, dtype=np.int64)) for v in jaxpr.outvars]
        arrays[f"random_orders/{name}/inputs"] = np.concatenate([p.reshape(n_points, -1) for p in pts], axis=1)
        out[name] = {}
        for k in range(n_orders):
            order = [int(v) for v in np.random.default_rng(100 * k + len(eliminable)).permutation(eliminable)]
            jacs = []
            for b in range(n_points):
                res, aux = graphax.jacve(fun, order=order, argnums=argnums, count_ops=True)(*[jnp.array(p[b]) for p in pts])
                jacs.append(pack(res, out_sizes, w.sizes))
            sparse = gcore.vertex_elimination_jaxpr(jaxpr, order, [], *[jnp.array(p[0]) for p in pts], argnums=argnums,
                                                    sparse_representation=True)
            sparse = [t for row in sparse for t in row]
            out[name][f"random{k}"] = {"order": order, "num_muls": int(aux["num_muls"]), "num_adds": int(aux["num_adds"]),
                                       "order_counts": [[int(v), int(c)] for v, c in aux["order_counts"]],
                                       "blocks": {f"{i // len(argnums)},{i % len(argnums)}": describe(t) for i, t in enumerate(sparse)}}
            arrays[f"random_orders/{name}/jac/random{k}"] = np.stack(jacs)
            print(f"  {name}:random{k}: muls {aux['num_muls']} adds {aux['num_adds']}")
    return out, arrays


def _random_vector_function(seed, n_ops):
    """A seeded random function of scalars, (1,) and (3,) vectors: elementwise rules, sum, dot, slices, concatenate."""
    def fn(*xs):
        r = np.random.default_rng(seed)
        vals = list(xs)
        for _ in range(n_ops):
            kind = r.choice(["un", "bi", "bi", "sum", "dot", "slice", "concat", "pow2"])
            a, b = vals[r.integers(len(vals))], vals[r.integers(len(vals))]
            v = None
            if kind == "un":
                v = [jnp.sin, jnp.cos, jnp.sqrt, jnp.abs, lambda t: -t, jnp.exp, jnp.tanh][r.integers(7)](a)
            elif kind == "bi":
                if a is b:
                    continue
                if a.shape == b.shape or a.ndim == 0 or b.ndim == 0 or (a.ndim == b.ndim == 1 and 1 in (a.shape[0], b.shape[0])):
                    v = [lambda p, q: p + q, lambda p, q: p * q, lambda p, q: p - q, lambda p, q: p / q][r.integers(4)](a, b)
            elif kind == "sum":
                if a.ndim == 1:
                    v = jnp.sum(a)
            elif kind == "dot":
                if a is not b and a.ndim == b.ndim == 1 and a.shape == b.shape and a.shape[0] > 1:
                    v = jnp.dot(a, b)
            elif kind == "slice":
                if a.ndim == 1 and a.shape[0] == 3:
                    v = a[int(r.integers(0, 2)):int(r.integers(2, 4))]
            elif kind == "concat":
                if a is not b and a.ndim == b.ndim == 1 and a.shape[0] + b.shape[0] <= 4:
                    v = jnp.concatenate([a, b])
            elif kind == "pow2":
                v = a**2
            if v is not None:
                vals.append(v)
        return tuple(vals[len(xs):][-2:])
    return fn


def run_random_functions(wanted=16, n_points=4):
    """Random functions through the reference, forward and reverse: jaxpr, count_ops, Jacobians.  Draws on which the
    reference raises (its slice / concatenate rules do for many graphs) or returns wrongly shaped blocks are skipped
    and counted."""
    out, arrays, skipped, trial = {}, {}, 0, 0
    while len(out) < wanted and trial < 400:
        trial += 1
        rng = np.random.default_rng(7000 + trial)
        shapes = [(1,), (3,), (3,), (1,), ()][:rng.integers(2, 6)]
        fun = _random_vector_function(trial, int(rng.integers(5, 18)))
        pts = [rng.uniform(0.5, 1.5, (n_points,) + s).astype(np.float32) for s in shapes]
        args0 = [jnp.array(p[0]) for p in pts]
        closed = jax.make_jaxpr(fun)(*args0)
        outvars = closed.jaxpr.outvars
        produced = {id(e.outvars[0]) for e in closed.jaxpr.eqns}
        if len(outvars) < 2 or len({id(v) for v in outvars}) < len(outvars) or any(id(v) not in produced for v in outvars):
            continue
        argnums = tuple(range(len(shapes)))
        out_sizes = [int(np.prod(v.aval.shape, dtype=np.int64)) for v in outvars]
        in_sizes = [int(np.prod(s, dtype=np.int64)) if s else 1 for s in shapes]
        rec_orders, jac = {}, {}
        try:
            for order in ("fwd", "rev"):
                vals = []
                for k in range(n_points):
                    res, aux = graphax.jacve(fun, order=order, argnums=argnums, count_ops=True)(*[jnp.array(p[k]) for p in pts])
                    vals.append(pack(res, out_sizes, in_sizes))
                rec_orders[order] = {"num_muls": int(aux["num_muls"]), "num_adds": int(aux["num_adds"]),
                                     "order_counts": [[int(v), int(c)] for v, c in aux["order_counts"]]}
                jac[order] = np.stack(vals)
        except Exception:      # noqa: BLE001  (the reference's own limits)
            skipped += 1
            continue
        if not all(np.isfinite(v).all() for v in jac.values()):
            continue
        _, rec = dump_jaxpr(closed.jaxpr)
        rec["orders"] = rec_orders
        name = f"seed{trial}"
        out[name] = rec
        arrays[f"random_functions/{name}/inputs"] = np.concatenate([p.reshape(n_points, -1) for p in pts], axis=1)
        for order, v in jac.items():
            arrays[f"random_functions/{name}/jac/{order}"] = v
    print(f"  {len(out)} random functions recorded, {skipped} draws skipped because the reference raises on them")
    return {"functions": out, "skipped_reference_raises": skipped, "draws": trial}, arrays


def run_g():
    """The reference's only literal known answers: `# fwd: 632, rev: 566` (`examples/randoms.py:91`)."""
    rng = np.random.default_rng(7)
    xs = [jnp.array(np.float32(x)) for x in rng.uniform(0.5, 1.5, 15)]
    rec = {"function": "graphax.examples.g", "n_eqns": len(jax.make_jaxpr(gex.g)(*xs).jaxpr.eqns), "orders": {}}
    for order in ("fwd", "rev"):
        _, aux = graphax.jacve(gex.g, order=order, argnums=tuple(range(15)), count_ops=True)(*xs)
        rec["orders"][order] = {"num_muls": int(aux["num_muls"]), "num_adds": int(aux["num_adds"]),
                                "order_counts": [[int(v), int(c)] for v, c in aux["order_counts"]]}
        print(f"  g:{order}: muls {aux['num_muls']} (reference comment: {'632' if order == 'fwd' else '566'})")
    return rec


def run_roe3d_vmapped(batch=4):
    """The form the reference's own test uses: `jacve(jax.vmap(RoeFlux_3d), order)` with orders written against
    the 146-equation vmapped jaxpr (`tests/examples/example_test.py:146-166`)."""
    from graphax_b200.examples import orders as ref_orders
    w = WORKLOADS["roe3d"]
    xs = [jnp.array(x) for x in w.inputs(batch)]
    vf = jax.vmap(gex.RoeFlux_3d)
    jaxpr = jax.make_jaxpr(vf)(*xs).jaxpr
    pads = [n for n, e in enumerate(jaxpr.eqns, 1)
            if e.primitive.name == "broadcast_in_dim" and not isinstance(e.invars[0], jax.core.Literal)]
    rec = {"function": "jax.vmap(graphax.examples.RoeFlux_3d)", "batch": batch, "n_eqns": len(jaxpr.eqns),
           "identity_pads": pads, "primitives": [e.primitive.name for e in jaxpr.eqns], "orders": {}}
    arrays = {"inputs": np.concatenate([np.asarray(x).reshape(batch, -1) for x in xs], axis=1).astype(np.float32)}
    orders = {"fwd": "fwd", "rev": "rev", "mixed": list(ref_orders.ROEFLUX_3D_VMAP_MIXED),
              "mixed2": list(ref_orders.ROEFLUX_3D_VMAP_MIXED2)}
    for oname, order in orders.items():
        out, aux = graphax.jacve(vf, order=order, argnums=tuple(range(6)), count_ops=True)(*xs)
        rows = []
        for row in out:      # blocks [B, out, B, in]
            rows.append(np.concatenate([np.asarray(b, np.float32).reshape(batch, -1, batch, c) for b, c in zip(row, w.sizes)], axis=3))
        full = np.concatenate(rows, axis=1)                       # [B, 5, B, 10]
        off = full.copy()
        for b in range(batch):
            off[b, :, b, :] = 0
        rec["orders"][oname] = {"order": order, "num_muls": int(aux["num_muls"]), "num_adds": int(aux["num_adds"]),
                                "off_diagonal_max": float(np.abs(off).max())}
        arrays[f"jac/{oname}"] = np.stack([full[b, :, b, :] for b in range(batch)])
        print(f"  roe3d_vmap:{oname}: muls {aux['num_muls']} adds {aux['num_adds']} off-diagonal max {np.abs(off).max()}")
    return rec, arrays


def run_zoo(n_points=16):
    """Every elementwise elemental rule (`primitives.py:150-239`) through the reference, on the engine's coverage
    function `graphax_b200/examples/zoo.py` (its source is executed here against jaxlite's `jnp`).  Recorded twice:
    with the reference's rules as they are, and with the `asinh` rule replaced through the reference's own extension
    point (`defelemental`): `primitives.py:172` registers sqrt(1 + x^2), the reciprocal of the derivative."""
    import types
    import jax.lax as lax
    from graphax.primitives import defelemental, elemental_rules
    ns = types.SimpleNamespace(**{k: getattr(jnp, k) for k in dir(jnp) if not k.startswith("_")})
    ns.sigmoid, ns.erf = lax.logistic, jax.scipy.special.erf          # conveniences of the engine's jnp shim
    scope = {"jnp": ns}
    exec((REPO / "graphax_b200" / "examples" / "zoo.py").read_text().replace("from ..tracer import jnp", ""), scope)
    zoo = scope["ElementalZoo"]
    w = EXTRA_WORKLOADS["zoo"]
    pts = w.inputs(n_points)
    rec = {"function": "graphax_b200.examples.zoo.ElementalZoo (engine fixture, reference rules)", "orders": {}}
    arrays = {"inputs": np.concatenate([p.reshape(n_points, -1) for p in pts], axis=1)}
    jaxpr = jax.make_jaxpr(zoo)(*[jnp.array(p[0]) for p in pts]).jaxpr
    rec["primitives"] = [e.primitive.name for e in jaxpr.eqns]
    original = elemental_rules[lax.asinh_p]
    for variant in ("reference_rules", "asinh_corrected"):
        if variant == "asinh_corrected":
            defelemental(lax.asinh_p, lambda x: 1. / jnp.sqrt(1. + x**2))
        for order in ("fwd", "rev"):
            vals = []
            for k in range(n_points):
                out, aux = graphax.jacve(zoo, order=order, argnums=(0, 1, 2), count_ops=True)(*[jnp.array(p[k]) for p in pts])
                vals.append(pack(out, [1, 3, 1], w.sizes))
            arrays[f"{variant}/jac/{order}"] = np.stack(vals)
            rec["orders"][order] = {"num_muls": int(aux["num_muls"]), "num_adds": int(aux["num_adds"])}
    elemental_rules[lax.asinh_p] = original
    # a second quirk, recorded as a known answer: both partials of `x*x` land on the same graph edge
    # (`core.py:376-378` assigns graph[invar][outvar] twice), so the reference returns x instead of 2x
    dup = graphax.jacve(lambda x: x * x, order="fwd", argnums=(0,))(jnp.array(3.0))
    rec["quirks"] = {"d(x*x)/dx at x=3": float(np.asarray(dup[0] if isinstance(dup, (tuple, list)) else dup))}
    # a third: the `select_n` rule (`primitives.py:225-234`) builds an inconsistent SparseTensor, so the reference
    # cannot differentiate through `jnp.where` at all
    try:
        graphax.jacve(lambda x: jnp.where(x > 0., x, 0.1 * x), order="rev", argnums=(0,))(jnp.array([0.5, -0.3, 1.2]))
        rec["quirks"]["jnp.where"] = "ok"
    except Exception as e:       # noqa: BLE001
        rec["quirks"]["jnp.where"] = type(e).__name__
    d = np.abs(arrays["reference_rules/jac/rev"] - arrays["asinh_corrected/jac/rev"]).max()
    print(f"  zoo: muls {rec['orders']} ; asinh rule changes the Jacobian by up to {d:.3f}")
    return rec, arrays


def run_examples(n_points=4):
    """The remaining scalar examples of the reference (`examples/easy.py`, `minpack.py`, `economics.py`): jaxpr,
    `count_ops` and Jacobians, forward and reverse.  The tests rebuild each function from its recorded jaxpr."""
    from graphax.examples import minpack
    cases = {
        "KerrSenn_metric": (gex.KerrSenn_metric, [1.0, 2.0, 0.7, 0.3]),
        "FreeEnergy": (gex.FreeEnergy, [[0.05, 0.15, 0.25, 0.35]]),
        "CloudSchemes_step": (gex.CloudSchemes_step, [0.8, 1.1, 0.9, 1.2, 0.7, 1.3, 0.6, 1.4, 0.5, 1.5, 1.05]),
        "PropaneCombustion": (gex.PropaneCombustion, [0.9, 1.1, 1.3, 0.7, 1.2, 0.8, 1.4, 0.6, 1.5, 0.5, 2.0]),
        "HumanHeartDipole": (gex.HumanHeartDipole, [0.3, 0.7, 0.4, 0.6, 1.1, 0.9, 1.3, 0.8]),
        "SteadyStateCombustion": (minpack.SteadyStateCombustion, [0.1, 0.2, 0.3, 0.4]),
        "BlackScholes": (gex.BlackScholes, [1.0, 0.9, 0.05, 0.3, 1.5]),
    }
    out, arrays = {}, {}
    rng = np.random.default_rng(5)
    for name, (fun, base) in cases.items():
        base = [np.asarray(b, np.float32) for b in base]
        pts = [[jnp.array(b * (1 + 0.02 * rng.uniform(-1, 1, b.shape)).astype(np.float32)) for b in base] for _ in range(n_points)]
        closed = jax.make_jaxpr(fun)(*pts[0])
        _, rec = dump_jaxpr(closed.jaxpr)
        rec["function"] = f"graphax.examples.{fun.__name__}"
        argnums = tuple(range(len(base)))
        out_sizes = [int(np.prod(v.aval.shape, dtype=np.int64)) for v in closed.jaxpr.outvars]
        in_sizes = [int(b.size) for b in base]
        arrays[f"examples/{name}/inputs"] = np.stack([np.concatenate([np.asarray(a).reshape(-1) for a in p]) for p in pts])
        rec["orders"] = {}
        for order in ("fwd", "rev"):
            jacs = []
            for p in pts:
                res, aux = graphax.jacve(fun, order=order, argnums=argnums, count_ops=True)(*p)
                if len(out_sizes) == 1:
                    res = (res,)
                if len(argnums) == 1:
                    res = tuple(r if isinstance(r, tuple) else (r,) for r in res)
                jacs.append(pack(res, out_sizes, in_sizes))
            rec["orders"][order] = {"num_muls": int(aux["num_muls"]), "num_adds": int(aux["num_adds"])}
            arrays[f"examples/{name}/jac/{order}"] = np.stack(jacs)
        out[name] = rec
        print(f"  {name}: {rec['n_eqns']} eqns, {rec['orders']}")
    return out, arrays


def run_closures(n_points=4):
    """Functions that close over arrays (`jaxpr.constvars`, core.py:374): a least-squares residual with a scalar and a
    vector output, and the reference's own Gaussian data fitting residual (`examples/minpack.py:fi`) on the data of
    graphax_b200.examples.minpack (the reference draws its data with jax.random)."""
    from graphax.examples import minpack
    from graphax_b200.examples import minpack as mine
    X = jnp.array((np.arange(15, dtype=np.float32).reshape(5, 3) % 7 - 3) / 4)
    y = jnp.array(np.array([0.5, -1.0, 0.25, 2.0, -0.75], np.float32))
    scale = jnp.array(np.array([1.0, 0.5, 2.0, -1.5, 0.25], np.float32))

    def least_squares(w):
        r = X @ w - y
        weighted = scale * r
        return jnp.sum(r * weighted), weighted        # (not r * r: the reference halves d(x*x), core.py:376-378)

    data, tis = jnp.array(mine.FIT_DATA), jnp.array(mine.FIT_TIMES)

    def gaussian_fit(x1, x2, x3, x4, x5, x6, x7, x8, x9, x10, x11):
        return minpack.fi(x1, x2, x3, x4, x5, x6, x7, x8, x9, x10, x11, data, tis)

    cases = {
        "least_squares": (least_squares, [[0.3, -0.2, 0.5]]),
        "GaussianDataFitting": (gaussian_fit, [1.3, 0.65, 0.65, 0.7, 0.6, 3.0, 5.0, 7.0, 2.0, 4.5, 5.5]),   # MINPACK-2 start
    }
    out, arrays = {}, {}
    rng = np.random.default_rng(11)
    for name, (fun, base) in cases.items():
        base = [np.asarray(b, np.float32) for b in base]
        pts = [[jnp.array(b * (1 + 0.02 * rng.uniform(-1, 1, b.shape)).astype(np.float32)) for b in base] for _ in range(n_points)]
        closed = jax.make_jaxpr(fun)(*pts[0])
        _, rec = dump_jaxpr(closed.jaxpr)
        for i, c in enumerate(closed.consts):
            arrays[f"closures/{name}/const{i}"] = np.asarray(c, dtype=np.float32)
        argnums = tuple(range(len(base)))
        out_sizes = [int(np.prod(v.aval.shape, dtype=np.int64)) for v in closed.jaxpr.outvars]
        in_sizes = [int(b.size) for b in base]
        arrays[f"closures/{name}/inputs"] = np.stack([np.concatenate([np.asarray(a).reshape(-1) for a in p]) for p in pts])
        rec["orders"] = {}
        for order in ("fwd", "rev"):
            jacs = []
            for p in pts:
                res, aux = graphax.jacve(fun, order=order, argnums=argnums, count_ops=True)(*p)
                if len(out_sizes) == 1:
                    res = (res,)
                if len(argnums) == 1:
                    res = tuple(r if isinstance(r, tuple) else (r,) for r in res)
                jacs.append(pack(res, out_sizes, in_sizes))
            rec["orders"][order] = {"num_muls": int(aux["num_muls"]), "num_adds": int(aux["num_adds"]),
                                    "order_counts": [[int(a), int(b)] for a, b in aux["order_counts"]]}
            arrays[f"closures/{name}/jac/{order}"] = np.stack(jacs)
        out[name] = rec
        print(f"  {name}: {rec['n_eqns']} eqns, {len(closed.consts)} constants, "
              f"{ {k: (v['num_muls'], v['num_adds']) for k, v in rec['orders'].items()} }")
    return out, arrays


PYTREE_CASES = {   # name: (source of the function over `jnp`, per-sample shapes, argnums)
    "1out_1in": ("lambda x: jnp.sin(x) * x", [(3,)], (0,)),
    "1out_2in_both": ("lambda x, y: jnp.sin(x) * y", [(3,), (3,)], (0, 1)),
    "1out_2in_first": ("lambda x, y: jnp.sin(x) * y", [(3,), (3,)], (0,)),
    "1out_2in_second": ("lambda x, y: jnp.sin(x) * y", [(3,), (3,)], (1,)),
    "2out_1in": ("lambda x: (jnp.sin(x), jnp.sum(x * 2.))", [(3,)], (0,)),
    "2out_2in_both": ("lambda x, y: (jnp.sin(x) * y, jnp.sum(x) * jnp.cos(y))", [(3,), ()], (0, 1)),
    "2out_2in_first": ("lambda x, y: (jnp.sin(x) * y, jnp.sum(x) * jnp.cos(y))", [(3,), ()], (0,)),
    "scalar_1out_1in": ("lambda x: jnp.sin(x) * x", [()], (0,)),
    "independent_block": ("lambda x, y: (jnp.sin(x), jnp.cos(y))", [(3,), (2,)], (0, 1)),
}


def tree_shape(t):
    if isinstance(t, tuple):
        return {"tuple": [tree_shape(c) for c in t]}
    if isinstance(t, list):
        return {"list": [tree_shape(c) for c in t]}
    if isinstance(t, dict):
        return {"dict": sorted(t)}
    return list(np.shape(t))


def run_pytree_cases():
    """Result conventions of `jacve` / `jacfun` (`core.py:60-93`, `:555-581`): nesting and block shapes for every
    combination of output count, argument count and `argnums`, without and with `has_aux` / `count_ops`."""
    out = {}
    for name, (src, shapes, argnums) in PYTREE_CASES.items():
        fun = eval(src, {"jnp": jnp})
        args = [jnp.array(np.linspace(0.3, 0.9, int(np.prod(s, dtype=np.int64)) if s else 1).reshape(s).astype(np.float32)) for s in shapes]
        rec = {"source": src, "in_shapes": [list(s) for s in shapes], "argnums": list(argnums)}
        rec["plain"] = tree_shape(graphax.jacve(fun, order="rev", argnums=argnums)(*args))
        rec["has_aux"] = tree_shape(graphax.jacve(fun, order="rev", argnums=argnums, has_aux=True)(*args))
        rec["count_ops"] = tree_shape(graphax.jacve(fun, order="rev", argnums=argnums, count_ops=True)(*args))
        out[name] = rec
        print(f"  {name}: {rec['plain']}")
    return out


RANK_N = {   # per-sample shapes and argnums of the reference's own tests (tests/examples/example_test.py:168-349)
    "NeuralNetwork": ([(4,), (8, 4), (8,), (4, 8), (4,), (4,)], (1, 2, 3, 4)),
    "Perceptron": ([(4,), (4,), (8, 4), (8,), (4, 8), (4,), (), ()], tuple(range(8))),
    "Encoder": ([(4, 4), (2, 4)] + [(4, 4)] * 6 + [(4, 4), (2, 4), (4,), (2, 1)] + [()] * 4, tuple(range(2, 16))),
    "EncoderDecoder": ([(4, 4), (2, 4)] + [(4, 4)] * 9 + [(4, 4), (2, 4), (4,), (2, 1)] + [()] * 6, tuple(range(2, 21))),
}


def NeuralNetwork(x, W1, b1, W2, b2, y):
    """The function defined inside the reference's `test_NeuralNetwork` (example_test.py:170-179)."""
    a1 = jnp.tanh(W1 @ x + b1)
    a2 = jnp.tanh(W2 @ a1 + b2)
    d = a2 - y
    return .5*jnp.sum(d**2)


def run_rank_n(name, n_points=2):
    """Matrix-valued vertices (`_mixed_mul`, dot_general / transpose / reduce_max edges): values, counts, equations."""
    shapes, argnums = RANK_N[name]
    fun = NeuralNetwork if name == "NeuralNetwork" else getattr(gex, name)
    rng = np.random.default_rng(11)
    pts = [rng.normal(0.0, 0.6, (n_points,) + s) if s else rng.uniform(0.5, 1.5, (n_points,)) for s in shapes]
    pts = [p.astype(np.float32) for p in pts]
    jaxpr = jax.make_jaxpr(fun)(*[jnp.array(p[0]) for p in pts]).jaxpr
    rec = {"function": name, "in_shapes": [list(s) for s in shapes], "argnums": list(argnums), "n_eqns": len(jaxpr.eqns),
           "primitives": [e.primitive.name for e in jaxpr.eqns], "out_shapes": [list(v.aval.shape) for v in jaxpr.outvars],
           "orders": {}}
    arrays = {f"in{i}": p for i, p in enumerate(pts)}
    for order in ("fwd", "rev"):
        vals = []
        for k in range(n_points):
            out, aux = graphax.jacve(fun, order=order, argnums=argnums, count_ops=True)(*[jnp.array(p[k]) for p in pts])
            rows = out if len(jaxpr.outvars) > 1 else (out,)
            vals.append(np.concatenate([np.concatenate([np.asarray(b, np.float32).reshape(-1, int(np.prod(shapes[a], dtype=np.int64)))
                                                        for b, a in zip(row, argnums)], axis=1) for row in rows], axis=0))
        rec["orders"][order] = {"num_muls": int(aux["num_muls"]), "num_adds": int(aux["num_adds"])}
        arrays[f"jac/{order}"] = np.stack(vals)
        print(f"  {name}:{order}: muls {aux['num_muls']} adds {aux['num_adds']} tile {vals[0].shape}")
    return rec, arrays


def run_mlp_batched(batch=12, widths=(6, 10, 4)):
    """The batched MLP of the reference's `test_vmap_NeuralNetwork` (example_test.py:204-215: the per-sample network
    under `jax.vmap(in_axes=(0, None, None, None, None, 0))`, summed to a scalar loss) at NON-SQUARE widths, so that
    the shape comparisons in the reference's `_swap_axes` (sparse/tensor.py:529-569) are unambiguous: weight
    Jacobians of a scalar loss through `_mixed_mul` contractions over the batch.  This pins the tensor-valued path
    (graphax_b200/dense.py + oracle/dense_numpy.py) to a run of the reference's own code."""
    from functools import partial
    n_in, n_hidden, n_out = widths

    @partial(jax.vmap, in_axes=(0, None, None, None, None, 0))
    def NeuralNetwork(x, W1, b1, W2, b2, y):
        y1 = W1 @ x
        z1 = y1 + b1
        a1 = jnp.tanh(z1)
        y2 = W2 @ a1
        z2 = y2 + b2
        return 0.5*(jnp.tanh(z2) - y)**2

    def f(x, W1, b1, W2, b2, y):
        out = NeuralNetwork(x, W1, b1, W2, b2, y)
        return out.sum()

    rng = np.random.default_rng(23)
    shapes = [(batch, n_in), (n_hidden, n_in), (n_hidden,), (n_out, n_hidden), (n_out,), (batch, n_out)]
    args = [(rng.standard_normal(sh) * sc).astype(np.float32) for sh, sc in
            zip(shapes, (1.0, n_in ** -0.5, 0.1, n_hidden ** -0.5, 0.1, 1.0))]
    argnums = (1, 2, 3, 4)
    closed = jax.make_jaxpr(f)(*[jnp.array(a) for a in args])
    rec = {"function": "sum(vmap(NeuralNetwork, in_axes=(0, None, None, None, None, 0))) (example_test.py:204-215)",
           "in_shapes": [list(sh) for sh in shapes], "argnums": list(argnums), "n_eqns": len(closed.jaxpr.eqns),
           "primitives": [e.primitive.name for e in closed.jaxpr.eqns], "orders": {}}
    arrays = {f"in{i}": a for i, a in enumerate(args)}
    for order in ("fwd", "rev"):
        (loss, grads), = [graphax.jacve(f, order=order, argnums=argnums, has_aux=True)(*[jnp.array(a) for a in args])]
        _, aux = graphax.jacve(f, order=order, argnums=argnums, count_ops=True)(*[jnp.array(a) for a in args])
        rec["orders"][order] = {"num_muls": int(aux["num_muls"]), "num_adds": int(aux["num_adds"])}
        arrays[f"loss/{order}"] = np.asarray(loss, np.float32)
        for a, g in zip(argnums, grads):
            assert tuple(np.shape(g)) == tuple(shapes[a]), (np.shape(g), shapes[a])
            arrays[f"grad{a}/{order}"] = np.asarray(g, np.float32)
        print(f"  mlp_batched:{order}: muls {aux['num_muls']} adds {aux['num_adds']} loss {float(loss):.6f}")
    return rec, arrays


def run_f(n_points=2):
    """`examples/randoms.py:f` (80 equations on vectors and matrices: dot_general in four layouts, transpose, reshape,
    squeeze, reduce_max / reduce_min, slice, stop_gradient; its test is commented out in the reference).  It calls
    arcsinh twice, so it is run with the `asinh` rule corrected through `defelemental` (see `run_zoo`)."""
    import jax.lax as lax
    from graphax.primitives import defelemental, elemental_rules
    shapes = [(4,), (2, 3), (4, 4), (4, 1)]
    rng = np.random.default_rng(3)
    pts = [rng.uniform(0.2, 0.8, (n_points,) + s).astype(np.float32) for s in shapes]
    closed = jax.make_jaxpr(gex.f)(*[jnp.array(p[0]) for p in pts])
    _, rec = dump_jaxpr(closed.jaxpr)
    rec["function"] = "graphax.examples.f (asinh rule corrected)"
    arrays = {f"in{i}": p for i, p in enumerate(pts)}
    original = elemental_rules[lax.asinh_p]
    defelemental(lax.asinh_p, lambda x: 1. / jnp.sqrt(1. + x**2))
    rec["orders"] = {}
    for order in ("fwd", "rev"):
        vals = []
        for k in range(n_points):
            out, aux = graphax.jacve(gex.f, order=order, argnums=(0, 1, 2, 3), count_ops=True)(*[jnp.array(p[k]) for p in pts])
            vals.append(np.concatenate([np.concatenate([np.asarray(b, np.float32).reshape(max(1, int(np.prod(o.aval.shape, dtype=np.int64))), -1)
                                                        for b in row], axis=1) for o, row in zip(closed.jaxpr.outvars, out)], axis=0))
        rec["orders"][order] = {"num_muls": int(aux["num_muls"]), "num_adds": int(aux["num_adds"])}
        arrays[f"jac/{order}"] = np.stack(vals)
    elemental_rules[lax.asinh_p] = original
    print(f"  f: {rec['n_eqns']} eqns {rec['orders']}")
    return rec, arrays


def main():
    out = {"generator": "tests/golden/make_reference_goldens.py",
           "reference": "jamielohoff/graphax, unmodified sources under /root/reference/src",
           "backend": BACKEND, "workloads": {}}
    arrays = {}
    for name, fun in FUNCTIONS.items():
        w = WORKLOADS.get(name) or EXTRA_WORKLOADS[name]
        print(name, flush=True)
        rec, arr = run_workload(name, w, fun, N_POINTS.get(name, 8))
        out["workloads"][name] = rec
        for k, v in arr.items():
            arrays[f"{name}/{k}"] = v
    out["rank_n"] = {}
    for name in RANK_N:
        print(name, flush=True)
        rec, arr = run_rank_n(name)
        out["rank_n"][name] = rec
        for k, v in arr.items():
            arrays[f"rank_n/{name}/{k}"] = v
    print("random orders", flush=True)
    out["random_orders"], arr = run_random_orders()
    arrays.update(arr)
    print("random functions", flush=True)
    out["random_functions"], arr = run_random_functions()
    arrays.update(arr)
    print("further examples", flush=True)
    out["examples"], arr = run_examples()
    arrays.update(arr)
    print("closures", flush=True)
    out["closures"], arr = run_closures()
    arrays.update(arr)
    print("pytree conventions", flush=True)
    out["pytree"] = run_pytree_cases()
    print("zoo", flush=True)
    rec, arr = run_zoo()
    out["zoo"] = rec
    for k, v in arr.items():
        arrays[f"zoo/{k}"] = v
    print("f", flush=True)
    rec, arr = run_f()
    out["rank_n_recorded"] = {"f": rec}
    for k, v in arr.items():
        arrays[f"rank_n_recorded/f/{k}"] = v
    print("mlp_batched", flush=True)
    rec, arr = run_mlp_batched()
    out["mlp_batched"] = rec
    for k, v in arr.items():
        arrays[f"mlp_batched/{k}"] = v
    print("roe3d_vmap", flush=True)
    rec, arr = run_roe3d_vmapped()
    out["workloads"]["roe3d_vmap"] = rec
    for k, v in arr.items():
        arrays[f"roe3d_vmap/{k}"] = v
    print("g", flush=True)
    out["workloads"]["g"] = run_g()
    (HERE / "reference_run.json").write_text(json.dumps(out, separators=(",", ":")))
    np.savez_compressed(HERE / "reference_run.npz", **arrays)
    print("wrote", HERE / "reference_run.json", (HERE / "reference_run.json").stat().st_size, "bytes;",
          HERE / "reference_run.npz", (HERE / "reference_run.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
