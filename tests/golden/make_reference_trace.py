"""Structure traces of the reference's OWN code on graphs with matrix-valued vertices.

Run once in the build container (where /root/reference exists):

    python tests/golden/make_reference_trace.py

Same set-up as make_reference_goldens.py (the unmodified reference sources on the NumPy stand-in for JAX).  For the
reference's NeuralNetwork / Perceptron / Encoder / EncoderDecoder tests, the batched MLP and `examples/randoms.py:f`
this records, per elimination order:

  * the SparseTensor structure of every elemental edge (`core.py:314-412`),
  * `num_muls`, `num_adds` and the per-vertex `order_counts` of `count_ops=True`,
  * the structure of every final Jacobian block (`sparse_representation=True`),
  * for the smaller graphs a TRACE of every `_mul` / `_add` the elimination performs - the structures of both
    operands and of the result, and what `get_num_muls` / `get_num_adds` returned for it,

i.e. the known answers for the engine's rank-N structure algebra (graphax_b200/structure_nd.py), which restates the
reference's dimension bookkeeping (`sparse/tensor.py:334-397, 826-1231, 1267-1493`, `primitives.py:42-107, 243-441,
606-1255`) without values.  Output: tests/golden/reference_trace_rank_n.json.gz.
"""
import gzip
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REFERENCE = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src")
sys.path.insert(0, str(REPO))
sys.path[:0] = [str(HERE / "jaxlite"), str(REFERENCE)]
import jax                      # noqa: E402  (jaxlite)
import jax.numpy as jnp         # noqa: E402
import graphax                  # noqa: E402  (the reference)
import graphax.core as gcore    # noqa: E402
import graphax.sparse.tensor as gtensor   # noqa: E402
from graphax import examples as gex       # noqa: E402
from graphax.sparse.tensor import SparseDimension, SparseTensor   # noqa: E402

assert str(REFERENCE) in graphax.__file__, "graphax must be the reference's own sources"

from make_reference_goldens import RANK_N, NeuralNetwork, dump_jaxpr   # noqa: E402  (shapes, argnums, points)


def dims(ds):
    return [["S" if isinstance(d, SparseDimension) else "D", d.id, d.size, d.val_dim,
             d.other_id if isinstance(d, SparseDimension) else None] for d in ds]


def describe(t):
    if t is None:
        return None
    return {"out": dims(t.out_dims), "primal": dims(t.primal_dims),
            "val": None if t.val is None else list(np.shape(t.val)), "pre": len(t.pre_transforms)}


TRACE = None      # list being filled while a traced elimination runs


def _install_tracing():
    """Wrap (not modify) the reference's product / sum / cost functions so that every call is logged."""
    orig_mul, orig_add = SparseTensor.__mul__, SparseTensor.__add__
    orig_nm, orig_na = gcore.get_num_muls, gcore.get_num_adds

    def mul(self, other):
        res = orig_mul(self, other)
        if TRACE is not None:
            TRACE.append(["mul", describe(self), describe(other), describe(res)])
        return res

    def add(self, other):
        res = orig_add(self, other)
        if TRACE is not None:
            TRACE.append(["add", describe(self), describe(other), describe(res)])
        return res

    def num_muls(lhs, rhs):
        n = orig_nm(lhs, rhs)
        if TRACE is not None:
            TRACE.append(["num_muls", int(n)])
        return n

    def num_adds(lhs, rhs):
        n = orig_na(lhs, rhs)
        if TRACE is not None:
            TRACE.append(["num_adds", int(n)])
        return n
    SparseTensor.__mul__, SparseTensor.__add__ = mul, add
    gcore.get_num_muls, gcore.get_num_adds = num_muls, num_adds


def mlp_batched():
    from functools import partial

    @partial(jax.vmap, in_axes=(0, None, None, None, None, 0))
    def net(x, W1, b1, W2, b2, y):
        return 0.5*(jnp.tanh(W2 @ jnp.tanh(W1 @ x + b1) + b2) - y)**2

    def f(x, W1, b1, W2, b2, y):
        return net(x, W1, b1, W2, b2, y).sum()
    shapes = [(12, 6), (10, 6), (10,), (4, 10), (4,), (12, 4)]
    return f, shapes, (1, 2, 3, 4)


CASES = {name: ((NeuralNetwork if name == "NeuralNetwork" else getattr(gex, name)),) + RANK_N[name] for name in RANK_N}
CASES["mlp_batched"] = mlp_batched()
CASES["f"] = (gex.f, [(4,), (2, 3), (4, 4), (4, 1)], (0, 1, 2, 3))
TRACED = {"NeuralNetwork": ("fwd", "rev"), "Perceptron": ("fwd", "rev"), "mlp_batched": ("fwd", "rev"),
          "Encoder": ("rev",), "f": ("rev",)}


def run(name):
    global TRACE
    fun, shapes, argnums = CASES[name]
    rng = np.random.default_rng(5)
    args = [jnp.array((rng.normal(0.0, 0.6, s) if s else rng.uniform(0.5, 1.5, ())).astype(np.float32)) for s in shapes]
    if name == "f":
        args = [jnp.array(rng.uniform(0.2, 0.8, s).astype(np.float32)) for s in shapes]
    closed = jax.make_jaxpr(fun)(*args)
    jaxpr = closed.jaxpr
    key, rec = dump_jaxpr(jaxpr)
    rec["argnums"] = list(argnums)
    env, graph, tgraph, vo = gcore._build_graph(jaxpr, args, closed.literals)
    rec["elemental"] = [[key[src], key[dst], describe(t)] for src in graph for dst, t in graph[src].items()]
    rec["orders"] = {}
    for order in ("fwd", "rev"):
        traced = order in TRACED.get(name, ())
        TRACE = [] if traced else None
        try:
            _, aux = graphax.jacve(fun, order=order, argnums=argnums, count_ops=True)(*args)
            trace = TRACE
        finally:
            TRACE = None
        sparse = gcore.vertex_elimination_jaxpr(jaxpr, order, closed.literals, *args, argnums=argnums,
                                                sparse_representation=True)
        if len(argnums) > 1:
            sparse = [t for row in sparse for t in row] if isinstance(sparse[0], (tuple, list)) else list(sparse)
        rec["orders"][order] = {"num_muls": int(aux["num_muls"]), "num_adds": int(aux["num_adds"]),
                                "order_counts": [[int(v), int(c)] for v, c in aux["order_counts"]],
                                "blocks": [describe(t) for t in sparse]}
        if traced:
            rec["orders"][order]["trace"] = trace
        print(f"  {name}:{order}: muls {aux['num_muls']} adds {aux['num_adds']}"
              + (f", {len(trace)} trace records" if traced else ""), flush=True)
    return rec


def main():
    _install_tracing()
    out = {"generator": "tests/golden/make_reference_trace.py",
           "reference": "jamielohoff/graphax, unmodified sources under /root/reference/src, on tests/golden/jaxlite",
           "cases": {}}
    for name in CASES:
        print(name, flush=True)
        out["cases"][name] = run(name)
    path = HERE / "reference_trace_rank_n.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps(out, separators=(",", ":"), sort_keys=True).encode())
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
