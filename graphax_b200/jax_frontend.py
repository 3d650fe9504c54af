"""Tracing through JAX itself, when JAX is installed.

graphax numbers vertices by the position of an equation in ``jax.make_jaxpr(f)(*xs).jaxpr.eqns``
(reference ``core.py:60-64``, ``core.py:184``).  ``tracer.py`` re-creates that lowering for functions written
against its own ``jnp`` stand-in, because this image has no JAX.  On a machine that does, a user's function is
written against the real ``jax.numpy``; this module lets ``graphax_b200.jacve`` accept it unchanged: the function is
traced by ``jax.make_jaxpr`` (so the numbering is JAX's by construction) and the resulting jaxpr is converted to the
engine's ``tracer.Jaxpr``.  Everything after that - elimination plan, CUDA kernels - is the same path.

``import jax`` happens inside the functions: importing this module never needs JAX.  The conversion only reads the
public jaxpr data model (``eqn.primitive.name``, ``eqn.params``, ``invars`` / ``outvars`` with ``aval.shape``,
``Literal.val``), which is also what the reference walks (``core.py:314-412``).  It is exercised in the tests against
``tests/golden/jaxlite``, the NumPy stand-in for the JAX API that the golden generator uses.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

from .tracer import Eqn, Jaxpr, Literal, Var

# equation parameters the engine's lowering reads (everything else - sharding, precision, ... - is ignored)
_PARAMS = {
    "integer_pow": ("y",), "reduce_sum": ("axes",), "reduce_max": ("axes",), "reduce_min": ("axes",),
    "dot_general": ("dimension_numbers",), "slice": ("start_indices", "limit_indices", "strides"),
    "concatenate": ("dimension",), "broadcast_in_dim": ("shape", "broadcast_dimensions"),
    "transpose": ("permutation",), "reshape": ("new_sizes", "dimensions"), "squeeze": ("dimensions",),
}
# call-like primitives: the reference registers a rule for `pjit` only (primitives.py:576-602): the call is ONE vertex
# whose value is computed and whose partials are dropped (the rule returns an empty list, so core.py:405-410 writes
# no edge).  The converter keeps that: the inner jaxpr travels in the equation's parameters and the lowering
# evaluates it for the value.  `jit` is the same primitive under its newer name.
_CALLS = ("pjit", "jit")
_NO_RULE = ("custom_jvp_call", "custom_vjp_call", "while", "cond", "scan")


def available() -> bool:
    try:
        import jax  # noqa: F401
        return True
    except ImportError:
        return False


def _plain(v):
    """Tuples of ints instead of whatever sequence type the jaxpr carries."""
    if isinstance(v, (tuple, list)):
        return tuple(_plain(x) for x in v)
    if isinstance(v, (np.integer,)):
        return int(v)
    return v


def from_jax(closed_jaxpr, out_is_sequence: bool = True) -> Jaxpr:
    """Convert ``jax.make_jaxpr(f)(*xs)`` to the engine's jaxpr; equation i stays vertex i."""
    import jax
    jaxpr = closed_jaxpr.jaxpr
    literal_cls = jax.core.Literal
    env: Dict[int, Var] = {}
    invars = []
    for i, v in enumerate(jaxpr.invars):
        env[id(v)] = Var(tuple(v.aval.shape), input_index=i)
        invars.append(env[id(v)])
    # arrays the function closes over (core.py:374 writes them into the environment next to the arguments)
    constvars, consts = [], []
    for v, val in zip(getattr(jaxpr, "constvars", ()) or (), getattr(closed_jaxpr, "consts", ()) or ()):
        env[id(v)] = Var(tuple(v.aval.shape), input_index=len(invars) + len(constvars), name=f"const{len(constvars)}")
        constvars.append(env[id(v)])
        consts.append(np.ascontiguousarray(np.asarray(val), dtype=np.float32))
    if len(constvars) != len(getattr(jaxpr, "constvars", ()) or ()):
        raise NotImplementedError("jaxpr with constvars but without their values (pass the ClosedJaxpr)")

    def atom(a):
        if isinstance(a, literal_cls):
            if np.shape(a.val) != ():
                raise NotImplementedError("array-valued literals are not lowered")
            return Literal(float(a.val))
        return env[id(a)]

    eqns: List[Eqn] = []
    pads: List[int] = []
    for eqn in jaxpr.eqns:
        name = eqn.primitive.name
        if name in _NO_RULE or len(eqn.outvars) != 1:
            # no elemental rule in the reference either (core.py:396-398 raises the same error), or several results
            raise NotImplementedError(f"{name} does not have registered elemental partial.")
        ins = tuple(atom(a) for a in eqn.invars)
        out = Var(tuple(eqn.outvars[0].aval.shape), eqn_index=len(eqns))
        env[id(eqn.outvars[0])] = out
        if name == "add_any":          # jax._src.ad_util.add_any_p: the reference registers the add rule for it
            name = "add"
        if name in _CALLS:
            params = {"jaxpr": from_jax(eqn.params["jaxpr"]), "name": str(eqn.params.get("name", ""))}
            name = "pjit"
        elif name == "iota":
            params = {"shape": _plain(eqn.params["shape"]), "dimension": int(eqn.params["dimension"])}
        else:
            params = {k: _plain(eqn.params[k]) for k in _PARAMS.get(name, ()) if k in eqn.params}
        eqns.append(Eqn(name, ins, out, params))
        pads.append(_vmap_pads(name, ins, eqns))
    outvars = [atom(v) for v in jaxpr.outvars]
    return Jaxpr(invars, eqns, outvars, pads, out_is_sequence, constvars, consts)


def _vmap_pads(name: str, ins: Sequence, eqns: List[Eqn]) -> int:
    """How many identity ``broadcast_in_dim`` equations ``jax.vmap`` inserts before this one (SURVEY.md A.1 rule 5):
    the lower-rank operand of a mixed-rank binary operation, and every un-batched operand of a concatenate."""
    traced = [a for a in ins if isinstance(a, Var)]
    if name == "concatenate":
        return sum(1 for a in traced if a.eqn_index is not None and eqns[a.eqn_index].primitive == "broadcast_in_dim"
                   and all(isinstance(x, Literal) for x in eqns[a.eqn_index].invars))
    if len(ins) == 2 and len(traced) == 2 and len(traced[0].shape) != len(traced[1].shape):
        return 1
    return 0


def make_jaxpr(fun: Callable, arg_shapes: Sequence[Tuple[int, ...]]) -> Jaxpr:
    """``jax.make_jaxpr`` on float32 avals of the per-sample shapes, converted."""
    import jax
    import jax.numpy as jnp
    specs = [jax.ShapeDtypeStruct(tuple(s), jnp.float32) for s in arg_shapes]
    return from_jax(jax.make_jaxpr(fun)(*specs))
