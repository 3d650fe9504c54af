"""User-extensible elemental rules - the reference's extension point (``primitives.py:110-146``).

graphax keeps a registry ``elemental_rules: Dict[Primitive, rule]`` that ``_build_graph`` consults for every
equation (``core.py:396-400``), filled through ``defelemental(primitive, rule)`` (``rule(*primals)`` returns the
partial derivative w.r.t. every operand), ``defelemental2`` (``rule(out, *primals)``, for functions whose derivative
is cheapest in terms of their value, e.g. ``exp``) or by direct assignment.  Users add primitives, or correct a rule
(the golden generator does exactly that for ``asinh``, whose shipped rule is the reciprocal of the derivative).

This module is that registry for the engine.  The engine never evaluates rules numerically: a rule is TRACED once,
while the plan is built - on stand-ins for the equation's operands, with the same ``graphax_b200.jnp`` functions
traced functions are written in - and the recorded operations are lowered into the plan's scalar expression DAG
(``lowering.lower_eqn``).  The partials then take the same route as the built-in ones: ``make_parallel_jacobian``
(``primitives.py:42-107`` -> ``lowering._parallel_jacobian``) picks the edge structure, the elimination multiplies
them, the kernels evaluate them in registers.

    from graphax_b200 import jnp
    from graphax_b200.primitives import Primitive, defelemental, defelemental2, elemental_rules

    softplus = Primitive("softplus", impl=lambda x: jnp.log1p(jnp.exp(x)))     # a new elementwise primitive
    defelemental(softplus, lambda x: jnp.sigmoid(x))                           # d softplus / dx
    defelemental2("exp", lambda out, x: out)                                   # override a built-in rule

    jac = jacve(lambda x, y: softplus(x * y) + y, order="rev", argnums=(0, 1))(x, y)

Scope: elementwise primitives (the result has the broadcast shape of the operands, one partial per operand of the
operand-or-result shape), which is what ``defelemental`` / ``defelemental2`` define in the reference too
(``standard_elemental`` hands every partial to ``make_parallel_jacobian``).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

from . import tracer as _tr

Rule = Tuple[str, Callable]          # ("primals", fn(*primals)) | ("out", fn(out, *primals))


class Primitive:
    """An elementwise primitive of the traced language.  ``impl(*operands)`` states its value in ``jnp`` operations
    (the role ``primitive.bind`` -> XLA plays in the reference); calling the primitive inside a traced function
    records ONE equation named after it, i.e. one vertex of the computational graph."""
    multiple_results = False

    def __init__(self, name: str, impl: Optional[Callable] = None):
        if not isinstance(name, str) or not name:
            raise TypeError("a primitive needs a non-empty name")
        self.name = name
        self.impl = impl
        _USER_PRIMITIVES[name] = self

    def bind(self, *args, **params):
        trace = _tr._current()
        atoms = [_tr._atom(a) for a in args]
        shape: Tuple[int, ...] = ()
        for a in atoms:
            shape = _tr._bcast_shape(shape, tuple(a.shape))
        return trace.emit(self.name, atoms, shape, **params)

    __call__ = bind

    def __repr__(self) -> str:
        return f"Primitive({self.name!r})"


_USER_PRIMITIVES: Dict[str, Primitive] = {}


def _name(primitive: Union[str, Primitive]) -> str:
    name = primitive if isinstance(primitive, str) else getattr(primitive, "name", None)
    if not isinstance(name, str):
        raise TypeError("primitive must be a graphax_b200.primitives.Primitive or the name of a primitive")
    return name


class _Rules(dict):
    """``elemental_rules``: primitive (object or name) -> rule.  Values stored through ``defelemental`` /
    ``defelemental2`` are tagged rule functions; a plain callable assigned directly is taken as
    ``rule(*primals) -> partials`` (the ``defelemental`` form)."""

    def __setitem__(self, primitive, rule) -> None:
        if callable(rule):
            rule = ("primals", rule)
        if not (isinstance(rule, tuple) and len(rule) == 2 and rule[0] in ("primals", "out") and callable(rule[1])):
            raise TypeError("an elemental rule is a callable returning the partial derivative(s)")
        super().__setitem__(_name(primitive), rule)

    def __getitem__(self, primitive):
        return super().__getitem__(_name(primitive))

    def __contains__(self, primitive) -> bool:
        try:
            return super().__contains__(_name(primitive))
        except TypeError:
            return False

    def __delitem__(self, primitive) -> None:
        super().__delitem__(_name(primitive))

    def pop(self, primitive, *default):
        return super().pop(_name(primitive), *default)


elemental_rules = _Rules()


def defelemental(primitive: Union[str, Primitive], elementalrule: Callable) -> None:
    """``elementalrule(*primals, **params)`` -> partial derivative w.r.t. every operand (a tuple for several
    operands), reference ``primitives.py:113-130``."""
    assert elementalrule is not None
    assert not getattr(primitive, "multiple_results", False)
    elemental_rules[primitive] = ("primals", elementalrule)


def defelemental2(primitive: Union[str, Primitive], elementalrule: Callable) -> None:
    """``elementalrule(out, *primals, **params)``: the rule also sees the primitive's value, reference
    ``primitives.py:132-146``."""
    assert elementalrule is not None
    assert not getattr(primitive, "multiple_results", False)
    elemental_rules[primitive] = ("out", elementalrule)


def lookup(name: str) -> Optional[Tuple[Optional[Primitive], Optional[Rule]]]:
    """(user primitive or None, registered rule or None) when the equation needs this module, else None."""
    prim = _USER_PRIMITIVES.get(name)
    rule = dict.get(elemental_rules, name)
    if prim is None and rule is None:
        return None
    return prim, rule


# ------------------------------------------------------------------------------------------------
# tracing a rule / an implementation and lowering the recorded operations into the plan's DAG
# ------------------------------------------------------------------------------------------------
def _evaluate(fun: Callable, operands: Sequence[Any], params: Dict[str, Any], dag, division: str, what: str) -> List[Any]:
    """Run ``fun`` on stand-ins for ``operands`` (lowering.Vec; literal operands stay Python floats, as in the
    reference where they are Python scalars) and lower what it records.  Returns one lowering.Vec per result."""
    from .lowering import Vec, lower_eqn
    trace = _tr._Trace()
    invars, args = [], []
    for v in operands:
        if v.is_literal:
            args.append(float(dag.cval[v.nodes[0]]))
        else:
            var = _tr.Var(v.shape, input_index=len(invars))
            invars.append((var, v))
            args.append(_tr.Tracer(trace, var))
    _tr._ACTIVE.append(trace)
    try:
        outs = fun(*args, **params)
    finally:
        _tr._ACTIVE.pop()
    outs = outs if isinstance(outs, tuple) else (outs,)
    env: Dict[int, Vec] = {}
    by_input = {id(var): vec for var, vec in invars}

    def read(atom) -> Vec:
        if isinstance(atom, _tr.Literal):
            return Vec((), [dag.const(atom.val)], is_literal=True)
        return by_input[id(atom)] if atom.input_index is not None else env[atom.eqn_index]

    for eqn in trace.eqns:
        # only VALUES are needed here: built-in primitives use the engine's own lowering (a rule for `exp` may call
        # `jnp.exp`), user primitives their implementation
        user = _USER_PRIMITIVES.get(eqn.primitive)
        if user is not None:
            if user.impl is None:
                raise NotImplementedError(f"primitive {eqn.primitive} has no implementation (Primitive(name, impl=...))")
            (out,) = _evaluate(user.impl, [read(a) for a in eqn.invars], eqn.params, dag, division,
                               f"impl of {eqn.primitive}")
        else:
            out, _ = lower_eqn(dag, eqn, read, division, _use_registry=False)
        env[eqn.outvar.eqn_index] = out
    results = []
    for o in outs:
        if isinstance(o, _tr.Tracer):
            results.append(read(o.var))
        elif isinstance(o, (bool, int, float)) or getattr(o, "shape", None) == ():
            results.append(Vec((), [dag.const(float(o))], is_literal=True))
        else:
            raise TypeError(f"{what} must return traced arrays or Python scalars, got {type(o).__name__}")
    return results


def lower_custom(dag, eqn, read, division: str, entry, builtin: Callable):
    """Primal value and elemental edges of an equation whose primitive is user-defined or has a user rule
    (``standard_elemental`` / ``standard_elemental2``).  ``builtin`` lowers the equation with the engine's own rule
    (used for the value of a built-in primitive whose rule was overridden)."""
    from .lowering import _parallel_jacobian
    prim, rule = entry
    ops = [read(a) for a in eqn.invars]
    if prim is not None:
        if prim.impl is None:
            raise NotImplementedError(f"primitive {eqn.primitive} has no implementation (Primitive(name, impl=...))")
        (out,) = _evaluate(prim.impl, ops, eqn.params, dag, division, f"impl of {eqn.primitive}")
        if out.is_literal or out.shape != eqn.outvar.shape:
            idx_shape = eqn.outvar.shape                     # e.g. impl returned a constant: broadcast it
            from .lowering import Vec, _bcast_index
            out = Vec(idx_shape, [out.nodes[int(i)] for i in _bcast_index(out, idx_shape)])
    else:
        out, _ = builtin()
    if rule is None:
        raise NotImplementedError(f"{eqn.primitive} does not have registered elemental partial.")   # core.py:398
    kind, fun = rule
    partials = _evaluate(fun, ([out] if kind == "out" else []) + ops, eqn.params, dag, division,
                         f"elemental rule of {eqn.primitive}")
    if len(partials) != len(ops):
        raise ValueError(f"elemental rule of {eqn.primitive} returned {len(partials)} partial(s) for {len(ops)} operand(s)")
    edges = []
    for i, a in enumerate(eqn.invars):
        if isinstance(a, _tr.Var):                            # partials w.r.t. literal operands are dropped (:127, :145)
            edges.append((a,) + _parallel_jacobian(ops[i], out, partials[i], len(ops)))
    return out, edges
