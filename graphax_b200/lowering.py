"""Scalar expression DAG and the elemental rules, lowered to scalars.

The reference evaluates every equation and its local partial derivatives as
small ``jnp`` arrays (``core.py:380-410``, ``primitives.py:113-205``).  Here the
same formulas are recorded once, per scalar component, as nodes of a hash-consed
expression DAG; the elimination later appends its products to the same DAG and
the whole thing is linearised into the flat plan the CUDA kernels run.

Exactness rules (they are what makes "bit-exact against the plan replay"
possible): every node is one IEEE-754 binary32 operation with round-to-nearest;
identical (op, operands) pairs are the same node (no CSE in the jaxpr, but equal
expressions have equal values); operands of commutative ops are ordered;
``x*1 -> x``, ``x*-1 -> -x``, ``fma(x, +-1, c) -> c +- x`` and all-constant
operations are folded at plan time with the same binary32 arithmetic.
"""
from __future__ import annotations

import math
import warnings
from fractions import Fraction
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import structure as st
from . import structure_nd as nd
from .tracer import Eqn, Literal, Var

# opcode table shared with csrc/gx_ops.h and oracle/replay.c (keep in sync; tests check it)
OPCODES = [
    "input", "const",
    "add", "sub", "mul", "div", "fma", "neg", "abs", "sqrt",
    "sin", "cos", "tan", "atan2", "exp", "log", "log1p", "tanh", "sinh", "cosh",
    "asin", "acos", "atan", "asinh", "acosh", "atanh", "erf", "pow", "logistic",
    "store_jac", "store_primal",
    "max", "min", "gt", "lt", "eq",
    "select",                     # select(which, on_true, on_false): which != 0 ? on_true : on_false
]
OP = {name: i for i, name in enumerate(OPCODES)}
_UNARY = {"neg", "abs", "sqrt", "sin", "cos", "tan", "exp", "log", "log1p", "tanh", "sinh", "cosh",
          "asin", "acos", "atan", "asinh", "acosh", "atanh", "erf", "logistic"}
_COMMUTATIVE = {"add", "mul"}

f32 = np.float32


def _f32_from_fraction(x: Fraction) -> np.float32:
    """Correctly rounded (nearest-even) binary32 value of an exact rational."""
    approx = f32(float(x))
    if not np.isfinite(approx):
        return approx
    best = approx
    for cand in (np.nextafter(approx, f32(-np.inf)), np.nextafter(approx, f32(np.inf))):
        if not np.isfinite(cand):
            continue
        db, dc = abs(Fraction(float(best)) - x), abs(Fraction(float(cand)) - x)
        if dc < db or (dc == db and (int(np.array(cand).view(np.uint32)) & 1) == 0):
            best = cand
    return f32(best)


class Dag:
    """Hash-consed scalar expression DAG. Node ids are dense ints."""

    def __init__(self, fuse: bool = True) -> None:
        # fuse=False ("reference" rounding): every product-and-add is two rounded operations, mul then add, like
        # the reference's separate jnp multiply and add (core.py:210, 262-272); no FMA node is ever created
        self.fuse = bool(fuse)
        self.op: List[str] = []
        self.args: List[Tuple[int, ...]] = []
        self.cval: List[Optional[np.float32]] = []     # value for const nodes
        self._memo: Dict[tuple, int] = {}
        self._mul_twin: Dict[Tuple[int, int], Tuple[int, bool]] = {}   # sign-stripped operands -> (mul node, negated?)

    # ---- node construction -------------------------------------------------
    def _new(self, key: tuple, op: str, args: Tuple[int, ...], cval=None) -> int:
        nid = self._memo.get(key)
        if nid is None:
            nid = len(self.op)
            self.op.append(op)
            self.args.append(args)
            self.cval.append(cval)
            self._memo[key] = nid
        return nid

    def const(self, v) -> int:
        v = f32(v)
        bits = int(np.array(v).view(np.uint32))
        return self._new(("const", bits), "const", (), v)

    def input(self, k: int) -> int:
        return self._new(("input", k), "input", (k,))

    def is_const(self, n: int) -> bool:
        return self.op[n] == "const"

    def _strip_sign(self, n: int) -> Tuple[int, bool]:
        """(node without its leading negation / the constant's magnitude, was it negative?)"""
        if self.op[n] == "neg":
            return self.args[n][0], True
        if self.op[n] == "const" and np.signbit(self.cval[n]) and not np.isnan(self.cval[n]):
            return self.const(-self.cval[n]), True
        return n, False

    def _is(self, n: int, v: float) -> bool:
        return self.op[n] == "const" and float(self.cval[n]) == v and not (v == 0.0)

    def _fold(self, op: str, vals: Sequence[np.float32]) -> Optional[int]:
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            if op == "add":
                return self.const(f32(vals[0]) + f32(vals[1]))
            if op == "sub":
                return self.const(f32(vals[0]) - f32(vals[1]))
            if op == "mul":
                return self.const(f32(vals[0]) * f32(vals[1]))
            if op == "div":
                if float(vals[1]) == 0.0:
                    return None
                return self.const(f32(vals[0]) / f32(vals[1]))
            if op == "neg":
                return self.const(-f32(vals[0]))
            if op == "abs":
                return self.const(np.abs(f32(vals[0])))
            if op == "sqrt":
                return self.const(np.sqrt(f32(vals[0])))
            if op == "fma":
                if all(np.isfinite(v) for v in vals):
                    exact = Fraction(float(vals[0])) * Fraction(float(vals[1])) + Fraction(float(vals[2]))
                    if exact == 0:
                        return None   # sign of an exact zero depends on operand signs; leave to run time
                    return self.const(_f32_from_fraction(exact))
        return None   # transcendental constants are evaluated on the device, not by host libm

    def emit(self, op: str, *a: int) -> int:
        if op == "neg" and self.op[a[0]] == "neg":
            return self.args[a[0]][0]
        if op == "select":
            which, on_true, on_false = a
            if on_true == on_false:
                return on_true
            if self.op[which] == "const":
                return on_true if float(self.cval[which]) != 0.0 else on_false
        if op == "mul":
            x, y = a
            if self._is(y, 1.0):
                return x
            if self._is(x, 1.0):
                return y
            if self._is(y, -1.0):
                return self.emit("neg", x)
            if self._is(x, -1.0):
                return self.emit("neg", y)
            # (-x)*y, x*(-y) and x*y round to the same magnitude (round-to-nearest is symmetric, zeros and
            # infinities included): a product whose sign-stripped twin already exists is that node, negated -
            # a free operand modifier in the instruction that reads it.  The first product of each magnitude
            # is created exactly as written.
            sx, sy = self._strip_sign(x), self._strip_sign(y)
            twin = self._mul_twin.get((min(sx[0], sy[0]), max(sx[0], sy[0])))
            if twin is not None:
                node, node_flip = twin
                return node if node_flip == (sx[1] != sy[1]) else self.emit("neg", node)
        if op == "fma" and not self.fuse:
            return self.emit("add", self.emit("mul", a[0], a[1]), a[2])
        if op == "fma":
            x, y, c = a
            for p, q in ((x, y), (y, x)):
                if self._is(q, 1.0):
                    return self.emit("add", p, c)
                if self._is(q, -1.0):
                    return self.emit("sub", c, p)
            if x > y:
                a = (y, x, c)
        if op in _COMMUTATIVE and a[0] > a[1]:
            a = (a[1], a[0])
        if all(self.op[n] == "const" for n in a):
            folded = self._fold(op, [self.cval[n] for n in a])
            if folded is not None:
                return folded
        nid = self._new((op,) + tuple(a), op, tuple(a))
        if op == "mul":
            sx, sy = self._strip_sign(a[0]), self._strip_sign(a[1])
            self._mul_twin.setdefault((min(sx[0], sy[0]), max(sx[0], sy[0])), (nid, sx[1] != sy[1]))
        return nid

    # convenience
    def add(self, a, b): return self.emit("add", a, b)
    def sub(self, a, b): return self.emit("sub", a, b)
    def mul(self, a, b): return self.emit("mul", a, b)
    def div(self, a, b): return self.emit("div", a, b)
    def fma(self, a, b, c): return self.emit("fma", a, b, c)
    def neg(self, a): return self.emit("neg", a)


class Vec:
    """A traced value lowered to scalar nodes: any shape, components in row-major order."""
    __slots__ = ("shape", "nodes", "is_literal")

    def __init__(self, shape: Tuple[int, ...], nodes: Sequence[int], is_literal: bool = False):
        self.shape = tuple(int(s) for s in shape)
        self.nodes = list(nodes)
        self.is_literal = is_literal
        assert len(self.nodes) == self.size

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def index_array(self) -> np.ndarray:
        """Flat component indices arranged in the value's shape (the currency of every shape op)."""
        return np.arange(self.size, dtype=np.int64).reshape(self.shape)


def _bshape(a: Tuple[int, ...], b: Tuple[int, ...]) -> Tuple[int, ...]:
    if a == ():
        return b
    if b == ():
        return a
    return tuple(max(x, y) for x, y in zip(a, b))


def _bcast_index(v: "Vec", shape: Tuple[int, ...]) -> np.ndarray:
    """For every component of a result of ``shape`` (flat order): the component of ``v`` it reads
    (lax broadcasting: equal rank with size-1 axes stretched, or a rank-0 operand)."""
    if v.shape == tuple(shape):
        return np.arange(v.size, dtype=np.int64)
    if v.ndim == 0:
        return np.zeros(int(np.prod(shape, dtype=np.int64)) if shape else 1, dtype=np.int64)
    return np.broadcast_to(v.index_array(), shape).reshape(-1)


class _Ew:
    """Elementwise helpers with lax broadcasting over Vec."""

    def __init__(self, dag: Dag):
        self.dag = dag

    def lit(self, v: float) -> Vec:
        return Vec((), [self.dag.const(v)], is_literal=True)

    def binary(self, op: str, a: Vec, b: Vec) -> Vec:
        shape = _bshape(a.shape, b.shape)
        ia, ib = _bcast_index(a, shape), _bcast_index(b, shape)
        return Vec(shape, [self.dag.emit(op, a.nodes[i], b.nodes[j]) for i, j in zip(ia, ib)])

    def unary(self, op: str, a: Vec) -> Vec:
        return Vec(a.shape, [self.dag.emit(op, x) for x in a.nodes])

    def add(self, a, b): return self.binary("add", a, b)
    def sub(self, a, b): return self.binary("sub", a, b)
    def mul(self, a, b): return self.binary("mul", a, b)
    def div(self, a, b): return self.binary("div", a, b)
    def neg(self, a): return self.unary("neg", a)

    def ones_like(self, a: Vec) -> Vec:
        return Vec(a.shape, [self.dag.const(1.0)] * a.size)

    def ipow(self, a: Vec, y: int) -> Vec:
        """lax.integer_pow by repeated squaring (y=2 -> x*x, y=3 -> (x*x)*x)."""
        if y == 0:
            return self.ones_like(a)
        if y < 0:
            return self.div(self.lit(1.0), self.ipow(a, -y))
        result, base, e = None, a, y
        while e:
            if e & 1:
                result = base if result is None else self.mul(result, base)
            e >>= 1
            if e:
                base = self.mul(base, base)
        return result


Entries = Dict[Tuple[int, int], int]


def _parallel_jacobian(primal: Vec, out: Vec, partial: Vec, n_operands: int) -> Tuple[Optional[st.Struct], Entries]:
    """``make_parallel_jacobian`` (reference primitives.py:42-107).  Entries are exact for any rank; the structure
    descriptor comes from ``structure.py`` for rank <= 1 and from ``structure_nd.py`` above."""
    if out.ndim > 1 or primal.ndim > 1:
        ip, iq = _bcast_index(primal, out.shape), _bcast_index(partial, out.shape)
        struct = nd.elementwise(primal.shape, out.shape, n_operands, partial.shape,
                                partial.is_literal or partial.size == 1)
        return struct, {(i, int(ip[i])): partial.nodes[int(iq[i])] for i in range(out.size)}
    n = out.size
    if primal.ndim == 0 and out.ndim == 0:
        return st.SCALAR, {(0, 0): partial.nodes[0]}
    if primal.ndim == 0:
        # rank-0 operand broadcast into a vector result
        return st.rank0_operand(n), {(i, 0): partial.nodes[i if partial.size == n else 0] for i in range(n)}
    if n_operands == 1:
        return st.diag(n), {(i, i): partial.nodes[i] for i in range(n)}
    if primal.shape != out.shape:
        # (1,) operand stretched to (n,): dense out dim, replicated primal dim
        return (st.bcast(n, primal.size),
                {(i, 0): partial.nodes[i if partial.size == n else 0] for i in range(n)})
    if partial.is_literal or partial.size == 1:
        return st.kron_scalar(n), {(i, i): partial.nodes[0] for i in range(n)}
    return st.diag(n), {(i, i): partial.nodes[i] for i in range(n)}


def lower_eqn(dag: Dag, eqn: Eqn, read, division: str = "ieee",
              _use_registry: bool = True) -> Tuple[Vec, List[Tuple[Var, st.Struct, Entries]]]:
    """Primal value and elemental edges of one equation.

    User-defined primitives and user-registered rules (``graphax_b200.primitives``: ``elemental_rules``,
    ``defelemental``, ``defelemental2`` - the reference's extension point, primitives.py:110-146) take precedence
    over the built-in rules below.

    Returns ``(out, [(operand_var, structure, entries), ...])`` with one edge per
    non-literal operand, or no edge at all when the rule yields fewer partials
    than traced operands (reference ``core.py:405-410``).

    ``division="ieee"`` lowers the reference's formulas verbatim: ``x/y``, ``1/y`` and
    ``-x/y**2`` are three correctly rounded divisions (primitives.py:197-199).
    ``division="reciprocal"`` evaluates one correctly rounded reciprocal ``r = 1/y``, the
    quotient ``q = x*r`` with one FMA correction step (``q + r*(x - y*q)``, correctly rounded
    up to rare ties) and the partials ``r`` and ``-(q*r)`` (likewise ``.5/out`` ->
    ``.5*(1/out)`` for sqrt and ``y*r, -x*r`` for atan2).  Same derivative, one IEEE
    reciprocal and a few FMAs instead of three IEEE divisions; the partial w.r.t. the
    denominator carries two roundings in both forms.
    """
    if _use_registry:
        from . import primitives as _registry
        entry = _registry.lookup(eqn.primitive)
        if entry is not None:
            return _registry.lower_custom(dag, eqn, read, division, entry,
                                          lambda: lower_eqn(dag, eqn, read, division, _use_registry=False))
    ew = _Ew(dag)
    recip = division == "reciprocal"
    if division not in ("ieee", "reciprocal"):
        raise ValueError("division must be 'ieee' or 'reciprocal'")
    prim = eqn.primitive
    ops = [read(a) for a in eqn.invars]
    traced = [i for i, a in enumerate(eqn.invars) if isinstance(a, Var)]

    def edges(out: Vec, partials: Sequence[Vec]):
        return out, [(eqn.invars[i],) + _parallel_jacobian(ops[i], out, partials[i], len(ops)) for i in traced]

    if prim in ("stop_gradient", "device_put"):    # primitives.py:452-462: value passes through, no edge
        return ops[0], []
    if prim == "iota":             # primitives.py:444-449: a constant array, no operands and no edge
        shape, dim = tuple(eqn.params["shape"]), int(eqn.params["dimension"])
        idx = np.indices(shape)[dim].reshape(-1) if shape else np.zeros(1, dtype=np.int64)
        return Vec(shape, [dag.const(float(i)) for i in idx]), []
    if prim == "pjit":
        # primitives.py:576-602: the reference evaluates the call and returns NO partials, so the call is one vertex
        # without in-edges (core.py:405-410) - derivatives do not flow through it.  Same here: the inner jaxpr is
        # evaluated for its value only.  (With real JAX, jnp.where / jnp.clip / jax.nn.* arrive as pjit equations.)
        inner = eqn.params["jaxpr"]
        warnings.warn(f"derivatives through the nested call {eqn.params.get('name') or 'pjit'!r} are dropped, as in the "
                      "reference (graphax primitives.py:576-602): it is a vertex without in-edges", stacklevel=2)
        env = {id(v): o for v, o in zip(inner.invars, ops)}

        def read_inner(a):
            return Vec((), [dag.const(a.val)], is_literal=True) if isinstance(a, Literal) else env[id(a)]
        for e in inner.eqns:
            env[id(e.outvar)], _ = lower_eqn(dag, e, read_inner, division)
        return read_inner(inner.outvars[0]), []

    # ---- elementwise binary (primitives.py:180-205, 237-239) ----------------
    if prim in ("max", "min"):     # primitives.py:208-215: ties send the derivative to the second operand
        x, y = ops
        out = ew.binary(prim, x, y)
        first = ew.binary("gt" if prim == "max" else "lt", x, y)          # 1.0 where x wins, else 0.0
        second = ew.sub(ew.lit(1.0), first)
        return edges(out, (first, second))
    if prim == "select_n":
        # lax.select_n(which, case0, case1) (jnp.where(c, x, y) = select_n(c, y, x)).  The reference's rule
        # (primitives.py:225-234) returns one zero Jacobian per case and none for the predicate, i.e. fewer partials
        # than traced operands, so it writes no edge at all (core.py:405-410) and the derivative through a `where`
        # is lost.  Here the selected case passes its derivative: masks `which` / `1 - which` (comparisons produce
        # exactly 1.0 / 0.0), and the predicate gets a zero partial.
        if len(ops) != 3:
            raise NotImplementedError("select_n with more than two cases")
        which, case0, case1 = ops
        shape = _bshape(_bshape(which.shape, case0.shape), case1.shape)
        iw, i0, i1 = (_bcast_index(v, shape) for v in (which, case0, case1))
        out = Vec(shape, [dag.emit("select", which.nodes[a], case1.nodes[c], case0.nodes[b]) for a, b, c in zip(iw, i0, i1)])
        on = Vec(shape, [dag.emit("select", which.nodes[a], dag.const(1.0), dag.const(0.0)) for a in iw])
        off = Vec(shape, [dag.emit("select", which.nodes[a], dag.const(0.0), dag.const(1.0)) for a in iw])
        zero = Vec(shape, [dag.const(0.0)] * len(iw))
        partials = (zero, off, on)
        return out, [(eqn.invars[i],) + _parallel_jacobian(ops[i], out, partials[i], 2) for i in traced]
    if prim == "convert_element_type":
        # primitives.py:1236-1255: a transform edge that converts the neighbouring edge's values - with one
        # floating-point type the identity, and like every transform edge it costs no multiplication
        x = ops[0]
        entries = {(i, i): dag.const(1.0) for i in range(x.size)}
        return x, ([(eqn.invars[0], nd.transform_edge(nd.NTransform("convert")), entries)] if traced else [])
    if prim in ("ge", "le", "ne"):   # complements of lt / gt / eq (a NaN operand compares as true here)
        x, y = ops
        base = ew.binary({"ge": "lt", "le": "gt", "ne": "eq"}[prim], x, y)
        out = Vec(base.shape, [dag.emit("select", n, dag.const(0.0), dag.const(1.0)) for n in base.nodes])
        zero = Vec(out.shape, [dag.const(0.0)] * out.size)
        return edges(out, (zero, zero))
    if prim in ("gt", "lt", "eq"):   # primitives.py:218-223: comparisons carry zero partials
        x, y = ops
        out = ew.binary(prim, x, y)
        zero = Vec(out.shape, [dag.const(0.0)] * out.size)
        return edges(out, (zero, zero))
    if prim in ("add", "sub", "mul", "div", "atan2", "pow"):
        x, y = ops
        if prim == "div" and recip:
            # r = RN(1/y); q0 = RN(x*r); q = RN(q0 + r*(x - y*q0)): one Newton-Markstein correction with
            # two FMAs brings the quotient back to the correctly rounded x/y (up to rare ties), so the
            # primal value agrees with the reference's true division
            r = ew.div(ew.lit(1.0), y)
            q0 = ew.mul(x, r)
            n = q0.size
            ix, iy, ir = _bcast_index(x, q0.shape), _bcast_index(y, q0.shape), _bcast_index(r, q0.shape)
            rem = Vec(q0.shape, [dag.fma(dag.neg(y.nodes[int(iy[i])]), q0.nodes[i], x.nodes[int(ix[i])])
                                 for i in range(n)])
            out = Vec(q0.shape, [dag.fma(rem.nodes[i], r.nodes[int(ir[i])], q0.nodes[i]) for i in range(n)])
        else:
            out = ew.binary(prim, x, y)
        if prim == "add":
            parts = (ew.ones_like(y), ew.ones_like(x))
        elif prim == "sub":
            parts = (ew.ones_like(y), ew.neg(ew.ones_like(x)))
        elif prim == "mul":
            parts = (y, x)
        elif prim == "div":
            if recip:
                parts = (r, ew.neg(ew.mul(out, r)))
            else:
                parts = (ew.div(ew.lit(1.0), y), ew.div(ew.neg(x), ew.ipow(y, 2)))
        elif prim == "atan2":
            abs2 = ew.add(ew.ipow(x, 2), ew.ipow(y, 2))
            if recip:
                r2 = ew.div(ew.lit(1.0), abs2)
                parts = (ew.mul(y, r2), ew.neg(ew.mul(x, r2)))
            else:
                parts = (ew.div(y, abs2), ew.div(ew.neg(x), abs2))
        else:  # pow
            parts = (ew.mul(y, ew.binary("pow", x, ew.sub(y, ew.lit(1.0)))), ew.mul(ew.unary("log", x), out))
        # ones_like / other-operand partials of a literal operand are Python floats in the reference
        parts = tuple(Vec(p.shape, p.nodes, is_literal=(prim in ("add", "sub", "mul") and ops[1 - i].is_literal))
                      for i, p in enumerate(parts))
        return edges(out, parts)

    # ---- elementwise unary (primitives.py:150-175) ---------------------------
    if prim == "integer_pow":
        (x,) = ops
        y = int(eqn.params["y"])
        out = ew.ipow(x, y)
        return edges(out, (ew.mul(ew.lit(float(y)), ew.ipow(x, y - 1)),))
    if prim == "square":
        (x,) = ops
        return edges(ew.mul(x, x), (ew.mul(ew.lit(2.0), x),))
    if prim in _UNARY:
        (x,) = ops
        out = ew.unary(prim, x)
        one = ew.lit(1.0)
        if prim == "neg":
            part = ew.neg(ew.ones_like(x))
        elif prim == "abs":
            part = ew.div(x, out)
        elif prim == "exp":
            part = out
        elif prim == "log":
            part = ew.div(one, x)
        elif prim == "sqrt":
            part = ew.mul(ew.lit(0.5), ew.div(one, out)) if recip else ew.div(ew.lit(0.5), out)
        elif prim == "logistic":
            part = ew.mul(out, ew.sub(one, out))
        elif prim == "log1p":
            part = ew.div(one, ew.add(one, x))
        elif prim == "sin":
            part = ew.unary("cos", x)
        elif prim == "asin":
            part = ew.div(one, ew.unary("sqrt", ew.sub(one, ew.ipow(x, 2))))
        elif prim == "cos":
            part = ew.neg(ew.unary("sin", x))
        elif prim == "acos":
            part = ew.div(ew.lit(-1.0), ew.unary("sqrt", ew.sub(one, ew.ipow(x, 2))))
        elif prim == "tan":
            part = ew.add(one, ew.ipow(out, 2))
        elif prim == "atan":
            part = ew.div(one, ew.add(one, ew.ipow(x, 2)))
        elif prim == "sinh":
            part = ew.unary("cosh", x)
        elif prim == "asinh":
            # the reference has sqrt(1+x^2) here (primitives.py:169), which is not the
            # derivative; this engine uses the correct 1/sqrt(1+x^2)
            part = ew.div(one, ew.unary("sqrt", ew.add(one, ew.ipow(x, 2))))
        elif prim == "cosh":
            part = ew.unary("sinh", x)
        elif prim == "acosh":
            part = ew.div(one, ew.unary("sqrt", ew.sub(ew.ipow(x, 2), one)))
        elif prim == "tanh":
            part = ew.sub(one, ew.ipow(out, 2))
        elif prim == "atanh":
            part = ew.div(one, ew.sub(one, ew.ipow(x, 2)))
        elif prim == "erf":
            two_over_sqrt_pi = ew.div(ew.lit(2.0), ew.unary("sqrt", ew.lit(math.pi)))
            part = ew.mul(ew.unary("exp", ew.neg(ew.ipow(x, 2))), two_over_sqrt_pi)
        else:  # pragma: no cover
            raise NotImplementedError(prim)
        return edges(out, (part,))

    # ---- structural rules ----------------------------------------------------
    if prim in ("reduce_sum", "dot_general", "slice", "concatenate", "broadcast_in_dim") and \
            (eqn.outvar.ndim > 1 or any(o.ndim > 1 for o in ops) or (prim == "broadcast_in_dim" and traced)
             or (prim == "dot_general" and eqn.params["dimension_numbers"][1] != ((), ()))):
        return lower_eqn(dag, Eqn(prim + "_nd", eqn.invars, eqn.outvar, eqn.params), read, division)
    if prim == "reduce_sum":          # primitives.py:243-272: row of ones over the reduced axis
        (x,) = ops
        if x.ndim == 0:
            return x, [(eqn.invars[0], st.SCALAR, {(0, 0): dag.const(1.0)})] if traced else []
        acc = x.nodes[0]
        for k in range(1, x.size):
            acc = dag.add(acc, x.nodes[k])
        out = Vec((), [acc])
        e = {(0, k): dag.const(1.0) for k in range(x.size)}
        return out, ([(eqn.invars[0], st.row(x.size), e)] if traced else [])
    if prim == "dot_general":         # primitives.py:349-441: each operand's edge carries the other operand
        x, y = ops
        if x.ndim != 1 or y.ndim != 1 or x.size != y.size:
            raise NotImplementedError("only 1-D . 1-D dot_general is lowered (scope row (f)1)")
        acc = dag.mul(x.nodes[0], y.nodes[0])
        for k in range(1, x.size):
            acc = dag.fma(x.nodes[k], y.nodes[k], acc)
        out = Vec((), [acc])
        es = []
        for i in traced:
            other = ops[1 - i]
            es.append((eqn.invars[i], st.row(x.size), {(0, k): other.nodes[k] for k in range(x.size)}))
        return out, es
    if prim == "slice":               # primitives.py:699-760: transform edge
        (x,) = ops
        (lo,), (hi,) = eqn.params["start_indices"], eqn.params["limit_indices"]
        out = Vec((hi - lo,), x.nodes[lo:hi])
        t = st.Transform("slice", lo, hi, x.size)
        e = {(i, lo + i): dag.const(1.0) for i in range(hi - lo)}
        return out, ([(eqn.invars[0], st.transform_edge(t), e)] if traced else [])
    if prim == "concatenate":         # primitives.py:971-1233: one transform edge per operand
        nodes, es, off = [], [], 0
        total = sum(o.size for o in ops)
        for a, o in zip(eqn.invars, ops):
            if o.ndim != 1:
                raise NotImplementedError("only 1-D concatenation is lowered")
            nodes.extend(o.nodes)
            if isinstance(a, Var):
                t = st.Transform("concat", off, off + o.size, total)
                es.append((a, st.transform_edge(t), {(off + i, i): dag.const(1.0) for i in range(o.size)}))
            off += o.size
        return Vec((total,), nodes), es
    if prim == "broadcast_in_dim":    # primitives.py:763-883; only the literal-operand form (jnp.zeros)
        (x,) = ops
        if traced:
            raise NotImplementedError("broadcast_in_dim of a traced operand is not lowered yet")
        shape = tuple(eqn.params["shape"])
        n = shape[0] if shape else 1
        return Vec(shape, [x.nodes[0]] * n), []
    # ---- shape ops and contractions of any rank: exact index maps, no tensor-level descriptor --------
    def gather_edge(x: Vec, index_map: np.ndarray, fill: Optional[int] = None):
        """out = x[index_map] (entries -1 read ``fill``): a selection Jacobian with structural ones."""
        flat = index_map.reshape(-1)
        nodes = [x.nodes[int(j)] if j >= 0 else fill for j in flat]
        entries = {(i, int(j)): dag.const(1.0) for i, j in enumerate(flat) if j >= 0}
        return Vec(tuple(index_map.shape), nodes), entries

    if prim in ("transpose", "reshape", "squeeze", "expand_dims"):
        (x,) = ops
        idx = x.index_array()
        if prim == "transpose":
            idx = np.transpose(idx, eqn.params["permutation"])
            t = nd.NTransform("transpose", perm=tuple(int(q) for q in eqn.params["permutation"]))
        else:
            idx = idx.reshape(eqn.outvar.shape)
            t = nd.NTransform("squeeze", dims=tuple(int(d) for d in eqn.params["dimensions"])) \
                if prim == "squeeze" and "dimensions" in eqn.params else \
                nd.NTransform("reshape", in_shape=x.shape, out_shape=tuple(eqn.outvar.shape))
        out, entries = gather_edge(x, idx)
        return out, ([(eqn.invars[0], nd.transform_edge(t), entries)] if traced else [])
    if prim == "slice_nd":
        (x,) = ops
        idx = x.index_array()[tuple(slice(a, b) for a, b in zip(eqn.params["start_indices"], eqn.params["limit_indices"]))]
        out, entries = gather_edge(x, idx)
        t = nd.NTransform("slice", in_shape=x.shape, out_shape=tuple(idx.shape))
        return out, ([(eqn.invars[0], nd.transform_edge(t), entries)] if traced else [])
    if prim == "concatenate_nd":
        axis, off, pieces, es = int(eqn.params["dimension"]), 0, [], []
        for o in ops:
            pieces.append(np.full(o.shape, -1, dtype=np.int64))
        full = np.concatenate(pieces, axis=axis)
        nodes: List[Optional[int]] = [None] * full.size
        out_idx = np.arange(full.size, dtype=np.int64).reshape(full.shape)
        for a, o in zip(eqn.invars, ops):
            sel = [slice(None)] * full.ndim
            sel[axis] = slice(off, off + o.shape[axis])
            dst = out_idx[tuple(sel)].reshape(-1)
            for i, j in zip(dst, range(o.size)):
                nodes[int(i)] = o.nodes[j]
            if isinstance(a, Var):
                t = nd.NTransform("concat", in_shape=o.shape, out_shape=tuple(full.shape), dims=(axis,),
                                  lo=off, hi=off + o.shape[axis])
                es.append((a, nd.transform_edge(t), {(int(i), j): dag.const(1.0) for i, j in zip(dst, range(o.size))}))
            off += o.shape[axis]
        return Vec(full.shape, nodes), es
    if prim == "broadcast_in_dim_nd":
        (x,) = ops
        shape, bdims = tuple(eqn.params["shape"]), tuple(eqn.params["broadcast_dimensions"])
        view = [1] * len(shape)
        for d, ax in zip(x.shape, bdims):
            view[ax] = d
        idx = np.broadcast_to(x.index_array().reshape(view) if x.ndim else np.zeros(view, dtype=np.int64), shape)
        out, entries = gather_edge(x, np.ascontiguousarray(idx))
        t = nd.NTransform("broadcast", in_shape=x.shape, out_shape=shape, dims=tuple(sorted(bdims)))
        return out, ([(eqn.invars[0], nd.transform_edge(t), entries)] if traced else [])
    if prim == "reduce_sum_nd":
        (x,) = ops
        axes = tuple(eqn.params["axes"])
        moved = np.moveaxis(x.index_array(), axes, range(-len(axes), 0)) if axes else x.index_array()
        groups = moved.reshape(int(np.prod(eqn.outvar.shape, dtype=np.int64)) if eqn.outvar.shape else 1, -1)
        nodes, entries = [], {}
        for i, grp in enumerate(groups):
            acc = x.nodes[int(grp[0])]
            for j in grp[1:]:
                acc = dag.add(acc, x.nodes[int(j)])
            nodes.append(acc)
            for j in grp:
                entries[(i, int(j))] = dag.const(1.0)
        return Vec(eqn.outvar.shape, nodes), ([(eqn.invars[0], nd.reduce_sum(x.shape, axes), entries)] if traced else [])
    if prim in ("reduce_max", "reduce_min"):      # primitives.py:275-346: where(x == out) / count
        (x,) = ops
        axes = tuple(eqn.params["axes"])
        moved = np.moveaxis(x.index_array(), axes, range(-len(axes), 0))
        groups = moved.reshape(int(np.prod(eqn.outvar.shape, dtype=np.int64)) if eqn.outvar.shape else 1, -1)
        nodes, entries = [], {}
        for i, grp in enumerate(groups):
            acc = x.nodes[int(grp[0])]
            for j in grp[1:]:
                acc = dag.emit("max" if prim == "reduce_max" else "min", acc, x.nodes[int(j)])
            nodes.append(acc)
            hits = [dag.emit("eq", x.nodes[int(j)], acc) for j in grp]
            count = hits[0]
            for h in hits[1:]:
                count = dag.add(count, h)
            for j, h in zip(grp, hits):
                entries[(i, int(j))] = dag.div(h, count)
        return Vec(eqn.outvar.shape, nodes), ([(eqn.invars[0], nd.reduce_extreme(x.shape, axes), entries)] if traced else [])
    if prim == "dot_general_nd":                  # primitives.py:349-441, any rank, batch dimensions allowed
        x, y = ops
        (cx, cy), (bx, by) = eqn.params["dimension_numbers"]
        fx = [d for d in range(x.ndim) if d not in cx and d not in bx]
        fy = [d for d in range(y.ndim) if d not in cy and d not in by]
        ix = np.transpose(x.index_array(), list(bx) + fx + list(cx)) if x.ndim else x.index_array()
        iy = np.transpose(y.index_array(), list(by) + fy + list(cy)) if y.ndim else y.index_array()
        nb = int(np.prod([x.shape[d] for d in bx], dtype=np.int64)) if bx else 1
        nk = int(np.prod([x.shape[d] for d in cx], dtype=np.int64)) if cx else 1
        ix, iy = ix.reshape(nb, -1, nk), iy.reshape(nb, -1, nk)
        nodes, ex, ey = [], {}, {}
        o = 0
        for b in range(nb):
            for i in range(ix.shape[1]):
                for j in range(iy.shape[1]):
                    acc = dag.mul(x.nodes[int(ix[b, i, 0])], y.nodes[int(iy[b, j, 0])])
                    for k in range(1, nk):
                        acc = dag.fma(x.nodes[int(ix[b, i, k])], y.nodes[int(iy[b, j, k])], acc)
                    nodes.append(acc)
                    for k in range(nk):
                        ex[(o, int(ix[b, i, k]))] = y.nodes[int(iy[b, j, k])]
                        ey[(o, int(iy[b, j, k]))] = x.nodes[int(ix[b, i, k])]
                    o += 1
        structs = nd.dot_general(x.shape, y.shape, eqn.params["dimension_numbers"])
        es = []
        for i in traced:
            es.append((eqn.invars[i], structs[i], ex if i == 0 else ey))
        return Vec(eqn.outvar.shape, nodes), es
    raise NotImplementedError(f"{prim} does not have registered elemental partial.")   # core.py:398
