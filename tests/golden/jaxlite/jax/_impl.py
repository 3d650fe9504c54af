"""jaxlite — a NumPy stand-in for the slice of the JAX API that graphax touches.

TEST INFRASTRUCTURE ONLY.  JAX is not installed in this image, so the reference
(`/root/reference/src/graphax`, pure Python on top of JAX) cannot be imported as
it is.  This package provides just enough of `jax`, `jax.numpy`, `jax.lax`,
`jax.tree_util`, `jax.core` / `jax._src.core` for the reference's UNMODIFIED
sources to run: `make_jaxpr` traces a Python function into equations with JAX's
primitive names, parameter names and equation order, and
every primitive evaluates eagerly with NumPy in float32.  It is used by
`tests/golden/make_reference_goldens.py` alone, to produce golden vectors from the
reference's own graph builder, elimination loop and SparseTensor algebra; nothing
under `graphax_b200/` may import it.

What it is not: XLA.  Each primitive is one NumPy call, so no fusion and no FMA
contraction; values agree with a real JAX run up to the backend's rounding of
individual operations (IEEE for + - * / sqrt, libm for transcendental functions).
"""
from __future__ import annotations

import builtins
import itertools
import math
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

np.seterr(all="ignore")

# ------------------------------------------------------------------------------------------------
# avals, variables, equations
# ------------------------------------------------------------------------------------------------


def _canon_dtype(dt) -> np.dtype:
    dt = np.dtype(dt)
    if dt == np.float64:
        return np.dtype(np.float32)
    if dt == np.int64:
        return np.dtype(np.int32)
    if dt == np.uint64:
        return np.dtype(np.uint32)
    return dt


class ShapedArray:
    def __init__(self, shape, dtype, weak_type: bool = False):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = _canon_dtype(dtype)
        self.weak_type = weak_type

    ndim = property(lambda self: len(self.shape))
    size = property(lambda self: int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1)

    def __repr__(self):
        return f"{self.dtype.name}[{','.join(map(str, self.shape))}]"

    def __eq__(self, other):
        return isinstance(other, ShapedArray) and (self.shape, self.dtype) == (other.shape, other.dtype)

    def __hash__(self):
        return hash((self.shape, self.dtype))


class Var:
    _ids = itertools.count()

    def __init__(self, aval: ShapedArray):
        self.aval = aval
        self.count = next(Var._ids)

    def __repr__(self):
        return f"v{self.count}:{self.aval}"


class Literal:
    def __init__(self, val, aval: ShapedArray):
        self.val = val
        self.aval = aval

    def __repr__(self):
        return f"{self.val}:{self.aval}"


class JaxprEqn:
    def __init__(self, invars, outvars, primitive, params):
        self.invars = list(invars)
        self.outvars = list(outvars)
        self.primitive = primitive
        self.params = dict(params)

    def __repr__(self):
        p = ",".join(f"{k}={v}" for k, v in self.params.items())
        return f"{self.outvars} = {self.primitive.name}[{p}] {self.invars}"


class Jaxpr:
    def __init__(self, constvars, invars, outvars, eqns):
        self.constvars = list(constvars)
        self.invars = self._invars = list(invars)
        self.outvars = self._outvars = list(outvars)
        self.eqns = list(eqns)

    def __repr__(self):
        return "{ lambda ; " + " ".join(map(repr, self.invars)) + " . let\n  " + \
            "\n  ".join(f"{i}: {e!r}" for i, e in enumerate(self.eqns, 1)) + f"\n  in {self.outvars} }}"


class ClosedJaxpr:
    def __init__(self, jaxpr: Jaxpr, consts: Sequence[Any]):
        self.jaxpr = jaxpr
        self.consts = list(consts)

    literals = property(lambda self: self.consts)
    eqns = property(lambda self: self.jaxpr.eqns)


# ------------------------------------------------------------------------------------------------
# values: concrete arrays and tracers share their operators
# ------------------------------------------------------------------------------------------------


class _Ops:
    """Operators of both concrete arrays and tracers, routed through the jnp functions below."""
    __array_ufunc__ = None          # NumPy scalars / arrays defer to our reflected operators
    __array_priority__ = 1000

    def __add__(s, o): return add(s, o)
    def __radd__(s, o): return add(o, s)
    def __sub__(s, o): return subtract(s, o)
    def __rsub__(s, o): return subtract(o, s)
    def __mul__(s, o): return multiply(s, o)
    def __rmul__(s, o): return multiply(o, s)
    def __truediv__(s, o): return divide(s, o)
    def __rtruediv__(s, o): return divide(o, s)
    def __pow__(s, o): return power(s, o)
    def __rpow__(s, o): return power(o, s)
    def __neg__(s): return negative(s)
    def __matmul__(s, o): return matmul(s, o)
    def __rmatmul__(s, o): return matmul(o, s)
    def __gt__(s, o): return greater(s, o)
    def __lt__(s, o): return less(s, o)
    def __eq__(s, o): return equal(s, o)
    __hash__ = object.__hash__
    def __getitem__(s, idx): return _getitem(s, idx)
    def reshape(s, *shape): return reshape(s, shape[0] if len(shape) == 1 and not isinstance(shape[0], (int, np.integer)) else shape)
    def sum(s, axis=None, keepdims=False): return sum(s, axis=axis, keepdims=keepdims)
    def astype(s, dtype): return convert_element_type(s, dtype)
    def mean(s, axis=None): return mean(s, axis)
    def max(s, axis=None, keepdims=False): return max(s, axis=axis, keepdims=keepdims)
    def min(s, axis=None, keepdims=False): return min(s, axis=axis, keepdims=keepdims)
    def squeeze(s, axis=None): return squeeze(s, axis)
    def transpose(s, *axes): return transpose(s, axes[0] if len(axes) == 1 and not isinstance(axes[0], int) else (axes or None))
    T = property(lambda s: transpose(s))
    def dot(s, o): return dot(s, o)

    shape = property(lambda s: s.aval.shape)
    ndim = property(lambda s: s.aval.ndim)
    size = property(lambda s: s.aval.size)
    dtype = property(lambda s: s.aval.dtype)

    def __len__(s):
        if not s.shape:
            raise TypeError("len() of unsized object")
        return s.shape[0]


class Array(_Ops):
    """A concrete value (what `jnp.array` returns); deliberately NOT an `np.ndarray` subclass, because the
    reference tells literals from traced operands with `isinstance(x, (float, np.ndarray, np.float32))`
    (`primitives.py:127`)."""

    def __init__(self, value):
        v = np.asarray(value)
        v = v.astype(_canon_dtype(v.dtype), copy=False)
        self._v = v if v.flags.c_contiguous else v.copy(order="C")
        self.aval = ShapedArray(self._v.shape, self._v.dtype)

    def __array__(self, dtype=None, copy=None):
        return self._v if dtype is None else self._v.astype(dtype)

    def __repr__(self):
        return f"Array({self._v!r})"

    def __float__(self): return float(self._v)
    def __int__(self): return int(self._v)
    def __bool__(self): return bool(self._v)
    def __iter__(self): return (Array(x) for x in self._v)
    def item(self): return self._v.item()
    def tolist(self): return self._v.tolist()


class Tracer(_Ops):
    def __init__(self, trace: "_Trace", var: Var):
        self._trace = trace
        self.var = var
        self.aval = var.aval

    def __repr__(self):
        return f"Traced<{self.aval}>"

    def __array__(self, *a, **k):
        raise TypeError("a traced value has no concrete contents")

    def __bool__(self):
        raise TypeError("truth value of a traced array")


class _Trace:
    def __init__(self):
        self.eqns: List[JaxprEqn] = []
        self.constvars: List[Var] = []
        self.consts: List[Any] = []
        self.const_of = {}


_TRACES: List[_Trace] = []


def _is_scalar_const(x) -> bool:
    return isinstance(x, (bool, int, float, np.generic)) or (isinstance(x, np.ndarray) and x.shape == ())


def _raw(x):
    """NumPy view of a concrete operand."""
    if isinstance(x, Array):
        return x._v
    if isinstance(x, Tracer):
        raise TypeError("traced value where a concrete one is required")
    return x


# ------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------


class Primitive:
    multiple_results = False

    def __init__(self, name: str, impl: Optional[Callable] = None):
        self.name = name
        self.impl = impl

    def __repr__(self):
        return self.name

    def def_impl(self, impl):
        self.impl = impl
        return impl

    def _eval(self, args, params):
        out = self.impl(*[_raw(a) for a in args], **params)
        return Array(out)

    def bind(self, *args, **params):
        if any(isinstance(a, BatchTracer) for a in args):
            return _batch_bind(self, args, params)
        if not _TRACES:
            return self._eval(args, params)
        # Inside make_jaxpr every bind is staged (JAX >= 0.4.36 has one current trace), constants included.
        trace = _TRACES[-1]
        invars, stand_ins = [], []
        for a in args:
            if isinstance(a, Tracer):
                invars.append(a.var)
                stand_ins.append(np.ones(a.aval.shape, a.aval.dtype))
            elif _is_scalar_const(a):
                v = np.asarray(a)
                v = v.astype(_canon_dtype(v.dtype))
                invars.append(Literal(v[()], ShapedArray((), v.dtype, weak_type=True)))
                stand_ins.append(v)
            else:   # a captured array constant becomes a constvar: one per object, as in JAX (constants are keyed by id)
                hit = trace.const_of.get(id(a))
                if hit is None:
                    c = a if isinstance(a, Array) else Array(a)
                    var = Var(c.aval)
                    trace.constvars.append(var)
                    trace.consts.append(c)
                    hit = trace.const_of[id(a)] = (a, c, var)     # `a` is kept alive so that its id stays unique
                _, c, var = hit
                invars.append(var)
                stand_ins.append(c._v)
        out = np.asarray(self.impl(*stand_ins, **params))       # abstract evaluation by example
        outvar = Var(ShapedArray(out.shape, out.dtype))
        trace.eqns.append(JaxprEqn(invars, [outvar], self, params))
        return Tracer(trace, outvar)


def _f(x):
    """float32 / int32 results like JAX with x64 disabled."""
    x = np.asarray(x)
    return x.astype(_canon_dtype(x.dtype), copy=False)


def _binary_impl(fn):
    def impl(x, y):
        x, y = np.asarray(x), np.asarray(y)
        if x.dtype != y.dtype:       # weak scalars and int/float mixes promote to the float32 side
            dt = _canon_dtype(np.result_type(x.dtype, y.dtype))
            x, y = x.astype(dt), y.astype(dt)
        return _f(fn(x, y))
    return impl


def _unary(name, fn):
    return Primitive(name, lambda x: _f(fn(np.asarray(x))))


neg_p = _unary("neg", np.negative)
abs_p = _unary("abs", np.abs)
sqrt_p = _unary("sqrt", np.sqrt)
exp_p = _unary("exp", np.exp)
log_p = _unary("log", np.log)
log1p_p = _unary("log1p", np.log1p)
sin_p = _unary("sin", np.sin)
cos_p = _unary("cos", np.cos)
tan_p = _unary("tan", np.tan)
asin_p = _unary("asin", np.arcsin)
acos_p = _unary("acos", np.arccos)
atan_p = _unary("atan", np.arctan)
sinh_p = _unary("sinh", np.sinh)
cosh_p = _unary("cosh", np.cosh)
tanh_p = _unary("tanh", np.tanh)
asinh_p = _unary("asinh", np.arcsinh)
acosh_p = _unary("acosh", np.arccosh)
atanh_p = _unary("atanh", np.arctanh)
square_p = _unary("square", lambda x: x * x)
logistic_p = _unary("logistic", lambda x: (1 / (1 + np.exp(-x))).astype(x.dtype))
stop_gradient_p = _unary("stop_gradient", lambda x: x)
device_put_p = _unary("device_put", lambda x: x)


def _erf(x):
    return np.vectorize(math.erf, otypes=[np.float64])(x).astype(x.dtype)


erf_p = _unary("erf", _erf)

add_p = Primitive("add", _binary_impl(np.add))
add_any_p = Primitive("add_any", _binary_impl(np.add))
sub_p = Primitive("sub", _binary_impl(np.subtract))
mul_p = Primitive("mul", _binary_impl(np.multiply))
div_p = Primitive("div", _binary_impl(np.divide))
atan2_p = Primitive("atan2", _binary_impl(np.arctan2))
pow_p = Primitive("pow", _binary_impl(np.power))
max_p = Primitive("max", _binary_impl(np.maximum))
min_p = Primitive("min", _binary_impl(np.minimum))
eq_p = Primitive("eq", lambda x, y: np.equal(x, y))
gt_p = Primitive("gt", lambda x, y: np.greater(x, y))
lt_p = Primitive("lt", lambda x, y: np.less(x, y))


def _integer_pow_impl(x, y):
    x = np.asarray(x)
    if y == 0:
        return np.ones_like(x)
    acc = None          # repeated squaring, as XLA expands integer_pow
    base, n = x, builtins.abs(int(y))
    while n:
        if n & 1:
            acc = base if acc is None else acc * base
        n >>= 1
        if n:
            base = base * base
    return _f(acc if y > 0 else 1 / acc)


integer_pow_p = Primitive("integer_pow", _integer_pow_impl)
reduce_sum_p = Primitive("reduce_sum", lambda x, axes, **_: _f(np.sum(np.asarray(x), axis=tuple(axes), dtype=np.asarray(x).dtype)))
reduce_max_p = Primitive("reduce_max", lambda x, axes, **_: _f(np.max(np.asarray(x), axis=tuple(axes))))
reduce_min_p = Primitive("reduce_min", lambda x, axes, **_: _f(np.min(np.asarray(x), axis=tuple(axes))))
transpose_p = Primitive("transpose", lambda x, permutation: np.transpose(np.asarray(x), permutation))
reshape_p = Primitive("reshape", lambda x, new_sizes, dimensions=None, **_: np.reshape(np.asarray(x), new_sizes))
squeeze_p = Primitive("squeeze", lambda x, dimensions: np.squeeze(np.asarray(x), axis=tuple(dimensions)))
convert_element_type_p = Primitive("convert_element_type",
                                   lambda x, new_dtype, weak_type=False, **_: np.asarray(x).astype(_canon_dtype(new_dtype)))
iota_p = Primitive("iota", lambda dtype, shape, dimension, **_: np.broadcast_to(
    np.arange(shape[dimension], dtype=_canon_dtype(dtype)).reshape([-1 if i == dimension else 1 for i in range(len(shape))]), shape))
pjit_p = Primitive("pjit")
pjit_p.multiple_results = True


def _slice_impl(x, start_indices, limit_indices, strides=None):
    x = np.asarray(x)
    strides = strides or (1,) * x.ndim
    return x[tuple(slice(int(a), int(b), int(s)) for a, b, s in zip(start_indices, limit_indices, strides))]


slice_p = Primitive("slice", _slice_impl)
concatenate_p = Primitive("concatenate", lambda *xs, dimension: _f(np.concatenate([np.asarray(x) for x in xs], axis=dimension)))


def _broadcast_in_dim_impl(x, shape, broadcast_dimensions, sharding=None):
    x = np.asarray(x)
    inter = [1] * len(shape)
    for src, dst in enumerate(broadcast_dimensions):
        inter[dst] = x.shape[src]
    return np.broadcast_to(x.reshape(inter), tuple(shape)).copy()


broadcast_in_dim_p = Primitive("broadcast_in_dim", _broadcast_in_dim_impl)


def _dot_general_impl(lhs, rhs, dimension_numbers, precision=None, preferred_element_type=None, **_):
    lhs, rhs = np.asarray(lhs), np.asarray(rhs)
    (lc, rc), (lb, rb) = dimension_numbers
    letters = iter("abcdefghijklmnopqrstuvwxyz")
    ls, rs = [None] * lhs.ndim, [None] * rhs.ndim
    for i, j in zip(lb, rb):
        ls[i] = rs[j] = next(letters)
    for i, j in zip(lc, rc):
        ls[i] = rs[j] = next(letters)
    for s in (ls, rs):
        for i in range(len(s)):
            if s[i] is None:
                s[i] = next(letters)
    out = [ls[i] for i in lb] + [c for i, c in enumerate(ls) if i not in lb and i not in lc] + \
          [c for i, c in enumerate(rs) if i not in rb and i not in rc]
    dt = _canon_dtype(np.result_type(lhs.dtype, rhs.dtype))
    return np.einsum(f"{''.join(ls)},{''.join(rs)}->{''.join(out)}", lhs.astype(dt), rhs.astype(dt)).astype(dt)


dot_general_p = Primitive("dot_general", _dot_general_impl)


def _select_n_impl(which, *cases):
    which = np.asarray(which)
    return np.choose(which.astype(np.int64), [np.asarray(c) for c in cases])


select_n_p = Primitive("select_n", _select_n_impl)

# ------------------------------------------------------------------------------------------------
# jax.numpy / jax.lax functions (the lowering rules of SURVEY.md Appendix A.1)
# ------------------------------------------------------------------------------------------------


def _is_value(x) -> bool:
    return isinstance(x, (Array, Tracer, BatchTracer))


def _shape_of(x) -> Tuple[int, ...]:
    return tuple(x.shape) if _is_value(x) or isinstance(x, np.ndarray) else np.shape(x)


def _lift(x):
    """Python lists / NumPy arrays entering a jnp function become concrete Arrays; scalars stay weak."""
    if _is_value(x) or _is_scalar_const(x):
        return x
    return Array(x)


def _same(x):
    return x if _is_value(x) or _TRACES else Array(x)


def _promote_ranks(x, y):
    """jnp's `_promote_shapes`: operands of different non-zero rank get leading unit axes through
    `lax.broadcast_to_rank` (one `broadcast_in_dim`); rank-0 operands rely on lax's scalar broadcasting."""
    sx, sy = _shape_of(x), _shape_of(y)
    if sx and sy and len(sx) != len(sy):
        if len(sx) < len(sy):
            x = _expand_leading(x, len(sy) - len(sx))
        else:
            y = _expand_leading(y, len(sx) - len(sy))
    return x, y


def _expand_leading(x, n):
    shape = (1,) * n + _shape_of(x)
    return broadcast_in_dim_p.bind(x, shape=shape, broadcast_dimensions=tuple(range(n, len(shape))), sharding=None)


def _weak(c, other):
    """A Python scalar next to an array takes the array's floating dtype (weak typing): `2*x` -> `mul 2.0:f32 x`."""
    dt = getattr(other, "dtype", np.dtype(np.float32))
    if np.issubdtype(dt, np.floating) or isinstance(c, (float, np.floating)):
        dt = np.dtype(np.float32)
    return np.asarray(c).astype(dt)[()]


def _binary(prim):
    def fn(x, y):
        x, y = _lift(x), _lift(y)
        if _is_scalar_const(x) and not _is_scalar_const(y):
            x = _weak(x, y)
        elif _is_scalar_const(y) and not _is_scalar_const(x):
            y = _weak(y, x)
        if _is_value(x) and (_is_value(y) or _is_scalar_const(y)) or _is_value(y) and _is_scalar_const(x):
            dx, dy = np.asarray(x).dtype if _is_scalar_const(x) else x.dtype, np.asarray(y).dtype if _is_scalar_const(y) else y.dtype
            if prim.name not in ("eq", "gt", "lt") or dx != dy:
                to = _canon_dtype(np.result_type(dx, dy))
                if _is_value(x) and x.dtype != to:
                    x = convert_element_type(x, to)
                if _is_value(y) and y.dtype != to:
                    y = convert_element_type(y, to)
        x, y = _promote_ranks(x, y)
        return prim.bind(x, y)
    fn.__name__ = prim.name
    return fn


add = _binary(add_p)
subtract = _binary(sub_p)
multiply = _binary(mul_p)
divide = true_divide = _binary(div_p)
arctan2 = _binary(atan2_p)
maximum = _binary(max_p)
minimum = _binary(min_p)
greater = _binary(gt_p)
less = _binary(lt_p)
equal = _binary(eq_p)


def power(x, y):
    if isinstance(y, (int, np.integer)) and not isinstance(y, bool):
        return integer_pow_p.bind(_lift(x), y=int(y))       # jnp.power with a static integer exponent
    return _binary(pow_p)(x, y)


def _unary_fn(prim):
    def fn(x):
        return prim.bind(_lift(x))
    fn.__name__ = prim.name
    return fn


negative = _unary_fn(neg_p)
abs = absolute = _unary_fn(abs_p)
sqrt = _unary_fn(sqrt_p)
exp = _unary_fn(exp_p)
log = _unary_fn(log_p)
log1p = _unary_fn(log1p_p)
sin = _unary_fn(sin_p)
cos = _unary_fn(cos_p)
tan = _unary_fn(tan_p)
arcsin = _unary_fn(asin_p)
arccos = _unary_fn(acos_p)
arctan = _unary_fn(atan_p)
sinh = _unary_fn(sinh_p)
cosh = _unary_fn(cosh_p)
tanh = _unary_fn(tanh_p)
arcsinh = _unary_fn(asinh_p)
arccosh = _unary_fn(acosh_p)
arctanh = _unary_fn(atanh_p)
square = _unary_fn(square_p)
logistic = _unary_fn(logistic_p)
erf = _unary_fn(erf_p)
stop_gradient = _unary_fn(stop_gradient_p)


def _axes(axis, ndim) -> Tuple[int, ...]:
    if axis is None:
        return tuple(range(ndim))
    if isinstance(axis, (int, np.integer)):
        axis = (axis,)
    return tuple(sorted(int(a) % builtins.max(ndim, 1) for a in axis))


def _reduce(prim):
    def fn(x, axis=None, keepdims=False, dtype=None):
        x = _lift(x)
        axes = _axes(axis, x.ndim)
        out = prim.bind(x, axes=axes)
        if keepdims:
            shape = tuple(1 if i in axes else s for i, s in enumerate(x.shape))
            out = broadcast_in_dim_p.bind(out, shape=shape, sharding=None,
                                          broadcast_dimensions=tuple(i for i in range(x.ndim) if i not in axes))
        return out
    return fn


sum = _reduce(reduce_sum_p)
max = amax = _reduce(reduce_max_p)
min = amin = _reduce(reduce_min_p)


def mean(x, axis=None):
    x = _lift(x)
    n = int(np.prod([x.shape[a] for a in _axes(axis, x.ndim)], dtype=np.int64))
    return divide(sum(x, axis=axis), float(n))


def dot_general(lhs, rhs, dimension_numbers, precision=None, preferred_element_type=None):
    (lc, rc), (lb, rb) = dimension_numbers
    dn = ((tuple(int(i) for i in lc), tuple(int(i) for i in rc)), (tuple(int(i) for i in lb), tuple(int(i) for i in rb)))
    return dot_general_p.bind(_lift(lhs), _lift(rhs), dimension_numbers=dn, precision=precision,
                              preferred_element_type=np.dtype(np.float32))


def dot(a, b):
    a, b = _lift(a), _lift(b)
    if a.ndim == 0 or b.ndim == 0:
        return multiply(a, b)
    return dot_general(a, b, (((a.ndim - 1,), (builtins.max(b.ndim - 2, 0),)), ((), ())))


def matmul(a, b):
    a, b = _lift(a), _lift(b)
    if not (1 <= a.ndim <= 2 and 1 <= b.ndim <= 2):
        raise NotImplementedError("jaxlite: matmul of ranks 1 and 2 only")
    return dot_general(a, b, (((a.ndim - 1,), (0,)), ((), ())))


def transpose(x, axes=None):
    x = _lift(x)
    perm = tuple(range(x.ndim))[::-1] if axes is None else tuple(int(a) for a in axes)
    if perm == tuple(range(x.ndim)):
        return _same(x)                 # lax.transpose returns its operand for the identity permutation
    return transpose_p.bind(x, permutation=perm)


def reshape(x, shape):
    x = _lift(x)
    shape = (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)
    if -1 in shape:
        known = int(np.prod([s for s in shape if s != -1], dtype=np.int64))
        shape = tuple(x.size // known if s == -1 else s for s in shape)
    if shape == tuple(x.shape):
        return _same(x)                 # lax.reshape returns its operand when nothing changes
    return reshape_p.bind(x, new_sizes=shape, dimensions=None, sharding=None)


def squeeze(x, axis=None):
    x = _lift(x)
    dims = tuple(i for i, s in enumerate(x.shape) if s == 1) if axis is None else _axes(axis, x.ndim)
    if not dims:
        return _same(x)
    return squeeze_p.bind(x, dimensions=dims)


def expand_dims(x, axis):
    x = _lift(x)
    axes = (axis,) if isinstance(axis, (int, np.integer)) else tuple(axis)
    nd = x.ndim + len(axes)
    axes = tuple(sorted(a % nd for a in axes))
    shape, it = [], iter(x.shape)
    for i in range(nd):
        shape.append(1 if i in axes else next(it))
    return broadcast_in_dim_p.bind(x, shape=tuple(shape), sharding=None,
                                   broadcast_dimensions=tuple(i for i in range(nd) if i not in axes))


def broadcast_in_dim(x, shape, broadcast_dimensions):
    x = _lift(x)
    shape = tuple(int(s) for s in shape)
    if _is_value(x) and tuple(x.shape) == shape and tuple(broadcast_dimensions) == tuple(range(len(shape))):
        return x
    return broadcast_in_dim_p.bind(x, shape=shape, broadcast_dimensions=tuple(broadcast_dimensions), sharding=None)


def convert_element_type(x, new_dtype):
    x = _lift(x)
    new_dtype = _canon_dtype(new_dtype)
    if _is_value(x) and x.dtype == new_dtype:
        return x
    return convert_element_type_p.bind(x, new_dtype=new_dtype, weak_type=False, sharding=None)


def lax_slice(x, start_indices, limit_indices, strides=None):
    return slice_p.bind(_lift(x), start_indices=tuple(int(i) for i in start_indices),
                        limit_indices=tuple(int(i) for i in limit_indices),
                        strides=None if strides is None else tuple(strides))


def slice_in_dim(x, start_index, limit_index, stride=1, axis=0):
    x = _lift(x)
    axis = axis % x.ndim
    start = [0] * x.ndim
    limit = list(x.shape)
    start[axis] = 0 if start_index is None else int(start_index)
    limit[axis] = x.shape[axis] if limit_index is None else int(limit_index)
    return lax_slice(x, start, limit)


def _getitem(x, idx):
    """Static basic indexing.  Slices with unit stride lower to one `slice` equation (jnp's
    `_attempt_rewriting_take_via_slice`); an integer index adds a `squeeze`."""
    idx = idx if isinstance(idx, tuple) else (idx,)
    if any(not isinstance(i, (slice, int, np.integer)) for i in idx):
        if _TRACES:
            raise NotImplementedError("jaxlite: only static basic indexing is traced")
        return Array(_raw(x)[tuple(_raw(i) for i in idx)])
    idx = idx + (slice(None),) * (x.ndim - len(idx))
    starts, limits, drop = [], [], []
    for ax, (i, n) in enumerate(zip(idx, x.shape)):
        if isinstance(i, slice):
            a, b, step = i.indices(n)
            if step != 1:
                raise NotImplementedError("jaxlite: strided slices")
            starts.append(a)
            limits.append(builtins.max(a, b))
        else:
            i = int(i) + (n if int(i) < 0 else 0)
            starts.append(i)
            limits.append(i + 1)
            drop.append(ax)
    out = x
    if any(a != 0 or b != n for a, b, n in zip(starts, limits, x.shape)):
        out = lax_slice(x, starts, limits)
    if drop:
        out = squeeze_p.bind(out, dimensions=tuple(drop))
    return out


def concatenate(parts, axis=0):
    parts = [_lift(p) for p in parts]
    axis = axis % builtins.max(parts[0].ndim, 1)
    if len(parts) == 1:
        return parts[0]
    return concatenate_p.bind(*parts, dimension=axis)


def full(shape, fill_value, dtype=None):
    shape = (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)
    dtype = _canon_dtype(dtype or np.float32)
    fill = np.asarray(fill_value, dtype=dtype)[()]
    if shape == ():
        return Array(fill) if not _TRACES else fill          # a rank-0 constant needs no broadcast
    if _TRACES:                                              # jnp.zeros(n) inside a traced function
        return broadcast_in_dim_p.bind(fill, shape=shape, broadcast_dimensions=(), sharding=None)
    return Array(np.full(shape, fill, dtype))


def zeros(shape, dtype=None): return full(shape, 0, dtype)
def ones(shape, dtype=None): return full(shape, 1, dtype)
def zeros_like(x, dtype=None): return full(_shape_of(x), 0, dtype or getattr(x, "dtype", np.float32))
def ones_like(x, dtype=None): return full(_shape_of(x), 1, dtype or getattr(x, "dtype", np.float32))


def split(x, indices_or_sections, axis=0):
    x = _lift(x)
    axis = axis % x.ndim
    n = x.shape[axis]
    if isinstance(indices_or_sections, (int, np.integer)):
        step = n // int(indices_or_sections)
        cuts = list(range(step, n, step))
    else:
        cuts = [int(i) for i in indices_or_sections]
    parts, lo = [], 0
    for hi in cuts + [n]:
        parts.append(slice_in_dim(x, lo, hi, axis=axis))
        lo = hi
    return parts


# ---- functions the reference only ever calls on concrete values ---------------------------------


def _concrete(fn):
    def wrapped(*args, **kwargs):
        if any(isinstance(a, (Tracer, BatchTracer)) for a in args):
            raise NotImplementedError(f"jaxlite: {fn.__name__} is not traceable")
        return fn(*args, **kwargs)      # inside a trace the result is a constant (JAX would stage iota / eq / ...)
    wrapped.__name__ = fn.__name__
    return wrapped


@_concrete
def array(x, dtype=None):
    if isinstance(x, (list, tuple)):
        x = [_raw(e) for e in x]
    v = np.asarray(_raw(x))
    return Array(v.astype(_canon_dtype(dtype)) if dtype is not None else v)


asarray = array


@_concrete
def eye(n, m=None, dtype=None):
    return Array(np.eye(int(n), None if m is None else int(m), dtype=_canon_dtype(dtype or np.float32)))


def tile(x, reps):
    reps = (reps,) if isinstance(reps, (int, np.integer)) else tuple(int(r) for r in reps)
    if not isinstance(x, (Tracer, BatchTracer)):
        return Array(np.tile(_raw(x), reps))
    reps = (1,) * (x.ndim - len(reps)) + reps
    if builtins.all(r == 1 for r in reps):
        return x if len(reps) == x.ndim else reshape(x, (1,) * (len(reps) - x.ndim) + tuple(x.shape))
    shape = (1,) * (len(reps) - x.ndim) + tuple(x.shape)
    inter = tuple(v for r, d in zip(reps, shape) for v in (r, d))
    y = broadcast_in_dim(reshape(x, shape), inter, tuple(range(1, 2 * len(shape), 2)))
    return reshape(y, tuple(r * d for r, d in zip(reps, shape)))


@_concrete
def diagonal(x, offset=0, axis1=0, axis2=1):
    return Array(np.diagonal(_raw(x), offset, axis1, axis2).copy())


def where(c, x, y):
    """`jnp.where(c, x, y)` = `lax.select(c, x, y)` = `select_n(c, y, x)`; eager calls keep jnp's weak int32 result
    for `where(cond, 1, 0)`."""
    if not _TRACES and not any(isinstance(v, BatchTracer) for v in (c, x, y)):
        return Array(_f(np.where(_raw(c), _raw(x), _raw(y))))
    c, x, y = _lift(c), _lift(x), _lift(y)
    if _is_scalar_const(x):
        x = _weak(x, y)
    if _is_scalar_const(y):
        y = _weak(y, x)
    shape = np.broadcast_shapes(_shape_of(c), _shape_of(x), _shape_of(y))
    full = lambda v: v if _shape_of(v) == shape or not _shape_of(v) and _is_scalar_const(v) and False else \
        broadcast_in_dim(v, shape, tuple(range(len(shape) - len(_shape_of(v)), len(shape))))
    return select_n_p.bind(full(c), full(y), full(x))


@_concrete
def einsum(subscripts, *operands, **kwargs):
    return Array(_f(np.einsum(subscripts, *[np.asarray(_raw(o)) for o in operands])))


@_concrete
def diag(v, k=0):
    return Array(np.diag(_raw(v), k))


@_concrete
def all(a, axis=None):      # noqa: A001
    return Array(np.all(_raw(a), axis=axis))


@_concrete
def allclose(a, b, rtol=1e-5, atol=1e-8, equal_nan=False):
    return Array(np.allclose(np.asarray(_raw(a), np.float64), np.asarray(_raw(b), np.float64), rtol=rtol, atol=atol, equal_nan=equal_nan))


@_concrete
def logical_and(a, b):
    return Array(np.logical_and(_raw(a), _raw(b)))


class ScatterDimensionNumbers:
    def __init__(self, update_window_dims, inserted_window_dims, scatter_dims_to_operand_dims,
                 operand_batching_dims=(), scatter_indices_batching_dims=()):
        self.update_window_dims = tuple(update_window_dims)
        self.inserted_window_dims = tuple(inserted_window_dims)
        self.scatter_dims_to_operand_dims = tuple(scatter_dims_to_operand_dims)


def _scatter_impl(operand, scatter_indices, updates, dimension_numbers, **_):
    op, idx, upd = np.array(operand), np.asarray(scatter_indices), np.asarray(updates)
    dn = dimension_numbers
    if idx.ndim != 1 or dn.inserted_window_dims or len(dn.update_window_dims) != upd.ndim:
        raise NotImplementedError("jaxlite: general scatter")
    if upd.ndim != op.ndim:
        raise NotImplementedError("jaxlite: scatter with collapsed dimensions")
    start = [0] * op.ndim
    for k, d in enumerate(dn.scatter_dims_to_operand_dims):
        start[d] = int(idx[k])
    if any(s + n > m for s, n, m in zip(start, upd.shape, op.shape)):
        return op                                          # out-of-bounds windows are dropped
    op[tuple(slice(s, s + n) for s, n in zip(start, upd.shape))] = upd.astype(op.dtype)
    return op


scatter_p = Primitive("scatter", _scatter_impl)


def scatter(operand, scatter_indices, updates, dimension_numbers, indices_are_sorted=False, unique_indices=False,
            mode=None):
    """`lax.scatter` for the one form the reference uses (`primitives.py:744-749, 1062-1067, 1114-1117`):
    a single index vector and no inserted window dimensions, i.e. the update block replaces the window of
    `operand` that starts at the given offsets."""
    return scatter_p.bind(_lift(operand), _lift(scatter_indices), _lift(updates), dimension_numbers=dimension_numbers)


# ------------------------------------------------------------------------------------------------
# make_jaxpr, pytrees
# ------------------------------------------------------------------------------------------------


class ShapeDtypeStruct:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), _canon_dtype(dtype)


def make_jaxpr(fun: Callable) -> Callable:
    def traced(*args, **kwargs):
        flat, _ = tree_flatten(args)
        trace = _Trace()
        invars = []
        for a in flat:
            if isinstance(a, (ShapeDtypeStruct, Tracer, BatchTracer)):   # tracers: make_jaxpr inside a trace / vmap
                invars.append(Var(ShapedArray(a.shape, a.dtype)))
                continue
            v = np.asarray(_raw(a))
            invars.append(Var(ShapedArray(v.shape, v.dtype)))
        tracers = [Tracer(trace, v) for v in invars]
        _, treedef = tree_flatten(args)
        _TRACES.append(trace)
        try:
            outs = fun(*tree_unflatten(treedef, tracers), **kwargs)
        finally:
            _TRACES.pop()
        flat_out, _ = tree_flatten(outs)
        outvars = []
        for o in flat_out:
            if isinstance(o, Tracer):
                outvars.append(o.var)
            else:
                v = np.asarray(_raw(o))
                outvars.append(Literal(v[()] if v.shape == () else v, ShapedArray(v.shape, v.dtype)))
        jaxpr = Jaxpr(trace.constvars, invars, outvars, trace.eqns)
        return ClosedJaxpr(jaxpr, trace.consts)
    return traced


class PyTreeDef:
    def __init__(self, kind, meta, children):
        self.kind, self.meta, self.children = kind, meta, children

    @property
    def num_leaves(self):
        return 1 if self.kind == "leaf" else builtins.sum(c.num_leaves for c in self.children)

    def __repr__(self):
        return f"PyTreeDef({self.kind}, {self.children})"

    def __eq__(self, other):
        return isinstance(other, PyTreeDef) and (self.kind, self.meta, self.children) == (other.kind, other.meta, other.children)


_LEAF = PyTreeDef("leaf", None, [])
_NODE_CLASSES: Dict[type, Tuple[Callable, Callable]] = {}


def register_pytree_node_class(cls):
    _NODE_CLASSES[cls] = (lambda x: x.tree_flatten(), cls.tree_unflatten)
    return cls


def tree_flatten(tree, is_leaf=None):
    leaves: List[Any] = []

    def go(t):
        if is_leaf is not None and is_leaf(t):
            leaves.append(t)
            return _LEAF
        if t is None:
            return PyTreeDef("none", None, [])
        if isinstance(t, tuple):
            return PyTreeDef("tuple", None, [go(c) for c in t])
        if isinstance(t, list):
            return PyTreeDef("list", None, [go(c) for c in t])
        if isinstance(t, dict):
            keys = sorted(t)
            return PyTreeDef("dict", tuple(keys), [go(t[k]) for k in keys])
        if type(t) in _NODE_CLASSES:
            children, aux = _NODE_CLASSES[type(t)][0](t)
            return PyTreeDef(("custom", type(t)), aux, [go(c) for c in children])
        leaves.append(t)
        return _LEAF
    return leaves, go(tree)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def go(d):
        if d.kind == "leaf":
            return next(it)
        if d.kind == "none":
            return None
        kids = [go(c) for c in d.children]
        if d.kind == "tuple":
            return tuple(kids)
        if d.kind == "list":
            return kids
        if d.kind == "dict":
            return dict(zip(d.meta, kids))
        return _NODE_CLASSES[d.kind[1]][1](d.meta, kids)
    return go(treedef)


def tree_structure(tree):
    return tree_flatten(tree)[1]


def tree_leaves(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[0]


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, treedef = tree_flatten(tree, is_leaf)
    others = [tree_flatten(r, is_leaf)[0] for r in rest]
    return tree_unflatten(treedef, [f(*xs) for xs in zip(leaves, *others)])


def tree_reduce(f, tree, *init):
    import functools
    return functools.reduce(f, tree_leaves(tree), *init)


def safe_map(f, *args):
    args = [list(a) for a in args]
    n = len(args[0])
    for a in args[1:]:
        assert len(a) == n, f"length mismatch: {[len(a) for a in args]}"
    return list(map(f, *args))


def jit(fun=None, **kwargs):
    """No compilation here: `jit(f)` only turns Python / NumPy arguments into arrays, as tracing would."""
    if fun is None:
        return lambda f: jit(f)

    def wrapped(*args, **kw):
        return fun(*tree_map(lambda a: a if _is_value(a) else Array(a), args), **kw)
    wrapped.__wrapped__ = fun
    return wrapped


def _unavailable(name):
    def fn(*a, **k):
        raise NotImplementedError(f"jaxlite does not provide jax.{name}")
    return fn



# ------------------------------------------------------------------------------------------------
# vmap: the batching rules of the primitives the reference's vmapped tests reach (batch axis kept in front)
# ------------------------------------------------------------------------------------------------


class BatchTracer(_Ops):
    """A value with a hidden leading batch axis; `aval` is the per-sample view, as inside `jax.vmap`."""

    def __init__(self, val):
        self.val = val
        self.aval = ShapedArray(val.shape[1:], val.dtype)

    def __repr__(self):
        return f"Batched<{self.aval}>"


def _ndim(x) -> int:
    return len(_shape_of(x))


def _broadcast_batcher(prim, args, params):
    """jax.interpreters.batching.broadcast_batcher: equal shapes (and scalars) bind directly; otherwise unbatched
    arrays get a unit batch axis and lower-rank operands trailing unit axes (`lax.expand_dims`), the primitive's
    own broadcasting does the rest."""
    vals = [a.val if isinstance(a, BatchTracer) else a for a in args]
    shapes = {_shape_of(v) for v in vals if _ndim(v)}
    if len(shapes) > 1 or any(_ndim(a) and not isinstance(a, BatchTracer) for a in args):
        vals = [v if isinstance(a, BatchTracer) or not _ndim(v) else _expand_leading(v, 1) for a, v in zip(args, vals)]
        nd = builtins.max(_ndim(v) for v in vals)
        vals = [v if not isinstance(a, BatchTracer) or _ndim(v) == nd else expand_dims(v, tuple(range(_ndim(v), nd)))
                for a, v in zip(args, vals)]
    return BatchTracer(prim.bind(*vals, **params))


def _batch_bind(prim, args, params):
    name = prim.name
    B = next(a.val.shape[0] for a in args if isinstance(a, BatchTracer))
    if len(args) == 1 and name not in _BATCH_RULES:
        return BatchTracer(prim.bind(args[0].val, **params))         # elementwise unary
    if name in ("add", "add_any", "sub", "mul", "div", "atan2", "pow", "max", "min", "eq", "gt", "lt"):
        return _broadcast_batcher(prim, args, params)
    if name in _BATCH_RULES:
        return _BATCH_RULES[name](prim, args, params, B)
    raise NotImplementedError(f"jaxlite: no batching rule for {name}")


def _b_reduce(prim, args, params, B):
    return BatchTracer(prim.bind(args[0].val, axes=tuple(a + 1 for a in params["axes"])))


def _b_dot_general(prim, args, params, B):
    lhs, rhs = args
    (lc, rc), (lb, rb) = params["dimension_numbers"]
    if not (isinstance(lhs, BatchTracer) and isinstance(rhs, BatchTracer)):
        if isinstance(lhs, BatchTracer) and not lb:      # the batch axis is the first free axis of lhs
            dn = ((tuple(i + 1 for i in lc), rc), ((), ()))
            return BatchTracer(prim.bind(lhs.val, rhs, **dict(params, dimension_numbers=dn)))
        if isinstance(rhs, BatchTracer) and not lb:      # the batch axis is the first free axis of rhs: move it to the front
            dn = ((lc, tuple(i + 1 for i in rc)), ((), ()))
            out = prim.bind(lhs, rhs.val, **dict(params, dimension_numbers=dn))
            pos = _ndim(lhs) - len(lc)
            perm = (pos,) + tuple(i for i in range(_ndim(out)) if i != pos)
            return BatchTracer(out if pos == 0 else transpose_p.bind(out, permutation=perm))
        raise NotImplementedError("jaxlite: dot_general with batch dimensions and one batched operand")
    sh = lambda t: tuple(i + 1 for i in t)
    dn = ((sh(lc), sh(rc)), ((0,) + sh(lb), (0,) + sh(rb)))
    return BatchTracer(prim.bind(lhs.val, rhs.val, **dict(params, dimension_numbers=dn)))


def _b_slice(prim, args, params, B):
    strides = params.get("strides")
    return BatchTracer(prim.bind(args[0].val, start_indices=(0,) + tuple(params["start_indices"]),
                                 limit_indices=(B,) + tuple(params["limit_indices"]),
                                 strides=None if strides is None else (1,) + tuple(strides)))


def _b_concatenate(prim, args, params, B):
    vals = []
    for a in args:
        if isinstance(a, BatchTracer):
            vals.append(a.val)
        else:       # lax.broadcast(op, (size,))
            shape = (B,) + _shape_of(a)
            vals.append(broadcast_in_dim_p.bind(a, shape=shape, broadcast_dimensions=tuple(range(1, len(shape))), sharding=None))
    return BatchTracer(prim.bind(*vals, dimension=params["dimension"] + 1))


def _b_broadcast_in_dim(prim, args, params, B):
    return BatchTracer(prim.bind(args[0].val, shape=(B,) + tuple(params["shape"]), sharding=None,
                                 broadcast_dimensions=(0,) + tuple(d + 1 for d in params["broadcast_dimensions"])))


def _b_transpose(prim, args, params, B):
    return BatchTracer(prim.bind(args[0].val, permutation=(0,) + tuple(p + 1 for p in params["permutation"])))


def _b_reshape(prim, args, params, B):
    return BatchTracer(prim.bind(args[0].val, new_sizes=(B,) + tuple(params["new_sizes"]), dimensions=None, sharding=None))


def _b_squeeze(prim, args, params, B):
    return BatchTracer(prim.bind(args[0].val, dimensions=tuple(d + 1 for d in params["dimensions"])))


def _b_scatter(prim, args, params, B):
    """The windowed-assignment form of `_scatter_impl` with a leading batch axis on operand and / or updates (static
    indices): the batch axis becomes one more window dimension that starts at 0 - `jax.vmap(graphax.jacve(f))`
    reaches it through the slice / concatenate transforms (`primitives.py:744-749, 1062-1067`)."""
    operand, idx, upd = args
    if isinstance(idx, BatchTracer):
        raise NotImplementedError("jaxlite: scatter with batched indices")

    def lead(a):
        if isinstance(a, BatchTracer):
            return a.val
        shape = (B,) + _shape_of(a)
        return broadcast_in_dim_p.bind(_lift(a), shape=shape, broadcast_dimensions=tuple(range(1, len(shape))), sharding=None)
    dn = params["dimension_numbers"]
    new = ScatterDimensionNumbers((0,) + tuple(d + 1 for d in dn.update_window_dims), (),
                                  tuple(d + 1 for d in dn.scatter_dims_to_operand_dims))
    return BatchTracer(prim.bind(lead(operand), idx, lead(upd), dimension_numbers=new))


_BATCH_RULES = {"scatter": _b_scatter, "reduce_sum": _b_reduce, "reduce_max": _b_reduce, "reduce_min": _b_reduce, "dot_general": _b_dot_general,
                "slice": _b_slice, "concatenate": _b_concatenate, "broadcast_in_dim": _b_broadcast_in_dim,
                "transpose": _b_transpose, "reshape": _b_reshape, "squeeze": _b_squeeze}


def vmap(fun, in_axes=0, out_axes=0):
    """`jax.vmap` for `in_axes` of 0 / None and `out_axes=0`."""
    if out_axes != 0:
        raise NotImplementedError("jaxlite: vmap out_axes")

    def batched(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if any(a not in (0, None) for a in axes):
            raise NotImplementedError("jaxlite: vmap in_axes other than 0 / None")
        args = [a if _is_value(a) else Array(a) for a in args]
        B = next(a.shape[0] for a, ax in zip(args, axes) if ax == 0)
        outs = fun(*[BatchTracer(a) if ax == 0 else a for a, ax in zip(args, axes)])

        def unwrap(o):
            if isinstance(o, BatchTracer):
                return o.val
            shape = (B,) + _shape_of(o)       # an unbatched output is broadcast along the new axis
            return broadcast_in_dim_p.bind(_lift(o), shape=shape, broadcast_dimensions=tuple(range(1, len(shape))), sharding=None)
        return tree_map(unwrap, outs)
    return batched

jacfwd = _unavailable("jacfwd")
jacrev = _unavailable("jacrev")
grad = _unavailable("grad")


class custom_jvp:
    """Accepted as a decorator (the reference's SNN example defines one at import time, `neuromorphic.py:8`);
    calling the wrapped function under a trace is not supported."""

    def __init__(self, fun):
        self.fun = fun

    def defjvp(self, jvp):
        self.jvp = jvp
        return jvp

    def __call__(self, *args):
        if _TRACES:
            raise NotImplementedError("jaxlite: custom_jvp under a trace")
        return self.fun(*args)


@_concrete
def arange(start, stop=None, step=1, dtype=None):
    v = np.arange(start, stop, step) if stop is not None else np.arange(start)
    return Array(v if dtype is None else v.astype(_canon_dtype(dtype)))
