"""Mini-jaxpr tracer.

graphax numbers the vertices of the computational graph by the position of an
equation in ``jax.make_jaxpr(f)(*xs).jaxpr.eqns`` (reference ``core.py:184``,
``core.py:380-385``), so a caller-given elimination order only means something
relative to how JAX lowers the Python function.  JAX is not part of this image,
therefore this module re-creates that lowering for the ``jnp`` subset the
workloads use (rules: SURVEY.md Appendix A.1):

* one equation per array op, in Python evaluation order, no CSE;
* Python scalars become ``Literal`` operands that keep their source position;
* ``x**2`` -> ``integer_pow``, ``jnp.sum`` -> ``reduce_sum``, 1-D ``jnp.dot`` ->
  ``dot_general``, static ``x[a:b]`` -> ``slice``, ``jnp.zeros(n)`` ->
  ``broadcast_in_dim`` of a literal;
* rank-0 / size-1 operands rely on lax broadcasting (a single binary equation).

``vmap_numbering=True`` additionally records where ``jax.vmap`` would insert
identity ``broadcast_in_dim`` equations (Appendix A.1 rule 5) so that orders
written against ``jacve(jax.vmap(f))`` (reference ``example_test.py:150-159``)
can be translated to the per-sample numbering this engine runs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

Shape = Tuple[int, ...]


class Var:
    """A traced value: either a function input or the result of one equation."""
    __slots__ = ("shape", "eqn_index", "input_index", "name")

    def __init__(self, shape: Shape, eqn_index: Optional[int] = None,
                 input_index: Optional[int] = None, name: str = ""):
        self.shape = tuple(int(s) for s in shape)
        self.eqn_index = eqn_index      # 0-based position in Jaxpr.eqns (vertex id - 1)
        self.input_index = input_index  # position in Jaxpr.invars
        self.name = name

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    def __repr__(self) -> str:
        tag = f"in{self.input_index}" if self.input_index is not None else f"v{self.eqn_index + 1}"
        return f"{tag}{list(self.shape)}"


@dataclass(frozen=True)
class Literal:
    """A trace-time constant operand (weak-typed Python scalar -> f32)."""
    val: float

    @property
    def shape(self) -> Shape:
        return ()

    def __repr__(self) -> str:
        return repr(self.val)


Atom = Union[Var, Literal]


@dataclass
class Eqn:
    primitive: str
    invars: Tuple[Atom, ...]
    outvar: Var
    params: Dict[str, Any] = field(default_factory=dict)

    def __repr__(self) -> str:
        p = "" if not self.params else "[" + ",".join(f"{k}={v}" for k, v in self.params.items()) + "]"
        return f"{self.outvar!r} = {self.primitive}{p}({', '.join(map(repr, self.invars))})"


@dataclass
class Jaxpr:
    invars: List[Var]
    eqns: List[Eqn]
    outvars: List[Atom]
    # vmap_pads[i] = number of identity broadcast_in_dim equations jax.vmap
    # would insert immediately before equation i (0-based)
    vmap_pads: List[int] = field(default_factory=list)
    out_is_sequence: bool = True
    # arrays the function closes over (jax: ClosedJaxpr.consts / jaxpr.constvars, reference core.py:374): variables
    # with known values - sources of edges like the inputs (their input_index continues after the arguments), never
    # differentiated
    constvars: List[Var] = field(default_factory=list)
    consts: List[np.ndarray] = field(default_factory=list)

    def vertex_of(self, var: Var) -> int:
        return var.eqn_index + 1

    @property
    def output_vertices(self) -> List[int]:
        return [v.eqn_index + 1 for v in self.outvars if isinstance(v, Var) and v.eqn_index is not None]

    def to_vmap_numbering(self) -> Tuple[Dict[int, int], List[int]]:
        """Per-sample vertex id -> id under ``jacve(jax.vmap(f))`` and the list
        of ids that only exist in the vmapped jaxpr (identity pads)."""
        mapping, pads, shift = {}, [], 0
        for i in range(len(self.eqns)):
            for _ in range(self.vmap_pads[i] if self.vmap_pads else 0):
                shift += 1
                pads.append(i + shift)
            mapping[i + 1] = i + 1 + shift
        return mapping, pads

    def translate_vmap_order(self, order: Sequence[int]) -> List[int]:
        """Map an order written in the ``jacve(vmap(f))`` numbering onto the
        per-sample numbering by dropping the identity pad vertices."""
        mapping, pads = self.to_vmap_numbering()
        inverse = {v: k for k, v in mapping.items()}
        padset = set(pads)
        out = []
        for o in order:
            o = int(o)
            if o in padset:
                continue
            if o not in inverse:
                raise ValueError(f"vertex {o} does not exist in the vmapped numbering")
            out.append(inverse[o])
        return out

    def pretty(self) -> str:
        lines = ["{ lambda " + " ".join(map(repr, self.constvars)) + " ; " + " ".join(map(repr, self.invars)) + " . let"]
        for i, e in enumerate(self.eqns, start=1):
            lines.append(f"  {i:4d}: {e!r}")
        lines.append("  in (" + ", ".join(map(repr, self.outvars)) + ") }")
        return "\n".join(lines)


class _Trace:
    def __init__(self, n_args: int = 0) -> None:
        self.eqns: List[Eqn] = []
        self.vmap_pads: List[int] = []
        self.n_args = n_args
        self.constvars: List[Var] = []
        self.consts: List[np.ndarray] = []
        self._const_of: Dict[int, Tuple[Any, Var]] = {}     # id(object) -> (object kept alive, its variable)

    def constant(self, value: Any) -> Var:
        """The variable of an array the traced function closes over (one per object, like jax's constvars)."""
        hit = self._const_of.get(id(value))
        if hit is not None:
            return hit[1]
        def has_tracer(v) -> bool:
            return isinstance(v, Tracer) or (isinstance(v, (list, tuple)) and any(has_tracer(e) for e in v))
        if has_tracer(value):     # (NumPy would index the tracers while probing the sequence and stage equations)
            raise TypeError("a sequence of traced values is not an array: use jnp.concatenate / jnp.stack-like operations")
        arr = value.value if isinstance(value, ConstArray) else \
            value.detach().cpu().numpy() if hasattr(value, "detach") else np.asarray(value)
        if arr.dtype == object or not (np.issubdtype(arr.dtype, np.number) or arr.dtype == bool):
            raise TypeError(f"cannot trace operand of type {type(value).__name__}")
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        var = Var(arr.shape, input_index=self.n_args + len(self.constvars), name=f"const{len(self.constvars)}")
        self.constvars.append(var)
        self.consts.append(arr)
        self._const_of[id(value)] = (value, var)
        return var

    def emit(self, primitive: str, invars: Sequence[Atom], out_shape: Shape,
             pads: int = 0, **params) -> "Tracer":
        out = Var(out_shape, eqn_index=len(self.eqns))
        self.eqns.append(Eqn(primitive, tuple(invars), out, dict(params)))
        self.vmap_pads.append(pads)
        return Tracer(self, out)


_ACTIVE: List[_Trace] = []


def _current() -> _Trace:
    if not _ACTIVE:
        raise RuntimeError("graphax_b200.jnp functions must be called on traced values inside jacve/make_jaxpr")
    return _ACTIVE[-1]


def _atom(x: Any) -> Atom:
    if isinstance(x, Tracer):
        return x.var
    if isinstance(x, (bool, int, float, np.floating, np.integer)):
        return Literal(float(x))
    if isinstance(x, np.ndarray) and x.shape == ():
        return Literal(float(x))
    if _ACTIVE and (isinstance(x, (np.ndarray, list, tuple, ConstArray)) or (hasattr(x, "detach") and hasattr(x, "numpy"))):
        return _current().constant(x)       # an array the function closes over: a constvar, like under jax.make_jaxpr
    raise TypeError(f"cannot trace operand of type {type(x).__name__}")


def _bcast_shape(a: Shape, b: Shape) -> Shape:
    """lax-style broadcasting: equal ranks (size-1 dims stretch) or a rank-0 operand."""
    if a == ():
        return b
    if b == ():
        return a
    if len(a) != len(b):
        raise NotImplementedError(
            f"rank promotion between {a} and {b} is not lowered (jnp would insert reshape equations)")
    out = []
    for x, y in zip(a, b):
        if x != y and 1 not in (x, y):
            raise TypeError(f"incompatible shapes for broadcasting: {a} and {b}")
        out.append(max(x, y))
    return tuple(out)


class Tracer:
    """Array stand-in that records one equation per operation."""
    __slots__ = ("trace", "var")
    __array_priority__ = 1000

    def __init__(self, trace: _Trace, var: Var):
        self.trace = trace
        self.var = var

    shape = property(lambda self: self.var.shape)
    ndim = property(lambda self: self.var.ndim)
    size = property(lambda self: self.var.size)

    def __repr__(self) -> str:
        return f"Tracer<{self.var!r}>"

    # ---- binary arithmetic -------------------------------------------------
    def _binary(self, prim: str, other: Any, reflected: bool = False) -> "Tracer":
        a, b = _atom(self), _atom(other)
        if reflected:
            a, b = b, a
        if a.shape and b.shape and len(a.shape) != len(b.shape):
            # jnp rank promotion: the lower-rank operand gets leading size-1 axes through one
            # broadcast_in_dim equation (lax.broadcast_to_rank), e.g. [B,H] + [H] -> [B,H] + [1,H]
            lo, hi = (a, b) if len(a.shape) < len(b.shape) else (b, a)
            diff = len(hi.shape) - len(lo.shape)
            new_shape = (1,) * diff + tuple(lo.shape)
            promoted = self.trace.emit("broadcast_in_dim", (lo,), new_shape, shape=new_shape,
                                       broadcast_dimensions=tuple(range(diff, diff + len(lo.shape)))).var
            a, b = (promoted, b) if lo is a else (a, promoted)
        shape = _bcast_shape(a.shape, b.shape)
        # jax.vmap pads the lower-rank traced operand of a mixed-rank binary op
        pads = 1 if (isinstance(a, Var) and isinstance(b, Var) and len(a.shape) != len(b.shape)) else 0
        return self.trace.emit(prim, (a, b), shape, pads=pads)

    __add__ = lambda s, o: s._binary("add", o)
    __radd__ = lambda s, o: s._binary("add", o, True)
    __sub__ = lambda s, o: s._binary("sub", o)
    __rsub__ = lambda s, o: s._binary("sub", o, True)
    __mul__ = lambda s, o: s._binary("mul", o)
    __rmul__ = lambda s, o: s._binary("mul", o, True)
    __truediv__ = lambda s, o: s._binary("div", o)
    __rtruediv__ = lambda s, o: s._binary("div", o, True)

    def __neg__(self) -> "Tracer":
        return self.trace.emit("neg", (self.var,), self.shape)

    # comparisons (values 1.0 / 0.0: the tracer has one dtype); `==` stays identity for dict / list use
    __gt__ = lambda s, o: s._binary("gt", o)
    __lt__ = lambda s, o: s._binary("lt", o)
    __ge__ = lambda s, o: s._binary("ge", o)
    __le__ = lambda s, o: s._binary("le", o)

    def __pow__(self, y: Any) -> "Tracer":
        if isinstance(y, (int, np.integer)) and not isinstance(y, bool):
            return self.trace.emit("integer_pow", (self.var,), self.shape, y=int(y))
        if isinstance(y, float) and float(y).is_integer():
            # jnp.power with a Python float exponent stays `pow` against a literal
            return self._binary("pow", y)
        return self._binary("pow", y)

    def __rpow__(self, x: Any) -> "Tracer":
        return self._binary("pow", x, True)

    # ---- static indexing ---------------------------------------------------
    def __getitem__(self, idx: Any) -> "Tracer":
        """Static basic indexing: slices become one ``slice`` equation, integer indices a ``slice`` followed
        by a ``squeeze`` (how jnp lowers static indexing without gathers)."""
        idx = idx if isinstance(idx, tuple) else (idx,)
        if len(idx) > self.ndim or any(not isinstance(i, (slice, int, np.integer)) for i in idx):
            raise NotImplementedError("only static basic indexing (slices and integers) is lowered")
        idx = idx + (slice(None),) * (self.ndim - len(idx))
        starts, limits, squeeze = [], [], []
        for ax, (i, n) in enumerate(zip(idx, self.shape)):
            if isinstance(i, slice):
                a, b, step = i.indices(n)
                if step != 1:
                    raise NotImplementedError("strided slices are not lowered")
                starts.append(a)
                limits.append(max(b, a))
            else:
                i = int(i) + (n if int(i) < 0 else 0)
                if not 0 <= i < n:
                    raise IndexError(f"index {i} out of bounds for axis {ax} of size {n}")
                starts.append(i)
                limits.append(i + 1)
                squeeze.append(ax)
        out = self
        if any(a != 0 or b != n for a, b, n in zip(starts, limits, self.shape)):
            out = self.trace.emit("slice", (self.var,), tuple(b - a for a, b in zip(starts, limits)),
                                  start_indices=tuple(starts), limit_indices=tuple(limits), strides=None)
        if squeeze:
            shape = tuple(s for ax, s in enumerate(out.shape) if ax not in squeeze)
            out = out.trace.emit("squeeze", (out.var,), shape, dimensions=tuple(squeeze))
        return out

    def reshape(self, *shape) -> "Tracer":
        shape = tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else tuple(shape)
        shape = tuple(int(s) for s in shape)
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1], dtype=np.int64))
            shape = tuple(self.size // known if s == -1 else s for s in shape)
        if int(np.prod(shape, dtype=np.int64)) != self.size:
            raise TypeError(f"cannot reshape {self.shape} into {shape}")
        return self.trace.emit("reshape", (self.var,), shape, new_sizes=shape, dimensions=None)

    @property
    def T(self) -> "Tracer":
        return _transpose(self)

    def __matmul__(self, other) -> "Tracer":
        return _matmul(self, other)

    def __rmatmul__(self, other) -> "Tracer":
        return _matmul(other, self)

    def sum(self, axis=None, keepdims: bool = False) -> "Tracer":
        return _reduce_sum(self, axis, keepdims)

    def mean(self, axis=None) -> "Tracer":
        return _mean(self, axis)

    def dot(self, other) -> "Tracer":
        return _dot(self, other)


class ConstArray:
    """A concrete array inside a traced function (``jnp.asarray(np_array)``): like a jax Array under ``make_jaxpr`` it
    becomes a constant of the trace (a constvar) the first time an operation uses it, and every operation on it is
    staged.  Outside a trace it behaves like its NumPy value."""
    __slots__ = ("value",)
    __array_priority__ = 1000

    def __init__(self, value: Any):
        value = value.detach().cpu().numpy() if hasattr(value, "detach") else np.asarray(value)
        self.value = np.ascontiguousarray(value, dtype=np.float32)

    shape = property(lambda self: self.value.shape)
    ndim = property(lambda self: self.value.ndim)
    size = property(lambda self: self.value.size)

    def __array__(self, dtype=None, copy=None):
        return self.value if dtype is None else self.value.astype(dtype)

    def _operand(self):
        return Tracer(_current(), _current().constant(self)) if _ACTIVE else self.value

    def __repr__(self) -> str:
        return f"ConstArray({self.value!r})"


def _forward_to_operand(name: str):
    def method(self, *args, **kwargs):
        args = [a.value if isinstance(a, ConstArray) and not _ACTIVE else a for a in args]
        return getattr(self._operand(), name)(*args, **kwargs)
    method.__name__ = name
    return method


for _name in ("__neg__", "__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__",
              "__rtruediv__", "__pow__", "__rpow__", "__matmul__", "__rmatmul__", "__getitem__", "__gt__", "__lt__",
              "sum", "reshape", "transpose"):
    if hasattr(Tracer, _name):
        setattr(ConstArray, _name, _forward_to_operand(_name))
ConstArray.T = property(lambda self: self._operand().T)


_CONST_UNARY = {"sqrt": np.sqrt, "abs": np.abs, "neg": np.negative, "exp": np.exp, "log": np.log, "sin": np.sin,
                "cos": np.cos, "tan": np.tan, "tanh": np.tanh}


def _unary(prim: str):
    def fn(x: Any) -> Tracer:
        x = _lift(x)
        if not isinstance(x, Tracer):
            if isinstance(x, (int, float, np.floating, np.integer)):
                if _ACTIVE:
                    # inside a traced function JAX stages every operation, constants included:
                    # jnp.sqrt(2.) is the equation `sqrt 2.0` - a vertex without in-edges (it still counts
                    # in the numbering and must appear in a custom order, like jnp.zeros)
                    return _current().emit(prim, (Literal(float(x)),), ())
                if prim in _CONST_UNARY:
                    return float(np.float32(_CONST_UNARY[prim](np.float32(x))))
            raise TypeError(f"{prim} expects a traced value")
        return x.trace.emit(prim, (x.var,), x.shape)
    fn.__name__ = prim
    return fn


def _binary_fn(prim: str):
    def fn(x: Any, y: Any) -> Tracer:
        x = _lift(x) if not isinstance(y, Tracer) else x
        if isinstance(x, Tracer):
            return x._binary(prim, y)
        if isinstance(y, Tracer):
            return y._binary(prim, x, True)
        raise TypeError(f"{prim} expects at least one traced value")
    fn.__name__ = prim
    return fn


def _keepdims(x: Tracer, reduced: Tracer, axes) -> Tracer:
    """``keepdims=True``: jnp re-inserts the reduced axes with a broadcast_in_dim of the result."""
    shape = tuple(1 if i in axes else s for i, s in enumerate(x.shape))
    bdims = tuple(i for i in range(x.ndim) if i not in axes)
    return x.trace.emit("broadcast_in_dim", (reduced.var,), shape, shape=shape, broadcast_dimensions=bdims)


def _array_first(fn):
    """The first argument may be an array the function closes over: it enters the trace as a constant."""
    import functools

    @functools.wraps(fn)
    def wrapped(x, *args, **kwargs):
        return fn(_lift(x), *args, **kwargs)
    return wrapped


@_array_first
def _reduce_sum(x: Tracer, axis=None, keepdims: bool = False) -> Tracer:
    if keepdims:
        r = _reduce_sum(x, axis)
        axes = tuple(range(x.ndim)) if axis is None else tuple(sorted(a % x.ndim for a in ((axis,) if isinstance(axis, int) else axis)))
        return _keepdims(x, r, axes)
    if axis is None:
        axes = tuple(range(x.ndim))
    elif isinstance(axis, int):
        axes = (axis % max(x.ndim, 1),)
    else:
        axes = tuple(sorted(a % x.ndim for a in axis))
    shape = tuple(s for i, s in enumerate(x.shape) if i not in axes)
    return x.trace.emit("reduce_sum", (x.var,), shape, axes=axes)


def _lift(x: Any) -> Any:
    """An array the function closes over as a traced constant (numbers and traced values pass through)."""
    if isinstance(x, Tracer) or not _ACTIVE or isinstance(x, (bool, int, float, np.floating, np.integer)):
        return x
    if isinstance(x, (np.ndarray, list, tuple, ConstArray)) or (hasattr(x, "detach") and hasattr(x, "numpy")):
        trace = _current()
        return Tracer(trace, trace.constant(x))
    return x


def _dot(a: Tracer, b: Tracer) -> Tracer:
    a, b = _lift(a), _lift(b)
    if not (isinstance(a, Tracer) and isinstance(b, Tracer)):
        raise TypeError("dot expects two traced values")
    if a.ndim > 1 or b.ndim > 1:
        return _matmul(a, b)
    if a.ndim == 1 and b.ndim == 1:
        if a.shape != b.shape:
            raise TypeError(f"dot: incompatible shapes {a.shape} and {b.shape}")
        dn = (((0,), (0,)), ((), ()))
        return a.trace.emit("dot_general", (a.var, b.var), (), dimension_numbers=dn,
                            pads=0)
    raise NotImplementedError("only 1-D . 1-D dot products are lowered")


def _dot_general(a: Tracer, b: Tracer, dimension_numbers, precision=None, preferred_element_type=None) -> Tracer:
    """``lax.dot_general``: out = batch dims, then free dims of a, then free dims of b."""
    a, b = _lift(a), _lift(b)
    (ca, cb), (ba, bb) = dimension_numbers
    ca, cb = tuple(int(c) for c in ca), tuple(int(c) for c in cb)
    ba, bb = tuple(int(c) for c in ba), tuple(int(c) for c in bb)
    if [a.shape[i] for i in ca] != [b.shape[i] for i in cb] or [a.shape[i] for i in ba] != [b.shape[i] for i in bb]:
        raise TypeError(f"dot_general: dimension sizes differ: {a.shape} {ca} {ba} vs {b.shape} {cb} {bb}")
    shape = tuple(a.shape[i] for i in ba) + \
        tuple(s for i, s in enumerate(a.shape) if i not in ca and i not in ba) + \
        tuple(s for i, s in enumerate(b.shape) if i not in cb and i not in bb)
    return a.trace.emit("dot_general", (a.var, b.var), shape, dimension_numbers=((ca, cb), (ba, bb)))


def _matmul(a: Tracer, b: Tracer) -> Tracer:
    """``a @ b`` / ``jnp.dot`` for ranks 1 and 2: one dot_general contracting the last axis of a with the
    first (for a 1-D b: only) axis of b."""
    a, b = _lift(a), _lift(b)
    if not (isinstance(a, Tracer) and isinstance(b, Tracer)) or not (1 <= a.ndim <= 2 and 1 <= b.ndim <= 2):
        raise NotImplementedError("matmul of ranks 1 and 2 only")
    return _dot_general(a, b, (((a.ndim - 1,), (0,)), ((), ())))


@_array_first
def _transpose(x: Tracer, axes=None) -> Tracer:
    perm = tuple(range(x.ndim))[::-1] if axes is None else tuple(int(a) for a in axes)
    if x.ndim < 2:
        return x
    return x.trace.emit("transpose", (x.var,), tuple(x.shape[p] for p in perm), permutation=perm)


def _outer(a: Tracer, b: Tracer) -> Tracer:
    """jnp.outer: ravel both, then multiply [n,1] * [1,m] (two broadcast_in_dim-free reshapes in jnp)."""
    a, b = _lift(a), _lift(b)
    return a.reshape(a.size, 1) * b.reshape(1, b.size)


@_array_first
def _squeeze(x: Tracer, axis=None) -> Tracer:
    axes = tuple(i for i, s in enumerate(x.shape) if s == 1) if axis is None else \
        tuple(sorted(a % x.ndim for a in ((axis,) if isinstance(axis, int) else axis)))
    return x.trace.emit("squeeze", (x.var,), tuple(s for i, s in enumerate(x.shape) if i not in axes), dimensions=axes)


@_array_first
def _expand_dims(x: Tracer, axis: int) -> Tracer:
    axis = axis % (x.ndim + 1)
    shape = x.shape[:axis] + (1,) + x.shape[axis:]
    bdims = tuple(i for i in range(len(shape)) if i != axis)
    return x.trace.emit("broadcast_in_dim", (x.var,), shape, shape=shape, broadcast_dimensions=bdims)


def _reduce_extreme(prim: str):
    def fn(x: Tracer, axis=None, keepdims: bool = False) -> Tracer:
        x = _lift(x)
        axes = tuple(range(x.ndim)) if axis is None else tuple(sorted(a % x.ndim for a in ((axis,) if isinstance(axis, int) else axis)))
        shape = tuple(s for i, s in enumerate(x.shape) if i not in axes)
        r = x.trace.emit(prim, (x.var,), shape, axes=axes)
        return _keepdims(x, r, axes) if keepdims else r
    return fn


@_array_first
def _mean(x: Tracer, axis=None) -> Tracer:
    """jnp.mean: reduce_sum, then a division by the literal element count."""
    axes = tuple(range(x.ndim)) if axis is None else tuple(a % x.ndim for a in ((axis,) if isinstance(axis, int) else axis))
    n = 1
    for a in axes:
        n *= x.shape[a]
    return _reduce_sum(x, axis) / float(n)


def _where(cond: Any, x: Any, y: Any) -> Tracer:
    """``jnp.where(cond, x, y)`` = ``lax.select(cond, x, y)`` = ``select_n(cond, y, x)``; scalars broadcast."""
    if not isinstance(cond, Tracer):
        raise TypeError("where expects a traced condition")
    trace = cond.trace
    atoms = [_atom(cond), _atom(y), _atom(x)]
    shape = ()
    for a in atoms:
        shape = _bcast_shape(shape, a.shape)
    return trace.emit("select_n", tuple(atoms), shape)


def _stop_gradient(x: Tracer) -> Tracer:
    return x.trace.emit("stop_gradient", (x.var,), x.shape)


def _softmax(x: Tracer, axis=-1) -> Tracer:
    """jax.nn.softmax without its custom_jvp wrapper (the reference has no rule for custom_jvp_call):
    exp(x - stop_gradient(max(x))) / sum(...), both reductions with keepdims."""
    x_max = _reduce_extreme("reduce_max")(x, axis, keepdims=True)
    unnormalized = _unary("exp")(x - _stop_gradient(x_max))
    return unnormalized / _reduce_sum(unnormalized, axis, keepdims=True)


class _LaxShim:
    """The slice of ``jax.lax`` the workloads touch."""
    dot_general = staticmethod(_dot_general)
    stop_gradient = staticmethod(_stop_gradient)


lax = _LaxShim()


class _NnShim:
    """The slice of ``jax.nn`` the reference's deep-learning examples touch."""
    softmax = staticmethod(lambda x, axis=-1: _softmax(x, axis))
    sigmoid = staticmethod(lambda x: _unary("logistic")(x))
    tanh = staticmethod(lambda x: _unary("tanh")(x))
    relu = staticmethod(lambda x: _binary_fn("max")(x, 0.0))      # jax.nn.relu = maximum(x, 0)


jnn = _NnShim()


@_array_first
def _square(x: Tracer) -> Tracer:
    return x.trace.emit("square", (x.var,), x.shape)


def _zeros(shape, dtype=None) -> Tracer:
    shape = (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)
    return _current().emit("broadcast_in_dim", (Literal(0.0),), shape,
                           shape=shape, broadcast_dimensions=())


def _full(shape, value: float) -> Tracer:
    shape = (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)
    return _current().emit("broadcast_in_dim", (Literal(float(value)),), shape, shape=shape, broadcast_dimensions=())


def _concatenate(parts: Sequence[Tracer], axis: int = 0) -> Tracer:
    parts = [_lift(p) for p in parts]
    if not parts or not all(isinstance(p, Tracer) for p in parts):
        raise TypeError("concatenate expects a sequence of traced values")
    if any(p.ndim != parts[0].ndim for p in parts):
        raise TypeError("concatenate: operands must have the same rank")
    if parts[0].ndim > 1:
        axis = axis % parts[0].ndim
        shape = list(parts[0].shape)
        shape[axis] = sum(p.shape[axis] for p in parts)
        return parts[0].trace.emit("concatenate", tuple(p.var for p in parts), tuple(shape), dimension=axis)
    if axis not in (0, -1):
        raise TypeError("axis out of range for 1-D concatenation")
    n = sum(p.shape[0] for p in parts)
    # vmap broadcasts an un-batched operand (the literal zeros) before concatenating
    pads = sum(1 for p in parts
               if p.var.eqn_index is not None
               and p.trace.eqns[p.var.eqn_index].primitive == "broadcast_in_dim"
               and all(isinstance(a, Literal) for a in p.trace.eqns[p.var.eqn_index].invars))
    return parts[0].trace.emit("concatenate", tuple(p.var for p in parts), (n,), pads=pads,
                               dimension=0)


class _JnpShim:
    """The slice of ``jax.numpy`` the reference workloads touch, over Tracers."""
    float32 = np.float32
    pi = math.pi

    sqrt = staticmethod(_unary("sqrt"))
    abs = staticmethod(_unary("abs"))
    sin = staticmethod(_unary("sin"))
    cos = staticmethod(_unary("cos"))
    tan = staticmethod(_unary("tan"))
    exp = staticmethod(_unary("exp"))
    log = staticmethod(_unary("log"))
    log1p = staticmethod(_unary("log1p"))
    tanh = staticmethod(_unary("tanh"))
    sinh = staticmethod(_unary("sinh"))
    cosh = staticmethod(_unary("cosh"))
    arcsin = staticmethod(_unary("asin"))
    arccos = staticmethod(_unary("acos"))
    arctan = staticmethod(_unary("atan"))
    arcsinh = staticmethod(_unary("asinh"))
    arccosh = staticmethod(_unary("acosh"))
    arctanh = staticmethod(_unary("atanh"))
    negative = staticmethod(_unary("neg"))
    square = staticmethod(_square)

    add = staticmethod(_binary_fn("add"))
    subtract = staticmethod(_binary_fn("sub"))
    multiply = staticmethod(_binary_fn("mul"))
    divide = staticmethod(_binary_fn("div"))
    arctan2 = staticmethod(_binary_fn("atan2"))
    power = staticmethod(_binary_fn("pow"))
    greater = staticmethod(_binary_fn("gt"))      # 1.0 / 0.0 (the tracer has one dtype); zero partials
    less = staticmethod(_binary_fn("lt"))
    equal = staticmethod(_binary_fn("eq"))
    maximum = staticmethod(_binary_fn("max"))
    minimum = staticmethod(_binary_fn("min"))
    erf = staticmethod(_unary("erf"))
    sigmoid = staticmethod(_unary("logistic"))

    sum = staticmethod(_reduce_sum)
    mean = staticmethod(_mean)
    max = staticmethod(_reduce_extreme("reduce_max"))
    min = staticmethod(_reduce_extreme("reduce_min"))
    dot = staticmethod(_dot)
    matmul = staticmethod(_matmul)
    outer = staticmethod(_outer)
    transpose = staticmethod(_transpose)
    squeeze = staticmethod(_squeeze)
    expand_dims = staticmethod(_expand_dims)

    @staticmethod
    def reshape(x, shape):
        return x.reshape(shape)

    @staticmethod
    def asarray(x, dtype=None):
        """A traced value as it is; numbers stay weak-typed scalars; an array becomes a constant of the trace."""
        if isinstance(x, (Tracer, ConstArray)) or isinstance(x, (bool, int, float, np.floating, np.integer)):
            return x
        return ConstArray(x)
    array = asarray
    zeros = staticmethod(_zeros)
    concatenate = staticmethod(_concatenate)
    where = staticmethod(_where)

    @staticmethod
    def split(x, indices_or_sections, axis: int = 0):
        """jnp.split: a chain of slices (reference tests/core/multi_output_primitives_test.py:14-21)."""
        axis = axis % x.ndim
        n = x.shape[axis]
        if isinstance(indices_or_sections, (int, np.integer)):
            if n % int(indices_or_sections):
                raise ValueError("array split does not result in an equal division")
            step = n // int(indices_or_sections)
            cuts = list(range(step, n, step))
        else:
            cuts = [int(i) for i in indices_or_sections]
        parts, lo = [], 0
        for hi in cuts + [n]:
            parts.append(x[tuple(slice(None) if d != axis else slice(lo, hi) for d in range(x.ndim))])
            lo = hi
        return parts

    @staticmethod
    def ones(shape, dtype=None):
        """``jnp.ones(())`` folds to a literal under make_jaxpr (function g)."""
        if shape == () or shape == []:
            return 1.0
        return _full(shape, 1.0)

    @staticmethod
    def full(shape, fill_value, dtype=None):
        """``jnp.full``: like ``jnp.zeros``, one ``broadcast_in_dim`` of a literal - a vertex without in-edges."""
        if isinstance(fill_value, Tracer):
            raise NotImplementedError("full() with a traced fill value")
        return _full(shape, float(fill_value))

    @staticmethod
    def zeros_like(x, dtype=None):
        return _full(tuple(np.shape(x)) if not isinstance(x, (Tracer, ConstArray)) else x.shape, 0.0)

    @staticmethod
    def ones_like(x, dtype=None):
        return _full(tuple(np.shape(x)) if not isinstance(x, (Tracer, ConstArray)) else x.shape, 1.0)


jnp = _JnpShim()


FRONTENDS = ("auto", "native", "jax")


def make_jaxpr(fun: Callable, arg_shapes: Sequence[Shape], frontend: str = "auto") -> Jaxpr:
    """Trace ``fun`` on per-sample avals of the given shapes (f32).

    ``frontend="native"`` runs ``fun`` on this module's tracers (functions written against ``graphax_b200.jnp``);
    ``"jax"`` traces with ``jax.make_jaxpr`` and converts the result (``jax_frontend.py``: functions written against
    the real ``jax.numpy``, needs JAX); ``"auto"`` tries the native tracers and falls back to JAX when the function
    rejects them and JAX is installed."""
    if frontend not in FRONTENDS:
        raise ValueError(f"frontend must be one of {FRONTENDS}")
    if frontend == "jax":
        from . import jax_frontend
        return jax_frontend.make_jaxpr(fun, arg_shapes)
    try:
        return _make_jaxpr_native(fun, arg_shapes)
    except (TypeError, AttributeError) as e:
        if frontend == "auto":
            from . import jax_frontend
            if jax_frontend.available():
                try:
                    return jax_frontend.make_jaxpr(fun, arg_shapes)
                except (TypeError, AttributeError):
                    raise e from None      # the function takes neither kind of tracer: report the native error
        raise


def _make_jaxpr_native(fun: Callable, arg_shapes: Sequence[Shape]) -> Jaxpr:
    trace = _Trace(len(arg_shapes))
    invars = [Var(tuple(s), input_index=i) for i, s in enumerate(arg_shapes)]
    _ACTIVE.append(trace)
    try:
        outs = fun(*[Tracer(trace, v) for v in invars])
    finally:
        _ACTIVE.pop()
    out_is_sequence = isinstance(outs, (tuple, list))
    if not out_is_sequence:
        outs = (outs,)
    outvars: List[Atom] = []
    for o in outs:
        if not isinstance(o, Tracer):
            raise TypeError("function outputs must be traced arrays")
        outvars.append(o.var)
    return Jaxpr(invars, trace.eqns, outvars, trace.vmap_pads, out_is_sequence, trace.constvars, trace.consts)
