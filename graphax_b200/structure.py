"""Tensor-level structure of edge Jacobians (no values).

Every edge of the computational graph carries, next to its scalar entries, the
structure descriptor graphax would give it: ``out_dims`` / ``primal_dims`` made
of dense and sparse (Kronecker) dimensions with their ``val_dim`` into the
compact value array (reference ``sparse/tensor.py:24-39, 44-85``).  The engine
does not need the descriptor to compute - entries are scalarised - but it is
what ``count_ops`` is defined on (``tensor.py:1414-1493``), what selects the
sparsity class of a product (``tensor.py:302-371``), and what the plan exports
for the bit-exact structure comparison with the oracle.

Scope: vertices of rank <= 1, i.e. every edge has at most one out dimension and
one primal dimension (all of configs 1-4).  Higher ranks
raise ``NotImplementedError`` (row (f)1 of the scope table).

Dimension ids are implicit at rank <= 1: the out dimension has id 0, the primal
dimension id ``len(out_dims)``; a sparse pair always points at each other.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import List, Optional, Tuple

DENSE, SPARSE = "D", "S"


@dataclass(frozen=True)
class Dim:
    kind: str                 # "D" dense | "S" sparse (one half of a Kronecker pair)
    size: int
    val_dim: Optional[int]    # axis of the compact value array, None = replicated / pure delta


@dataclass(frozen=True)
class Transform:
    """A deferred shape-op Jacobian (reference ``primitives.py:469-489``).

    kind "slice":  out = x[lo:hi]            (``primitives.py:699-760``)
    kind "concat": out[lo:hi] = x, rest zero (``primitives.py:971-1233``)
    ``full`` is the size of the larger side (slice operand / concat result).
    """
    kind: str
    lo: int
    hi: int
    full: int

    @property
    def part(self) -> int:
        return self.hi - self.lo


@dataclass(frozen=True)
class Struct:
    out: Optional[Dim]
    primal: Optional[Dim]
    val_shape: Optional[Tuple[int, ...]]       # None <=> val is None (pure transform edge)
    pre: Tuple[Transform, ...] = ()            # pending pre_transforms, applied to whatever precedes

    # ---- views -----------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(d.size for d in (self.out, self.primal) if d is not None)

    @property
    def has_val(self) -> bool:
        return self.val_shape is not None

    @property
    def val_size(self) -> int:
        n = 1
        for s in self.val_shape or ():
            n *= s
        return n

    def descriptor(self) -> dict:
        """Reference-compatible description: (kind, id, size, val_dim, other_id) per dim."""
        l = 1 if self.out is not None else 0

        def enc(d: Dim, idx: int, other: int):
            return [d.kind, idx, d.size, d.val_dim, other if d.kind == SPARSE else None]
        return {
            "out_dims": [enc(self.out, 0, l)] if self.out is not None else [],
            "primal_dims": [enc(self.primal, l, 0)] if self.primal is not None else [],
            "val_shape": None if self.val_shape is None else list(self.val_shape),
            "pre_transforms": [[t.kind, t.lo, t.hi, t.full] for t in self.pre],
        }

    def sparsity_class(self) -> str:
        if not self.has_val:
            return "transform"
        o, p = self.out, self.primal
        if o is None and p is None:
            return "scalar"
        if o is not None and p is not None and o.kind == SPARSE:
            return "diag" if o.val_dim is not None else "kron"
        if p is not None and p.kind == DENSE and p.val_dim is None:
            return "bcast"
        return "dense"


SCALAR = Struct(None, None, ())


def diag(n: int) -> Struct:
    return Struct(Dim(SPARSE, n, 0), Dim(SPARSE, n, 0), (n,))


def kron_scalar(n: int) -> Struct:
    return Struct(Dim(SPARSE, n, None), Dim(SPARSE, n, None), ())


def bcast(n_out: int, n_in: int) -> Struct:
    return Struct(Dim(DENSE, n_out, 0), Dim(DENSE, n_in, None), (n_out,))


def rank0_operand(n_out: int) -> Struct:
    return Struct(Dim(DENSE, n_out, 0), None, (n_out,))


def row(n_in: int) -> Struct:
    return Struct(None, Dim(DENSE, n_in, 0), (n_in,))


def transform_edge(t: Transform) -> Struct:
    return Struct(None, None, None, (t,))


# --------------------------------------------------------------------------
# multiplication  post * pre   (reference tensor.py:334-371)
# --------------------------------------------------------------------------
def _axis(s: Struct, d: Optional[Dim]) -> int:
    """Extent of the value array along the axis ``d`` maps to (1 if replicated)."""
    if d is None or d.val_dim is None:
        return 1
    return s.val_shape[d.val_dim]


def mul_class(lhs: Struct, rhs: Struct) -> str:
    ld, rd = lhs.primal, rhs.out
    if (ld is None) != (rd is None) or (ld is not None and ld.size != rd.size):
        raise AssertionError(f"{lhs.shape} and {rhs.shape} not compatible for multiplication!")
    if lhs.shape == () and rhs.shape == ():
        return "scalar"
    if ld is None or (ld.kind == DENSE and rd.kind == DENSE):
        return "dot"          # _is_pure_dot_product_mul: all (possibly zero) pairs dense-dense
    return "bcast"            # a single pair with a sparse side; _mixed_mul needs rank >= 2


def mul(lhs: Struct, rhs: Struct) -> Struct:
    """Structure of ``lhs * rhs`` with ``lhs`` the edge out of the eliminated
    vertex and ``rhs`` the edge into it (reference ``core.py:210``)."""
    cls = mul_class(lhs, rhs)
    if cls == "scalar":
        return SCALAR
    ld, rd = lhs.primal, rhs.out
    if cls == "dot":
        # tensor.py:999-1059: out dims of lhs survive, primal dims of rhs become dense
        if lhs.out is not None and (lhs.out.kind != DENSE or lhs.out.val_dim is None):
            raise NotImplementedError("dot-product with a sparse/replicated out dimension")
        out = None if lhs.out is None else Dim(DENSE, lhs.out.size, 0)
        l = 0 if out is None else 1
        primal, shape = None, (() if out is None else (out.size,))
        if rhs.primal is not None:
            if rhs.primal.val_dim is not None:
                primal = Dim(DENSE, rhs.primal.size, l)
                shape = shape + (rhs.primal.size,)
            else:
                primal = Dim(DENSE, rhs.primal.size, None)
        return Struct(out, primal, shape)

    # pure broadcast multiplication, tensor.py:826-866 (+ _pad_tensors, _get_output_tensor)
    if ld.kind == SPARSE and rd.kind == SPARSE:
        if ld.val_dim is None and rd.val_dim is None:
            return Struct(Dim(SPARSE, ld.size, None), Dim(SPARSE, ld.size, None), ())
        n = max(_axis(lhs, ld), _axis(rhs, rd))
        return Struct(Dim(SPARSE, ld.size, 0), Dim(SPARSE, ld.size, 0), (n,))
    if ld.kind == SPARSE:            # diag/kron * dense-out: row scaling, result looks like rhs
        if rd.val_dim is None:
            raise NotImplementedError("row scaling of an edge with a replicated out dimension")
        out = Dim(DENSE, rd.size, 0)
        shape: Tuple[int, ...] = (rd.size,)
        primal = None
        if rhs.primal is not None:
            if rhs.primal.kind != DENSE:
                raise AssertionError("inconsistent sparse pairing")
            if rhs.primal.val_dim is not None:
                primal = Dim(DENSE, rhs.primal.size, 1)
                shape = shape + (_axis(rhs, rhs.primal),)
            else:
                primal = Dim(DENSE, rhs.primal.size, None)
        return Struct(out, primal, shape)
    # dense-primal * diag/kron: column scaling, result looks like lhs
    out, shape = None, ()
    if lhs.out is not None:
        if lhs.out.kind != DENSE or lhs.out.val_dim is None:
            raise NotImplementedError("column scaling of an edge with a sparse/replicated out dimension")
        out = Dim(DENSE, lhs.out.size, 0)
        shape = (lhs.out.size,)
    if ld.val_dim is None and rd.val_dim is None:
        primal = Dim(DENSE, ld.size, None)
    else:
        primal = Dim(DENSE, ld.size, len(shape))
        shape = shape + (max(_axis(lhs, ld), _axis(rhs, rd)),)
    return Struct(out, primal, shape)


def num_muls(lhs: Struct, rhs: Struct) -> int:
    """``get_num_muls`` (reference tensor.py:1414-1456)."""
    n = 1
    if lhs.out is not None and lhs.out.kind == DENSE and lhs.out.val_dim is not None:
        n *= lhs.out.size
    ld, rd = lhs.primal, rhs.out
    if ld is not None and rd is not None:
        if ld.kind == DENSE:
            n *= ld.size
        elif rd.kind == DENSE:
            n *= rd.size
        elif ld.val_dim is not None and rd.val_dim is not None:
            n *= max(lhs.val_shape[ld.val_dim], rhs.val_shape[rd.val_dim])
        elif ld.val_dim is not None:
            n *= lhs.val_shape[ld.val_dim]
        elif rd.val_dim is not None:
            n *= rhs.val_shape[rd.val_dim]
        elif lhs.has_val and rhs.has_val:
            if lhs.val_size == 1 and rhs.val_size == 1:
                n *= 1
            elif lhs.val_size == 1:
                n *= rd.size
            elif rhs.val_size == 1:
                n *= ld.size
    if rhs.primal is not None and rhs.primal.kind == DENSE and rhs.primal.val_dim is not None:
        n *= rhs.primal.size
    return n


# --------------------------------------------------------------------------
# addition (reference tensor.py:374-397, 1267-1411)
# --------------------------------------------------------------------------
def add(lhs: Struct, rhs: Struct) -> Struct:
    if lhs.shape != rhs.shape:
        raise AssertionError(f"{lhs.shape} and {rhs.shape} not compatible for addition!")
    if lhs.pre or rhs.pre:
        raise AssertionError("pending transforms must be materialised before addition")
    out = primal = None
    shape: List[int] = []
    sparse_pair = False
    if lhs.out is not None:
        lo, ro = lhs.out, rhs.out
        if lo.kind == SPARSE and ro.kind == SPARSE:
            sparse_pair = True
            out = Dim(SPARSE, lo.size, 0)
            shape.append(max(_axis(lhs, lo), _axis(rhs, ro)))
        else:
            out = Dim(DENSE, lo.size, 0)
            if lo.kind == DENSE and ro.kind == DENSE and lo.val_dim is None and ro.val_dim is None:
                shape.append(1)
            else:
                shape.append(lo.size)
    if lhs.primal is not None:
        lp, rp = lhs.primal, rhs.primal
        if lp.kind == SPARSE and rp.kind == SPARSE and sparse_pair:
            primal = Dim(SPARSE, lp.size, 0)
        else:
            primal = Dim(DENSE, lp.size, len(shape))
            if lp.kind == DENSE and rp.kind == DENSE and lp.val_dim is None and rp.val_dim is None:
                shape.append(1)
            else:
                shape.append(lp.size)
    return Struct(out, primal, tuple(shape))


def num_adds(lhs: Struct, rhs: Struct) -> int:
    """``get_num_adds`` (reference tensor.py:1460-1493); the reference calls it
    with ``lhs`` = the already-summed edge and ``rhs`` = the old edge (core.py:252-253)."""
    n = 1
    if lhs.out is not None:
        lo, ro = lhs.out, rhs.out
        if lo.kind == SPARSE and ro.kind == SPARSE:
            if lo.val_dim is not None and ro.val_dim is not None:
                n *= lo.size
        else:
            n *= lo.size
    if lhs.primal is not None:
        lp, rp = lhs.primal, rhs.primal
        if not (lp.kind == SPARSE and rp.kind == SPARSE):
            n *= lp.size
    return n


# --------------------------------------------------------------------------
# transforms
# --------------------------------------------------------------------------
def apply_transform(t: Transform, pre: Struct) -> Struct:
    """``transform.apply(pre)``: act on the *out* side of the incoming edge.
    The result carries no pending transforms (the reference builds a fresh
    SparseTensor); the caller re-attaches ``pre.pre`` (core.py:223-224)."""
    if not pre.has_val:
        raise AssertionError("transforms are only unloaded onto edges with values")
    if t.kind == "slice":
        # primitives.py:706-725: densify, then lax.slice -> all-dense result
        out = Dim(DENSE, t.part, 0)
        if pre.primal is None:
            return Struct(out, None, (t.part,))
        return Struct(out, Dim(DENSE, pre.primal.size, 1), (t.part, pre.primal.size))
    if t.kind == "concat":
        d = pre.out
        if d is None:
            raise IndexError("concatenate transform needs an out dimension")      # primitives.py:994
        if d.kind == DENSE:
            if d.val_dim is None:
                raise NotImplementedError("DenseDimension without `val_dim` not yet supported!")  # :1015
            shape = list(pre.val_shape)
            shape[d.val_dim] = t.full
            return Struct(replace(d, size=t.full), pre.primal, tuple(shape))
        # sparse pair gets materialised into a zero-padded dense block (:1017-1129)
        return Struct(Dim(DENSE, t.full, 0), Dim(DENSE, d.size, 1), (t.full, d.size))
    raise ValueError(t.kind)


def apply_inverse_transform(t: Transform, post: Struct) -> Struct:
    """``transform.apply_inverse(post)``: act on the *primal* side of the outgoing edge."""
    if not post.has_val:
        raise AssertionError("transforms are only unloaded onto edges with values")
    if t.kind == "slice":
        # primitives.py:727-755: scatter the dense block into zeros of the operand's shape
        out = None if post.out is None else Dim(DENSE, post.out.size, 0)
        l = 0 if out is None else 1
        shape = (() if out is None else (out.size,)) + (t.full,)
        return Struct(out, Dim(DENSE, t.full, l), shape, ())
    if t.kind == "concat":
        d = post.primal
        if d is None:
            raise AttributeError("inverse concatenate transform needs a primal dimension")  # :1146
        if d.kind == DENSE:
            if d.val_dim is None:
                raise NotImplementedError("DenseDimension without `val_dim` not yet supported!")  # :1146
            shape = list(post.val_shape)
            shape[d.val_dim] = t.part
            return Struct(post.out, replace(d, size=t.part), tuple(shape), ())
        if d.val_dim is None:
            raise NotImplementedError("Finish the implementation!")                               # :1186
        return Struct(Dim(DENSE, post.out.size, 0), Dim(DENSE, t.part, 1), (post.out.size, t.part), ())
    raise ValueError(t.kind)


def materialise(s: Struct) -> Struct:
    """Flush pending pre_transforms (reference core.py:238-250, 541-552)."""
    pending = s.pre
    s = replace(s, pre=())
    for t in pending[::-1]:
        s = apply_inverse_transform(t, s)
    return s
