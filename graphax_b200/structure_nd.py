"""Tensor-level structure of edge Jacobians for vertices of any rank (no values).

``structure.py`` covers graphs whose vertices have rank <= 1; this module is the same bookkeeping for matrix-valued
vertices: which dimensions of an edge Jacobian ``J[out dims | primal dims]`` are dense, which are Kronecker pairs
(one out dimension tied to one primal dimension), and which of them the compact value array has an axis for
(reference ``sparse/tensor.py:24-85``).  The engine computes with scalarised entries; the structure is what
``count_ops`` is defined on (``get_num_muls`` / ``get_num_adds``, ``tensor.py:1414-1493``), what
``sparse_representation=True`` reports (``core.py:555-581``) and what selects the product class
(``_is_pure_dot_product_mul`` / ``_is_pure_broadcast_mul`` / ``_mixed_mul``, ``tensor.py:302-371``).

What is restated (structure only, every rule cites the reference code it follows):

* elemental edges: ``make_parallel_jacobian`` (``primitives.py:42-107``), ``reduce_sum / max / min`` (``:243-346``),
  ``dot_general`` with batch dimensions (``:349-441``);
* deferred shape-op Jacobians and how they act on a neighbouring edge: ``transpose``, ``reshape``, ``slice``,
  ``broadcast_in_dim``, ``squeeze``, ``concatenate``, ``convert_element_type`` (``:606-1255``);
* products: scalar, all-dense contraction (``tensor.py:999-1059``), Kronecker relabelling
  (``_get_output_tensor``, ``:764-823``) and the mixed case (``:1062-1231``);
* sums: ``_sparse_add`` (``:1267-1411``).

A dimension is ``NDim(kind, size, ext, other)``: ``ext`` is the extent of the value array along the dimension's axis
(``None`` = no axis: a replicated dense dimension or a pure Kronecker delta; 1 is legal for a sparse pair of larger
size - the reference allows broadcasting there), ``other`` the position of a sparse dimension's partner in the
opposite list.  The numbering of value axes (``val_dim``) is not tracked while eliminating - no count depends on it -
and is re-created in the reference's canonical order (``_swap_back_axes``, ``tensor.py:710-761``) by ``descriptor``.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import List, Optional, Sequence, Tuple

DENSE, SPARSE = "D", "S"


@dataclass(frozen=True)
class NDim:
    kind: str
    size: int
    ext: Optional[int]
    other: Optional[int] = None       # sparse only: index of the partner in the opposite dimension list


@dataclass(frozen=True)
class NTransform:
    """A deferred shape-op Jacobian (``JacobianTransform``, ``primitives.py:469-489``)."""
    kind: str                          # transpose | reshape | slice | broadcast | squeeze | concat | convert
    in_shape: Tuple[int, ...] = ()
    out_shape: Tuple[int, ...] = ()
    perm: Tuple[int, ...] = ()         # transpose: permutation
    dims: Tuple[int, ...] = ()         # broadcast: broadcast_dimensions; squeeze: dimensions; concat: (dimension,)
    lo: int = 0                        # concat: [lo, hi) of this operand along the concatenated axis
    hi: int = 0


@dataclass(frozen=True)
class NStruct:
    out: Tuple[NDim, ...]
    primal: Tuple[NDim, ...]
    has_val: bool = True
    pre: Tuple[NTransform, ...] = ()

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(d.size for d in self.out + self.primal)

    @property
    def val_size(self) -> int:
        n = 1
        for e in self.val_shape():
            n *= e
        return n

    def _axes(self) -> List[Tuple[str, int]]:
        """Dimensions that own a value axis, in the reference's canonical order (dense dimensions and the out half
        of sparse pairs in order of appearance, ``_swap_back_axes``)."""
        axes = [("out", i) for i, d in enumerate(self.out) if d.ext is not None]
        axes += [("primal", j) for j, d in enumerate(self.primal) if d.ext is not None and d.kind == DENSE]
        return axes

    def val_shape(self) -> Tuple[int, ...]:
        return tuple((self.out[i] if side == "out" else self.primal[i]).ext for side, i in self._axes())

    def descriptor(self) -> dict:
        """Reference-compatible description: (kind, id, size, val_dim, other_id) per dimension."""
        l = len(self.out)
        axis = {a: n for n, a in enumerate(self._axes())}

        def enc(side: str, i: int, d: NDim):
            if d.ext is None:
                vd = None
            elif side == "primal" and d.kind == SPARSE:
                vd = axis[("out", d.other)]
            else:
                vd = axis[(side, i)]
            other = None if d.kind == DENSE else (l + d.other if side == "out" else d.other)
            return [d.kind, i if side == "out" else l + i, d.size, vd, other]
        return {"out_dims": [enc("out", i, d) for i, d in enumerate(self.out)],
                "primal_dims": [enc("primal", j, d) for j, d in enumerate(self.primal)],
                "val_shape": list(self.val_shape()) if self.has_val else None,
                "pre_transforms": [[t.kind] for t in self.pre]}

    def sparsity_class(self) -> str:
        if not self.has_val:
            return "transform"
        if not self.out and not self.primal:
            return "scalar"
        kinds = {d.kind for d in self.out + self.primal}
        if kinds == {SPARSE}:
            return "diag" if any(d.ext is not None for d in self.out) else "kron"
        return "dense" if kinds == {DENSE} else "mixed"


SCALAR = NStruct((), ())


def transform_edge(t: NTransform) -> NStruct:
    return NStruct((), (), False, (t,))


def strip(s: NStruct) -> NStruct:
    return replace(s, pre=()) if s.pre else s


def _check(s: NStruct) -> NStruct:
    """Pairing consistency (the part of ``_assert_sparse_tensor_consistency`` that survives without values)."""
    for i, d in enumerate(s.out):
        if d.kind == SPARSE:
            p = s.primal[d.other]
            assert p.kind == SPARSE and p.other == i and p.size == d.size and p.ext == d.ext, f"{s} is not self-consistent!"
    for j, d in enumerate(s.primal):
        if d.kind == SPARSE:
            assert s.out[d.other].kind == SPARSE and s.out[d.other].other == j, f"{s} is not self-consistent!"
    return s


# --------------------------------------------------------------------------------------------------
# rank <= 1 descriptors (structure.py) -> this representation
# --------------------------------------------------------------------------------------------------
def from_struct(s) -> NStruct:
    if isinstance(s, NStruct):
        return s

    def dim(d):
        ext = None if (d.val_dim is None or s.val_shape is None) else s.val_shape[d.val_dim]
        return NDim(d.kind, d.size, ext, 0 if d.kind == SPARSE else None)
    out = () if s.out is None else (dim(s.out),)
    primal = () if s.primal is None else (dim(s.primal),)
    pre = []
    for t in s.pre:
        if t.kind == "slice":
            pre.append(NTransform("slice", in_shape=(t.full,), out_shape=(t.hi - t.lo,)))
        else:
            pre.append(NTransform("concat", in_shape=(t.hi - t.lo,), out_shape=(t.full,), dims=(0,), lo=t.lo, hi=t.hi))
    return NStruct(out, primal, s.has_val, tuple(pre))


# --------------------------------------------------------------------------------------------------
# elemental edges
# --------------------------------------------------------------------------------------------------
def elementwise(primal_shape: Sequence[int], out_shape: Sequence[int], n_operands: int,
                partial_shape: Sequence[int], partial_is_scalar: bool) -> NStruct:
    """``make_parallel_jacobian`` (primitives.py:42-107).  ``partial_shape`` is the shape of the elemental partial
    (the value array), ``partial_is_scalar`` whether it is a Python float or has a single element."""
    ps, os_ = tuple(primal_shape), tuple(out_shape)
    if len(ps) == 0 and len(os_) == 0:
        return SCALAR
    if len(ps) == 0:
        return NStruct(tuple(NDim(DENSE, e, e) for e in os_), ())
    if n_operands == 1:
        return _check(NStruct(tuple(NDim(SPARSE, e, e, i) for i, e in enumerate(os_)),
                              tuple(NDim(SPARSE, e, e, i) for i, e in enumerate(os_))))
    if ps != os_:
        out, primal = [], []
        for i, (o, p) in enumerate(zip(os_, ps)):
            if p != o:
                out.append(NDim(DENSE, o, o))
                primal.append(NDim(DENSE, p, None))
            else:
                out.append(NDim(SPARSE, o, o, i))
                primal.append(NDim(SPARSE, o, o, i))
        return _check(NStruct(tuple(out), tuple(primal)))
    if partial_is_scalar:
        return _check(NStruct(tuple(NDim(SPARSE, e, None, i) for i, e in enumerate(ps)),
                              tuple(NDim(SPARSE, e, None, i) for i, e in enumerate(ps))))
    exts = tuple(partial_shape) if len(partial_shape) == len(ps) else ps
    return _check(NStruct(tuple(NDim(SPARSE, e, x, i) for i, (e, x) in enumerate(zip(ps, exts))),
                          tuple(NDim(SPARSE, e, x, i) for i, (e, x) in enumerate(zip(ps, exts)))))


def reduce_sum(in_shape: Sequence[int], axes: Sequence[int]) -> NStruct:
    """primitives.py:243-272: ones over the reduced axes, pure Kronecker pairs for the kept ones."""
    out, primal = [], []
    for i, size in enumerate(in_shape):
        if i in axes:
            primal.append(NDim(DENSE, size, size))
        else:
            out.append(NDim(SPARSE, size, None, len(primal)))
            primal.append(NDim(SPARSE, size, None, len(out) - 1))
    return _check(NStruct(tuple(out), tuple(primal)))


def reduce_extreme(in_shape: Sequence[int], axes: Sequence[int]) -> NStruct:
    """primitives.py:275-346: the normalised argmax / argmin mask has the operand's shape."""
    out, primal = [], []
    for i, size in enumerate(in_shape):
        if i in axes:
            primal.append(NDim(DENSE, size, size))
        else:
            out.append(NDim(SPARSE, size, size, len(primal)))
            primal.append(NDim(SPARSE, size, size, len(out) - 1))
    return _check(NStruct(tuple(out), tuple(primal)))


def dot_general(lhs_shape: Sequence[int], rhs_shape: Sequence[int], dimension_numbers) -> Tuple[NStruct, NStruct]:
    """primitives.py:349-441: the edge of each operand carries the other operand as its value.  Result dimensions
    are ordered (batch, lhs free, rhs free); contracted dimensions are dense, batch dimensions sparse pairs with a
    value axis, an operand's own free dimensions pure Kronecker pairs, the other operand's free dimensions dense."""
    (lc, rc), (lb, rb) = dimension_numbers
    lc, rc, lb, rb = list(lc), list(rc), list(lb), list(rb)

    def edge(own_shape, own_c, own_b, oth_shape, oth_c, oth_b, own_is_lhs: bool) -> NStruct:
        n_batch = len(own_b)
        own_free = [d for d in range(len(own_shape)) if d not in own_c and d not in own_b]
        oth_free = [d for d in range(len(oth_shape)) if d not in oth_c and d not in oth_b]
        # result dimension order: batch, lhs free, rhs free
        out: List[Optional[NDim]] = [None] * (n_batch + len(own_free) + len(oth_free))
        own_free_pos = {d: n_batch + (k if own_is_lhs else len(oth_free) + k) for k, d in enumerate(own_free)}
        oth_free_pos = {d: n_batch + (len(own_free) + k if own_is_lhs else k) for k, d in enumerate(oth_free)}
        primal: List[NDim] = []
        for d, size in enumerate(own_shape):
            if d in own_c:
                other_dim = oth_c[own_c.index(d)]
                primal.append(NDim(DENSE, oth_shape[other_dim], oth_shape[other_dim]))
            elif d in own_b:
                pos = own_b.index(d)
                out[pos] = NDim(SPARSE, size, size, d)
                primal.append(NDim(SPARSE, size, size, pos))
            else:
                out[own_free_pos[d]] = NDim(SPARSE, size, None, d)
                primal.append(NDim(SPARSE, size, None, own_free_pos[d]))
        for d in oth_free:
            out[oth_free_pos[d]] = NDim(DENSE, oth_shape[d], oth_shape[d])
        return _check(NStruct(tuple(out), tuple(primal)))
    return (edge(lhs_shape, lc, lb, rhs_shape, rc, rb, True), edge(rhs_shape, rc, rb, lhs_shape, lc, lb, False))


# --------------------------------------------------------------------------------------------------
# products (tensor.py:334-371)
# --------------------------------------------------------------------------------------------------
def _combine(a: Optional[int], b: Optional[int]) -> Optional[int]:
    if a is None:
        return b
    if b is None:
        return a
    return max(a, b)


def mul_class(lhs: NStruct, rhs: NStruct) -> str:
    assert tuple(d.size for d in lhs.primal) == tuple(d.size for d in rhs.out), \
        f"{lhs.shape} and {rhs.shape} not compatible for multiplication!"
    if lhs.shape == () and rhs.shape == ():
        return "scalar"
    pairs = list(zip(lhs.primal, rhs.out))
    if all(a.kind == DENSE and b.kind == DENSE for a, b in pairs):
        return "dot"
    if all(a.kind == SPARSE or b.kind == SPARSE for a, b in pairs):
        return "bcast"
    return "mixed"


def mul(lhs: NStruct, rhs: NStruct) -> NStruct:
    """Structure of ``lhs * rhs`` (lhs = edge out of the eliminated vertex, rhs = edge into it).  In all three
    product kinds the dimensions follow the same Kronecker algebra: a dense-dense pair is contracted away, a pair
    with a sparse side relabels (``_get_output_tensor`` tensor.py:764-823, ``_mixed_mul`` :1062-1231): the partner of
    the sparse side takes the kind of the other side, and it has a value axis when either factor had one."""
    cls = mul_class(lhs, rhs)
    if cls == "scalar":
        return SCALAR
    if cls == "dot":
        # tensor.py:999-1059: lhs.out_dims are kept as they are, rhs.primal_dims become dense
        return NStruct(lhs.out, tuple(NDim(DENSE, d.size, d.ext) for d in rhs.primal))
    n_out, n_primal = len(lhs.out), len(rhs.primal)
    out: List[Optional[NDim]] = [None] * n_out
    primal: List[Optional[NDim]] = [None] * n_primal
    for a, d in enumerate(lhs.out):
        if d.kind == DENSE:
            out[a] = d
    for m, d in enumerate(rhs.primal):
        if d.kind == DENSE:
            primal[m] = d
    for k, (ld, rd) in enumerate(zip(lhs.primal, rhs.out)):
        if ld.kind == DENSE and rd.kind == DENSE:
            continue                                            # contracted
        ext = _combine(ld.ext, rd.ext)
        if ld.kind == SPARSE and rd.kind == DENSE:
            out[ld.other] = NDim(DENSE, ld.size, ext)
        elif ld.kind == DENSE and rd.kind == SPARSE:
            primal[rd.other] = NDim(DENSE, ld.size, ext)
        else:
            out[ld.other] = NDim(SPARSE, ld.size, ext, rd.other)
            primal[rd.other] = NDim(SPARSE, ld.size, ext, ld.other)
    assert all(d is not None for d in out) and all(d is not None for d in primal)
    return _check(NStruct(tuple(out), tuple(primal)))


def num_muls(lhs: NStruct, rhs: NStruct) -> int:
    """``get_num_muls`` (tensor.py:1414-1456)."""
    n = 1
    for d in lhs.out:
        if d.kind == DENSE and d.ext is not None:
            n *= d.size
    for ld, rd in zip(lhs.primal, rhs.out):
        if ld.kind == DENSE:
            n *= ld.size
        elif rd.kind == DENSE:
            n *= rd.size
        elif ld.ext is not None and rd.ext is not None:
            n *= max(ld.ext, rd.ext)
        elif ld.ext is not None:
            n *= ld.ext
        elif rd.ext is not None:
            n *= rd.ext
        elif lhs.has_val and rhs.has_val:
            ls, rs = lhs.val_size, rhs.val_size
            if ls == 1 and rs == 1:
                n *= 1
            elif ls == 1:
                n *= rd.size
            elif rs == 1:
                n *= ld.size
    for d in rhs.primal:
        if d.kind == DENSE and d.ext is not None:
            n *= d.size
    return n


# --------------------------------------------------------------------------------------------------
# sums (tensor.py:1267-1411)
# --------------------------------------------------------------------------------------------------
def add(lhs: NStruct, rhs: NStruct) -> NStruct:
    assert lhs.shape == rhs.shape, f"{lhs.shape} and {rhs.shape} not compatible for addition!"
    assert not lhs.pre and not rhs.pre, "pending transforms must be materialised before addition"
    out: List[NDim] = []
    for ld, rd in zip(lhs.out, rhs.out):
        if ld.kind == SPARSE and rd.kind == SPARSE and ld.other == rd.other:
            out.append(NDim(SPARSE, ld.size, max(ld.ext or 1, rd.ext or 1), ld.other))
        elif ld.kind == SPARSE or rd.kind == SPARSE:
            out.append(NDim(DENSE, ld.size, ld.size))            # a sparse side is materialised with an eye mask
        else:
            out.append(NDim(DENSE, ld.size, 1 if (ld.ext is None and rd.ext is None) else ld.size))
    primal: List[NDim] = []
    for ld, rd in zip(lhs.primal, rhs.primal):
        if ld.kind == SPARSE and rd.kind == SPARSE and ld.other == rd.other:
            primal.append(NDim(SPARSE, ld.size, out[ld.other].ext, ld.other))
        elif ld.kind == SPARSE or rd.kind == SPARSE:
            primal.append(NDim(DENSE, ld.size, ld.size))
        else:
            primal.append(NDim(DENSE, ld.size, 1 if (ld.ext is None and rd.ext is None) else ld.size))
    # an out dimension whose partner was densified on the primal side cannot stay sparse
    for i, d in enumerate(out):
        if d.kind == SPARSE and primal[d.other].kind != SPARSE:
            out[i] = NDim(DENSE, d.size, d.size)
    return _check(NStruct(tuple(out), tuple(primal)))


def num_adds(summed: NStruct, old: NStruct) -> int:
    """``get_num_adds`` (tensor.py:1460-1493), called with the summed edge and the old one (core.py:252-253)."""
    n = 1
    for ld, rd in zip(summed.out, old.out):
        if ld.kind == SPARSE and rd.kind == SPARSE:
            if ld.ext is not None and rd.ext is not None:
                n *= ld.size
        else:
            n *= ld.size
    for ld, rd in zip(summed.primal, old.primal):
        if not (ld.kind == SPARSE and rd.kind == SPARSE):
            n *= ld.size
    return n


# --------------------------------------------------------------------------------------------------
# transforms: apply = act on the OUT side of the incoming edge, apply_inverse = on the PRIMAL side of the outgoing one
# --------------------------------------------------------------------------------------------------
def _all_dense(out_sizes: Sequence[int], primal_sizes: Sequence[int]) -> NStruct:
    return NStruct(tuple(NDim(DENSE, s, s) for s in out_sizes), tuple(NDim(DENSE, s, s) for s in primal_sizes))


def _densify_pair(out: List[NDim], primal: List[NDim], i: int, j: int, out_dim: NDim, primal_dim: NDim) -> None:
    out[i], primal[j] = out_dim, primal_dim


def apply_transform(t: NTransform, pre: NStruct) -> NStruct:
    if not pre.has_val:
        raise AssertionError("transforms are only unloaded onto edges with values")
    out, primal = list(pre.out), list(pre.primal)
    if t.kind == "convert":
        return strip(pre)
    if t.kind == "transpose":               # primitives.py:606-647: permute the out dimensions, partners follow
        new_out = [out[p] for p in t.perm]
        for new_i, p in enumerate(t.perm):
            if out[p].kind == SPARSE:
                primal[out[p].other] = replace(primal[out[p].other], other=new_i)
        return _check(NStruct(tuple(new_out), tuple(primal)))
    if t.kind in ("reshape", "slice"):      # :650-760: densify, then reshape / slice the out axes
        return _all_dense(t.out_shape, [d.size for d in primal])
    if t.kind == "broadcast":               # :763-808: every result axis that is not an operand axis is a size-1 dense one
        insert = [i for i in range(len(t.out_shape)) if i not in t.dims]
        for dim in insert:
            out.insert(dim, NDim(DENSE, 1, 1))
            for j, d in enumerate(primal):
                if d.kind == SPARSE and d.other >= dim:
                    primal[j] = replace(d, other=d.other + 1)
        return _check(NStruct(tuple(out), tuple(primal)))
    if t.kind == "squeeze":                 # :886-941: drop out dimensions; a sparse partner becomes a replicated dense one
        for dim in sorted(t.dims, reverse=True):
            d = out[dim]
            if d.kind == SPARSE:
                primal[d.other] = NDim(DENSE, primal[d.other].size, None)
            del out[dim]
            for j, p in enumerate(primal):
                if p.kind == SPARSE and p.other > dim:
                    primal[j] = replace(p, other=p.other - 1)
        return _check(NStruct(tuple(out), tuple(primal)))
    if t.kind == "concat":                  # :988-1129: zero-pad the concatenated out axis; a sparse pair is materialised
        dim = t.dims[0]
        d = out[dim]                        # IndexError without out dims, like the reference
        full = t.out_shape[dim]
        if d.kind == DENSE:
            if d.ext is None:
                raise NotImplementedError("DenseDimension without `val_dim` not yet supported!")
            out[dim] = NDim(DENSE, full, full)
        else:
            out[dim] = NDim(DENSE, full, full)
            primal[d.other] = NDim(DENSE, d.size, d.size)
        return _check(NStruct(tuple(out), tuple(primal)))
    raise ValueError(t.kind)


def apply_inverse_transform(t: NTransform, post: NStruct) -> NStruct:
    if not post.has_val:
        raise AssertionError("transforms are only unloaded onto edges with values")
    out, primal = list(post.out), list(post.primal)
    if t.kind == "convert":
        return strip(post)
    if t.kind == "transpose":               # primitives.py:628-645 (the reference iterates the inverse permutation)
        inv = [0] * len(t.perm)
        for i, p in enumerate(t.perm):
            inv[p] = i
        new_primal = [primal[p] for p in inv]
        for new_j, p in enumerate(inv):
            if primal[p].kind == SPARSE:
                out[primal[p].other] = replace(out[primal[p].other], other=new_j)
        return _check(NStruct(tuple(out), tuple(new_primal)))
    if t.kind in ("reshape", "slice"):      # :676-693, :727-755: densify, then reshape / scatter the primal axes
        return _all_dense([d.size for d in out], t.in_shape)
    if t.kind == "broadcast":               # :810-877: the inserted axes are summed away on the primal side
        rm = [d for d in range(len(t.out_shape)) if d not in t.dims]
        for n, dim in enumerate(rm):
            j = dim - n
            d = primal[j]
            if d.kind == SPARSE:
                out[d.other] = NDim(DENSE, out[d.other].size, None)
            del primal[j]
            for i, o in enumerate(out):
                if o.kind == SPARSE and o.other > j:
                    out[i] = replace(o, other=o.other - 1)
        return _check(NStruct(tuple(out), tuple(primal)))
    if t.kind == "squeeze":                 # :943-962: the squeezed axes come back as size-1 dense primal dimensions
        for dim in sorted(t.dims):
            primal.insert(dim, NDim(DENSE, 1, 1))
            for i, o in enumerate(out):
                if o.kind == SPARSE and o.other >= dim:
                    out[i] = replace(o, other=o.other + 1)
        return _check(NStruct(tuple(out), tuple(primal)))
    if t.kind == "concat":                  # :1132-1229: keep this operand's slice of the primal axis
        dim = t.dims[0]
        if not primal:
            raise AttributeError("'NoneType' object has no attribute 'other_id'")
        d = primal[dim]
        part = t.hi - t.lo
        if d.kind == DENSE:
            if d.ext is None:
                raise NotImplementedError("DenseDimension without `val_dim` not yet supported!")
            primal[dim] = NDim(DENSE, part, part)
        else:
            if d.ext is None:
                raise NotImplementedError("Finish the implementation!")
            o = out[d.other]
            out[d.other] = NDim(DENSE, o.size, o.size)
            primal[dim] = NDim(DENSE, part, part)
        return _check(NStruct(tuple(out), tuple(primal)))
    raise ValueError(t.kind)


def materialise(s: NStruct) -> NStruct:
    """Flush pending pre_transforms (core.py:238-250, 541-552)."""
    pending = s.pre
    s = strip(s)
    for t in pending[::-1]:
        s = apply_inverse_transform(t, s)
    return s
