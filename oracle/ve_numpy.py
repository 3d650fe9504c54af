"""ORACLE - test infrastructure, not product code.

CPU restatement (NumPy) of graphax's vertex-elimination path, vectorised over a
leading batch axis, i.e. the semantics of ``jax.vmap(graphax.jacve(f, order,
argnums))``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import this module; nothing under
``graphax_b200/`` does.

PINNED AGAINST A RUN OF THE REFERENCE'S OWN CODE (tests/test_reference_goldens.py): graphax
needs JAX, which is not in this image, so the unmodified reference sources were run here on
a NumPy stand-in for the JAX API (``tests/golden/jaxlite``, generator
``tests/golden/make_reference_goldens.py``; that pair reproduces the reference's literal
known answers, count_ops 632 / 566 for ``g``, ``examples/randoms.py:91``).  Against the
committed vectors of that run this module gives identical edge / block structures,
elimination orders and count_ops numbers, and float32 Jacobians that are equal BIT FOR BIT
for all 21 (function, order) pairs.  XLA's own arithmetic (fusion, FMA contraction) remains
unpinned: no JAX to run.  Further pins (tests/test_oracle.py): the vertex sets implied by
every elimination order in the reference's tests, the einsum identities of
``tests/core/*_mul_test.py`` (rank <= 1 cases), fp64 Jacobians from an independent autodiff.

What follows the reference, function by function (file:line into /root/reference/src/graphax):

* ``SparseTensor`` / dimensions ............ sparse/tensor.py:24-85
* ``dense`` ................................ sparse/tensor.py:123-171
* ``_mul`` dispatch + the three products ... sparse/tensor.py:334-371, 826-866, 999-1059
* ``_add`` / ``_sparse_add`` ............... sparse/tensor.py:374-397, 1267-1411
* ``get_num_muls`` / ``get_num_adds`` ...... sparse/tensor.py:1414-1493
* elemental rules .......................... primitives.py:42-107, 150-205, 243-272, 349-441
* slice / concatenate transforms ........... primitives.py:699-760, 971-1233
* ``_build_graph`` ......................... core.py:314-412
* ``_checkify_order`` ...................... core.py:275-311
* ``_eliminate_vertex`` .................... core.py:155-272
* flush + densify + regroup ................ core.py:541-568

Restrictions: vertices of rank <= 1 (every named workload of configs 1-4); values
carry a leading batch axis, so a per-sample ``val_dim`` k is array axis k+1.
Arithmetic is plain NumPy in the requested dtype: separate multiply and add, no
FMA contraction - this is the "reference-order, unfused" evaluation.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class DenseDimension:
    id: int
    size: int
    val_dim: Optional[int]


@dataclass
class SparseDimension:
    id: int
    size: int
    val_dim: Optional[int]
    other_id: int


def _is_s(d) -> bool:
    return isinstance(d, SparseDimension)


class JacobianTransform:
    def __init__(self, kind: str, lo: int, hi: int, full: int):
        self.kind, self.lo, self.hi, self.full = kind, lo, hi, full

    def apply(self, tensor: "SparseTensor") -> "SparseTensor":
        return (_slice_transform if self.kind == "slice" else _concatenate_transform)(self, tensor)

    def apply_inverse(self, tensor: "SparseTensor") -> "SparseTensor":
        return (_inverse_slice_transform if self.kind == "slice" else _inverse_concatenate_transform)(self, tensor)


class SparseTensor:
    """out_dims | primal_dims with a compact ``val`` of shape [B, *val_shape]."""

    def __init__(self, out_dims, primal_dims, val, pre_transforms=None):
        self.out_dims = tuple(out_dims)
        self.primal_dims = tuple(primal_dims)
        if len(self.out_dims) > 1 or len(self.primal_dims) > 1:
            raise NotImplementedError("oracle is restricted to vertices of rank <= 1")
        self.val = val
        self.pre_transforms = list(pre_transforms or [])
        self.shape = tuple(d.size for d in self.out_dims + self.primal_dims)
        _assert_consistency(self)

    @property
    def out(self):
        return self.out_dims[0] if self.out_dims else None

    @property
    def primal(self):
        return self.primal_dims[0] if self.primal_dims else None

    @property
    def val_shape(self):
        return None if self.val is None else tuple(self.val.shape[1:])

    def copy(self) -> "SparseTensor":
        return SparseTensor(copy.deepcopy(self.out_dims), copy.deepcopy(self.primal_dims), self.val,
                            self.pre_transforms)

    def descriptor(self) -> dict:
        def enc(d):
            return [("S" if _is_s(d) else "D"), d.id, d.size, d.val_dim, d.other_id if _is_s(d) else None]
        return {"out_dims": [enc(d) for d in self.out_dims],
                "primal_dims": [enc(d) for d in self.primal_dims],
                "val_shape": None if self.val is None else list(self.val_shape),
                "pre_transforms": [[t.kind, t.lo, t.hi, t.full] for t in self.pre_transforms]}

    # tensor.py:123-171
    def dense(self) -> np.ndarray:
        B = self.val.shape[0]
        o, p = self.out, self.primal
        if o is None and p is None:
            return self.val
        v = self.val
        if o is not None and p is not None and _is_s(o):
            n = o.size
            diag = v.reshape(B, -1) if o.val_dim is not None else v.reshape(B, 1)
            diag = np.broadcast_to(diag, (B, n))
            return diag[:, :, None] * np.eye(n, dtype=v.dtype)[None]
        # dense dims: insert replicated axes, then tile
        full, shape = [B], [B]
        for d in (o, p):
            if d is None:
                continue
            shape.append(d.size if d.val_dim is not None else 1)
            full.append(d.size)
        return np.broadcast_to(v.reshape(shape), full).copy()


def _assert_consistency(st: SparseTensor) -> None:
    """Subset of tensor.py:204-258 that is meaningful at rank <= 1."""
    l = len(st.out_dims)
    for i, d in enumerate(st.out_dims):
        assert d.id == i, f"{st.descriptor()} is not self-consistent!"
    for i, d in enumerate(st.primal_dims):
        assert d.id == l + i, f"{st.descriptor()} is not self-consistent!"
    for d in st.out_dims:
        if _is_s(d):
            other = st.primal_dims[d.other_id - l]
            assert _is_s(other) and other.other_id == d.id and other.size == d.size
    if st.val is not None:
        for d in st.out_dims + st.primal_dims:
            if d.val_dim is not None:
                ext = st.val.shape[d.val_dim + 1]
                assert ext == d.size or (_is_s(d) and ext == 1) or d.size == 1, \
                    f"{st.descriptor()} is not self-consistent!"


# --------------------------------------------------------------------------
# multiplication
# --------------------------------------------------------------------------
def _axis(st: SparseTensor, d) -> Optional[int]:
    return None if d is None or d.val_dim is None else d.val_dim + 1


def _mul(lhs: SparseTensor, rhs: SparseTensor) -> SparseTensor:
    """tensor.py:334-371; lhs = edge out of the eliminated vertex, rhs = edge into it."""
    l, r = len(lhs.out_dims), len(rhs.out_dims)
    assert lhs.shape[l:] == rhs.shape[:r], f"{lhs.shape} and {rhs.shape} not compatible for multiplication!"
    if lhs.shape == () and rhs.shape == ():
        return SparseTensor((), (), lhs.val * rhs.val)
    pairs = list(zip(lhs.primal_dims, rhs.out_dims))
    if all(not _is_s(a) and not _is_s(b) for a, b in pairs):
        return _pure_dot_product_mul(lhs.copy(), rhs.copy())
    if all(_is_s(a) or _is_s(b) for a, b in pairs):
        return _pure_broadcast_mul(lhs.copy(), rhs.copy())
    raise NotImplementedError("_mixed_mul needs vertices of rank >= 2")


def _pure_dot_product_mul(lhs: SparseTensor, rhs: SparseTensor) -> SparseTensor:
    """tensor.py:999-1059: replicate ``val_dim=None`` axes, then contract with dot_general."""
    B = lhs.val.shape[0]
    l = len(lhs.out_dims)
    lo, ld, rd, rp = lhs.out, lhs.primal, rhs.out, rhs.primal
    if lo is not None:
        assert not _is_s(lo) and lo.val_dim is not None
    L = lhs.val.reshape(B, lo.size if lo is not None else 1, -1)          # [B, o, v or 1]
    R = rhs.val
    if ld is not None:
        n = ld.size
        L = np.broadcast_to(L, (B, L.shape[1], n))                        # jnp.tile of a replicated axis
        rshape = [B, n if rd.val_dim is not None else 1, rp.size if (rp is not None and rp.val_dim is not None) else 1]
        R = np.broadcast_to(R.reshape(rshape), (B, n, rshape[2]))
    else:
        R = R.reshape(B, 1, -1)
    # out[b,o,p] = sum_v L[b,o,v] * R[b,v,p]: products first, then the sum in ascending v
    prod = L[:, :, :, None] * R[:, None, :, :]
    acc = prod[:, :, 0, :]
    for k in range(1, prod.shape[2]):
        acc = acc + prod[:, :, k, :]
    new_out = [DenseDimension(0, lo.size, 0)] if lo is not None else []
    new_primal, shape = [], [B] + ([lo.size] if lo is not None else [])
    if rp is not None:
        if rp.val_dim is not None:
            new_primal.append(DenseDimension(l, rp.size, l))
            shape.append(rp.size)
        else:
            new_primal.append(DenseDimension(l, rp.size, None))
    return SparseTensor(new_out, new_primal, acc.reshape(shape))


def _pure_broadcast_mul(lhs: SparseTensor, rhs: SparseTensor) -> SparseTensor:
    """tensor.py:826-866 with the padding of :570-707 and the dims of :764-823:
    S.S -> S pair, S.D or D.S -> dense (row / column scaling)."""
    B = lhs.val.shape[0]
    ld, rd = lhs.primal, rhs.out
    n = ld.size
    if _is_s(ld) and _is_s(rd):
        if ld.val_dim is None and rd.val_dim is None:
            val = lhs.val.reshape(B) * rhs.val.reshape(B)
            return SparseTensor([SparseDimension(0, n, None, 1)], [SparseDimension(1, n, None, 0)], val)
        val = lhs.val.reshape(B, -1) * rhs.val.reshape(B, -1)
        return SparseTensor([SparseDimension(0, n, 0, 1)], [SparseDimension(1, n, 0, 0)], val)
    if _is_s(ld):
        # diagonal (or scalar*I) times an edge that is dense along v: scale its rows
        assert rd.val_dim is not None, "replicated out dimension"
        rp = rhs.primal
        scale = lhs.val.reshape(B, -1)                                     # [B, n] or [B, 1]
        if rp is not None and rp.val_dim is not None:
            val = scale[:, :, None] * rhs.val
            return SparseTensor([DenseDimension(0, n, 0)], [DenseDimension(1, rp.size, 1)], val)
        val = scale * rhs.val.reshape(B, -1)
        new_primal = [DenseDimension(1, rp.size, None)] if rp is not None else []
        return SparseTensor([DenseDimension(0, n, 0)], new_primal, val)
    # edge dense along v times a diagonal (or scalar*I): scale its columns
    lo = lhs.out
    if lo is not None:
        assert not _is_s(lo) and lo.val_dim is not None
    l = len(lhs.out_dims)
    new_out = [DenseDimension(0, lo.size, 0)] if lo is not None else []
    if ld.val_dim is None and rd.val_dim is None:
        scale = rhs.val.reshape([B] + [1] * (lhs.val.ndim - 1))
        return SparseTensor(new_out, [DenseDimension(l, n, None)], lhs.val * scale)
    lead = [B] + ([lo.size] if lo is not None else [])
    Lv = lhs.val.reshape(lead + [-1]) if ld.val_dim is not None else lhs.val.reshape(lead + [1])
    Rv = rhs.val.reshape([B] + [1] * (len(lead) - 1) + [-1]) if rd.val_dim is not None \
        else rhs.val.reshape([B] + [1] * len(lead))
    return SparseTensor(new_out, [DenseDimension(l, n, l)], Lv * Rv)


def get_num_muls(lhs: SparseTensor, rhs: SparseTensor) -> int:
    """tensor.py:1414-1456 (per-sample shapes: the batch axis is not counted)."""
    n = 1
    for d in lhs.out_dims:
        if not _is_s(d) and d.val_dim is not None:
            n *= d.size
    for ld, rd in zip(lhs.primal_dims, rhs.out_dims):
        if not _is_s(ld):
            n *= ld.size
        elif not _is_s(rd):
            n *= rd.size
        elif ld.val_dim is not None and rd.val_dim is not None:
            n *= max(lhs.val_shape[ld.val_dim], rhs.val_shape[rd.val_dim])
        elif ld.val_dim is not None:
            n *= lhs.val_shape[ld.val_dim]
        elif rd.val_dim is not None:
            n *= rhs.val_shape[rd.val_dim]
        elif lhs.val is not None and rhs.val is not None:
            ls, rs = int(np.prod(lhs.val_shape)), int(np.prod(rhs.val_shape))
            if ls == 1 and rs == 1:
                n *= 1
            elif ls == 1:
                n *= rd.size
            elif rs == 1:
                n *= ld.size
    for d in rhs.primal_dims:
        if not _is_s(d) and d.val_dim is not None:
            n *= d.size
    return n


# --------------------------------------------------------------------------
# addition
# --------------------------------------------------------------------------
def _add(lhs: SparseTensor, rhs: SparseTensor) -> SparseTensor:
    """tensor.py:1267-1411: S+S with the same partner stays sparse, everything else
    is materialised (sparse side times an eye mask) and becomes dense."""
    assert lhs.shape == rhs.shape, f"{lhs.shape} and {rhs.shape} not compatible for addition!"
    B = lhs.val.shape[0]
    o, p = lhs.out, lhs.primal
    if o is None and p is None:
        return SparseTensor((), (), lhs.val + rhs.val)
    if o is not None and _is_s(o) and _is_s(rhs.out):
        n = o.size
        val = lhs.val.reshape(B, -1) + rhs.val.reshape(B, -1)
        return SparseTensor([SparseDimension(0, n, 0, 1)], [SparseDimension(1, n, 0, 0)], val)
    dims_l = [d for d in (lhs.out, lhs.primal) if d is not None]
    dims_r = [d for d in (rhs.out, rhs.primal) if d is not None]
    ext = [B]
    for dl, dr in zip(dims_l, dims_r):
        both_replicated = (not _is_s(dl) and not _is_s(dr) and dl.val_dim is None and dr.val_dim is None)
        ext.append(1 if both_replicated else dl.size)

    def materialise(st: SparseTensor) -> np.ndarray:
        full = st.dense()
        take = [slice(None)] + [slice(0, e) for e in ext[1:]]
        return full[tuple(take)]
    val = materialise(lhs) + materialise(rhs)
    new_out = [DenseDimension(0, o.size, 0)] if o is not None else []
    l = len(new_out)
    new_primal = [DenseDimension(l, p.size, l)] if p is not None else []
    return SparseTensor(new_out, new_primal, val)


def get_num_adds(lhs: SparseTensor, rhs: SparseTensor) -> int:
    """tensor.py:1460-1493."""
    n = 1
    for ld, rd in zip(lhs.out_dims, rhs.out_dims):
        if _is_s(ld) and _is_s(rd):
            if ld.val_dim is not None and rd.val_dim is not None:
                n *= ld.size
        else:
            n *= ld.size
    for ld, rd in zip(lhs.primal_dims, rhs.primal_dims):
        if not (_is_s(ld) and _is_s(rd)):
            n *= ld.size
    return n


# --------------------------------------------------------------------------
# transforms (shape-op Jacobians)
# --------------------------------------------------------------------------
def _slice_transform(t: JacobianTransform, pre: SparseTensor) -> SparseTensor:
    """primitives.py:706-725: densify the incoming edge and slice its out axis."""
    full = pre.dense()
    if pre.out is None:
        raise IndexError("slice of an edge without out dimension")
    new_val = full[:, t.lo:t.hi]
    new_out = [DenseDimension(0, t.hi - t.lo, 0)]
    new_primal = [DenseDimension(1, pre.primal.size, 1)] if pre.primal is not None else []
    return SparseTensor(new_out, new_primal, new_val)


def _inverse_slice_transform(t: JacobianTransform, post: SparseTensor) -> SparseTensor:
    """primitives.py:727-755: scatter the densified edge into zeros of the operand's shape."""
    full = post.dense()
    B = full.shape[0]
    if post.out is not None:
        zeros = np.zeros((B, post.out.size, t.full), dtype=full.dtype)
        zeros[:, :, t.lo:t.hi] = full
        return SparseTensor([DenseDimension(0, post.out.size, 0)], [DenseDimension(1, t.full, 1)], zeros)
    zeros = np.zeros((B, t.full), dtype=full.dtype)
    zeros[:, t.lo:t.hi] = full
    return SparseTensor([], [DenseDimension(0, t.full, 0)], zeros)


def _concatenate_transform(t: JacobianTransform, pre: SparseTensor) -> SparseTensor:
    """primitives.py:988-1129: zero-pad the out axis of the incoming edge to the
    concatenated size; sparse pairs are materialised first."""
    d = pre.out_dims[0]                       # IndexError without out dims, like the reference
    B = pre.val.shape[0]
    if not _is_s(d):
        if d.val_dim is None:
            raise NotImplementedError("DenseDimension without `val_dim` not yet supported!")
        shape = list(pre.val.shape)
        shape[d.val_dim + 1] = t.full
        new_val = np.zeros(shape, dtype=pre.val.dtype)
        idx = [slice(None)] * len(shape)
        idx[d.val_dim + 1] = slice(t.lo, t.hi)
        new_val[tuple(idx)] = pre.val
        new_out = [DenseDimension(0, t.full, d.val_dim)]
        return SparseTensor(new_out, copy.deepcopy(pre.primal_dims), new_val)
    block = pre.dense()                       # [B, n, n] = val * eye
    new_val = np.zeros((B, t.full, d.size), dtype=pre.val.dtype)
    new_val[:, t.lo:t.hi, :] = block
    return SparseTensor([DenseDimension(0, t.full, 0)], [DenseDimension(1, d.size, 1)], new_val)


def _inverse_concatenate_transform(t: JacobianTransform, post: SparseTensor) -> SparseTensor:
    """primitives.py:1132-1229: keep the slice of the primal axis that belongs to this operand."""
    d = post.primal
    if d is None:
        raise AttributeError("'NoneType' object has no attribute 'other_id'")
    if not _is_s(d):
        if d.val_dim is None:
            raise NotImplementedError("DenseDimension without `val_dim` not yet supported!")
        idx = [slice(None)] * post.val.ndim
        idx[d.val_dim + 1] = slice(t.lo, t.hi)
        new_primal = [DenseDimension(d.id, t.hi - t.lo, d.val_dim)]
        return SparseTensor(copy.deepcopy(post.out_dims), new_primal, post.val[tuple(idx)])
    if d.val_dim is None:
        raise NotImplementedError("Finish the implementation!")
    block = post.dense()
    new_val = block[:, :, t.lo:t.hi]
    return SparseTensor([DenseDimension(0, post.out.size, 0)], [DenseDimension(1, t.hi - t.lo, 1)], new_val)


# --------------------------------------------------------------------------
# elemental rules on batched arrays
# --------------------------------------------------------------------------
class _Lit:
    """A jaxpr literal (Python float in the reference)."""
    def __init__(self, v: float):
        self.v = v


def _ipow(x, y: int):
    if y == 0:
        return np.ones_like(x)
    if y < 0:
        return 1.0 / _ipow(x, -y)
    result, base = None, x
    while y:
        if y & 1:
            result = base if result is None else result * base
        y >>= 1
        if y:
            base = base * base
    return result


def _make_parallel_jacobian(primal, out, elemental, n_operands: int, per_sample_shape) -> SparseTensor:
    """primitives.py:42-107 (rank <= 1); all arrays carry the batch axis."""
    pshape, oshape = per_sample_shape(primal), per_sample_shape(out)
    B = out.shape[0]
    eshape = per_sample_shape(elemental)
    e = elemental if not isinstance(elemental, _Lit) else np.full((B,), elemental.v, dtype=out.dtype)
    if len(pshape) == 0 and len(oshape) == 0:
        return SparseTensor((), (), np.broadcast_to(e.reshape(B), (B,)))
    n = oshape[0]
    if len(pshape) == 0:
        return SparseTensor([DenseDimension(0, n, 0)], [], np.broadcast_to(e.reshape(B, -1), (B, n)))
    if n_operands == 1:
        return SparseTensor([SparseDimension(0, n, 0, 1)], [SparseDimension(1, n, 0, 0)], e.reshape(B, n))
    if pshape != oshape:
        return SparseTensor([DenseDimension(0, n, 0)], [DenseDimension(1, pshape[0], None)],
                            np.broadcast_to(e.reshape(B, -1), (B, n)))
    if isinstance(elemental, _Lit) or int(np.prod(eshape)) == 1:
        return SparseTensor([SparseDimension(0, n, None, 1)], [SparseDimension(1, n, None, 0)], e.reshape(B))
    return SparseTensor([SparseDimension(0, n, 0, 1)], [SparseDimension(1, n, 0, 0)], e.reshape(B, n))


_UNARY_RULES: Dict[str, Tuple[Callable, Callable]] = {
    # name: (primal fn, partial fn(out, x))         primitives.py:150-175
    "neg": (lambda x: -x, lambda out, x: -np.ones_like(x)),
    "abs": (np.abs, lambda out, x: x / out),
    "exp": (np.exp, lambda out, x: out),
    "log": (np.log, lambda out, x: 1. / x),
    "sqrt": (np.sqrt, lambda out, x: .5 / out),
    "logistic": (lambda x: 1. / (1. + np.exp(-x)), lambda out, x: out * (1. - out)),
    "log1p": (np.log1p, lambda out, x: 1. / (1. + x)),
    "sin": (np.sin, lambda out, x: np.cos(x)),
    "asin": (np.arcsin, lambda out, x: 1. / np.sqrt(1. - x * x)),
    "cos": (np.cos, lambda out, x: -np.sin(x)),
    "acos": (np.arccos, lambda out, x: -1. / np.sqrt(1. - x * x)),
    "tan": (np.tan, lambda out, x: 1. + out * out),
    "atan": (np.arctan, lambda out, x: 1. / (1. + x * x)),
    "sinh": (np.sinh, lambda out, x: np.cosh(x)),
    "asinh": (np.arcsinh, lambda out, x: 1. / np.sqrt(1. + x * x)),   # corrected, see SURVEY App. E
    "cosh": (np.cosh, lambda out, x: np.sinh(x)),
    "acosh": (np.arccosh, lambda out, x: 1. / np.sqrt(x * x - 1.)),
    "tanh": (np.tanh, lambda out, x: 1. - out * out),
    "atanh": (np.arctanh, lambda out, x: 1. / (1. - x * x)),
    "erf": (None, lambda out, x: 2. * np.exp(-(x * x)) / np.sqrt(np.pi)),
}


def _elemental(eqn, invals, dtype, batch: int):
    """Primal output and the SparseTensor partial of every traced operand."""
    from graphax_b200.tracer import Var      # the jaxpr is the *input* of the oracle, like jax.make_jaxpr's

    def shp(a):
        return () if isinstance(a, _Lit) else tuple(a.shape[1:])

    def arr(a):
        return np.asarray(a.v, dtype=dtype) if isinstance(a, _Lit) else a

    def bc(a, ref_ndim):
        """align a per-sample rank-0 batched array with rank-1 operands"""
        a = arr(a)
        if isinstance(a, np.ndarray) and a.ndim == 1 and ref_ndim == 1:
            return a[:, None]
        return a

    prim = eqn.primitive
    traced = [i for i, a in enumerate(eqn.invars) if isinstance(a, Var)]
    if prim in ("add", "sub", "mul", "div", "atan2", "pow", "max", "min"):
        X, Y = invals
        nd = max(len(shp(X)), len(shp(Y)))
        x, y = bc(X, nd), bc(Y, nd)
        if prim == "add":
            out, parts = x + y, (None, None)
        elif prim == "sub":
            out, parts = x - y, (None, None)
        elif prim == "mul":
            out, parts = x * y, (None, None)
        elif prim == "div":
            out = x / y
            parts = (1. / y, -x / (y * y))
        elif prim == "atan2":
            out = np.arctan2(x, y)
            abs2 = x * x + y * y
            parts = (y / abs2, -x / abs2)
        elif prim in ("max", "min"):       # primitives.py:208-215
            first = (x > y) if prim == "max" else (x < y)
            out = np.where(first, x, y) + 0 * (x + y)
            parts = (np.where(first, 1., 0.) + 0 * (x + y), np.where(first, 0., 1.) + 0 * (x + y))
        else:
            out = np.power(x, y)
            parts = (y * np.power(x, y - 1.), np.log(x) * out)
        out = np.asarray(out, dtype=dtype)
        if out.ndim == 0 or out.shape[0] != batch:
            out = np.broadcast_to(out, (batch,) + out.shape[1:] if out.ndim else (batch,))
        ones = lambda a: _Lit(1.0) if isinstance(a, _Lit) else np.ones_like(arr(a))
        if prim == "add":
            parts = (ones(Y), ones(X))
        elif prim == "sub":
            ox = ones(X)
            parts = (ones(Y), _Lit(-1.0) if isinstance(ox, _Lit) else -ox)
        elif prim == "mul":
            parts = (Y, X)
        parts = [np.asarray(p, dtype=dtype) if not isinstance(p, _Lit) else p for p in parts]
        parts = [p if isinstance(p, _Lit) or p.ndim >= 1 else _Lit(float(p)) for p in parts]
        # a partial keeps the batch axis only if it depends on a traced value
        fixed = []
        for p in parts:
            if not isinstance(p, _Lit) and p.shape[0] != batch:
                p = np.broadcast_to(p, (batch,) + p.shape[1:])
            fixed.append(p)
        edges = [_make_parallel_jacobian(invals[i], out, fixed[i], 2, shp) for i in traced]
        return out, edges
    if prim == "integer_pow":
        (x,) = invals
        y = int(eqn.params["y"])
        out = _ipow(x, y)
        part = np.asarray(y, dtype=dtype) * _ipow(x, y - 1)
        return out, [_make_parallel_jacobian(x, out, part, 1, shp)]
    if prim == "square":
        (x,) = invals
        return x * x, [_make_parallel_jacobian(x, x, np.asarray(2., dtype=dtype) * x, 1, shp)]
    if prim in _UNARY_RULES:
        (x,) = invals
        f, df = _UNARY_RULES[prim]
        if f is None:
            import math
            f = np.vectorize(math.erf, otypes=[np.float64])
        out = f(x).astype(dtype)
        return out, [_make_parallel_jacobian(x, out, df(out, x).astype(dtype), 1, shp)]
    if prim == "reduce_sum":            # primitives.py:243-272
        (x,) = invals
        acc = x[:, 0]
        for k in range(1, x.shape[1]):
            acc = acc + x[:, k]
        n = x.shape[1]
        edge = SparseTensor([], [DenseDimension(0, n, 0)], np.ones((batch, n), dtype=np.float32).astype(dtype))
        return acc, [edge]
    if prim == "dot_general":           # primitives.py:349-441
        x, y = invals
        prod = x * y
        acc = prod[:, 0]
        for k in range(1, prod.shape[1]):
            acc = acc + prod[:, k]
        n = x.shape[1]
        edges = []
        for i in traced:
            edges.append(SparseTensor([], [DenseDimension(0, n, 0)], (y, x)[i]))
        return acc, edges
    if prim == "slice":                 # primitives.py:699-760
        (x,) = invals
        (lo,), (hi,) = eqn.params["start_indices"], eqn.params["limit_indices"]
        t = JacobianTransform("slice", lo, hi, x.shape[1])
        return x[:, lo:hi], [SparseTensor([], [], None, [t])]
    if prim == "concatenate":           # primitives.py:971-1233
        vals = [np.broadcast_to(arr(v), (batch,) + arr(v).shape[1:]) for v in invals]
        out = np.concatenate(vals, axis=1)
        edges, off = [], 0
        for a, v in zip(eqn.invars, vals):
            if isinstance(a, Var):
                edges.append(SparseTensor([], [], None, [JacobianTransform("concat", off, off + v.shape[1], out.shape[1])]))
            off += v.shape[1]
        return out, edges
    if prim == "broadcast_in_dim":      # literal operand only (jnp.zeros): no edge
        (x,) = invals
        assert isinstance(x, _Lit)
        shape = tuple(eqn.params["shape"])
        return np.full((batch,) + shape, x.v, dtype=dtype), []
    raise NotImplementedError(f"{prim} does not have registered elemental partial.")


# --------------------------------------------------------------------------
# graph + elimination
# --------------------------------------------------------------------------
def _key(var) -> int:
    return -(var.input_index + 1) if var.input_index is not None else var.eqn_index + 1


def build_graph(jaxpr, args: Sequence[np.ndarray], dtype):
    """core.py:314-412."""
    from graphax_b200.tracer import Literal, Var
    batch = args[0].shape[0]
    env: Dict[int, np.ndarray] = {}
    graph: Dict[int, Dict[int, SparseTensor]] = {}
    transpose_graph: Dict[int, Dict[int, SparseTensor]] = {}
    vo_vertices = set()
    for var, a in zip(jaxpr.invars, args):
        a = np.asarray(a, dtype=dtype)
        assert a.shape[1:] == var.shape or (var.shape == () and a.ndim == 1), (a.shape, var.shape)
        env[_key(var)] = a
        graph[_key(var)] = {}
        transpose_graph[_key(var)] = {}
    # closed-over arrays (core.py:374): in the environment like the arguments, the same value for every sample
    for var, c in zip(getattr(jaxpr, "constvars", ()), getattr(jaxpr, "consts", ())):
        c = np.asarray(c, dtype=dtype)
        env[_key(var)] = np.ascontiguousarray(np.broadcast_to(c, (batch,) + c.shape))
        graph[_key(var)] = {}
        transpose_graph[_key(var)] = {}
    outkeys = {_key(v) for v in jaxpr.outvars}
    elemental = []
    for eqn in jaxpr.eqns:
        v = _key(eqn.outvar)
        graph.setdefault(v, {})
        transpose_graph.setdefault(v, {})
        for a in eqn.invars:
            if isinstance(a, Var) and a.eqn_index is not None and _key(a) in outkeys:
                vo_vertices.add(_key(a))
        invals = [(_Lit(a.val) if isinstance(a, Literal) else env[_key(a)]) for a in eqn.invars]
        out, partials = _elemental(eqn, invals, dtype, batch)
        env[v] = np.ascontiguousarray(out, dtype=dtype)
        invars = [a for a in eqn.invars if isinstance(a, Var)]
        if len(invars) != len(partials):
            continue
        for a, val in zip(invars, partials):
            src = _key(a)
            if v in graph[src]:
                val = _add(val, graph[src][v])       # x*x: accumulate instead of overwrite (App. E)
            graph[src][v] = val
            transpose_graph[v][src] = val
            elemental.append([src, v, val.descriptor()])
    return env, graph, transpose_graph, vo_vertices, elemental


def checkify_order(order, jaxpr, vo_vertices) -> List[int]:
    """core.py:275-311."""
    outkeys = {_key(v) for v in jaxpr.outvars}
    valid = [i for i in range(1, len(jaxpr.eqns) + 1) if i not in outkeys or i in vo_vertices]
    if isinstance(order, str):
        if order in ("forward", "fwd"):
            return valid
        if order in ("reverse", "rev"):
            return valid[::-1]
        raise ValueError(f"{order} is not a valid order identifier!")
    missing = set(valid).difference(set(order))
    if missing:
        raise ValueError(f"Supplied order is missing vertices {missing}!")
    return list(order)


def _flush(tensor: SparseTensor) -> SparseTensor:
    pending = tensor.pre_transforms
    tensor = SparseTensor(tensor.out_dims, tensor.primal_dims, tensor.val)
    for t in pending[::-1]:
        tensor = t.apply_inverse(tensor)
    return tensor


def eliminate_vertex(vertex: int, graph, transpose_graph, vo_vertices) -> Tuple[int, int]:
    """core.py:155-272."""
    num_mul = num_add = 0
    for out_edge in list(graph[vertex].keys()):
        post_val = graph[vertex][out_edge]
        for in_edge in list(transpose_graph[vertex].keys()):
            pre_val = transpose_graph[vertex][in_edge]
            _pre_val, _post_val = pre_val, post_val
            if post_val.pre_transforms and pre_val.val is not None:
                _pre_val = SparseTensor(pre_val.out_dims, pre_val.primal_dims, pre_val.val)
                for t in post_val.pre_transforms:
                    _pre_val = t.apply(_pre_val)
            if pre_val.val is not None and post_val.val is not None:
                lhs = SparseTensor(_post_val.out_dims, _post_val.primal_dims, _post_val.val)
                rhs = SparseTensor(_pre_val.out_dims, _pre_val.primal_dims, _pre_val.val)
                edge_outval = _mul(lhs, rhs)
                num_mul += get_num_muls(lhs, rhs)
                pending = list(pre_val.pre_transforms)
            elif pre_val.val is not None:
                edge_outval = SparseTensor(_pre_val.out_dims, _pre_val.primal_dims, _pre_val.val)
                pending = list(pre_val.pre_transforms)
            else:
                edge_outval = SparseTensor(_post_val.out_dims, _post_val.primal_dims, _post_val.val)
                pending = list(pre_val.pre_transforms) + list(post_val.pre_transforms)
            edge_outval.pre_transforms = pending

            if graph[in_edge].get(out_edge) is not None:
                _edge = _flush(transpose_graph[out_edge][in_edge])
                edge_outval = _flush(edge_outval)
                edge_outval = _add(edge_outval, _edge)
                num_add += get_num_adds(edge_outval, _edge)
            graph[in_edge][out_edge] = edge_outval
            transpose_graph[out_edge][in_edge] = edge_outval
    if vertex not in vo_vertices:
        for in_vertex in transpose_graph[vertex]:
            del graph[in_vertex][vertex]
    for out_vertex in graph[vertex]:
        del transpose_graph[out_vertex][vertex]
    del graph[vertex]
    if vertex not in vo_vertices:
        del transpose_graph[vertex]
    return num_mul, num_add


def jacve(jaxpr, order, argnums: Sequence[int], args: Sequence[np.ndarray], dtype=np.float32):
    """``vmap(jacve(f, order, argnums, count_ops=True))`` on a traced jaxpr.

    Returns ``(jac, primal, aux)`` with ``jac[o][a]`` of shape
    ``[B, *out_o.shape, *arg_a.shape]`` (core.py:555-568), ``primal[o]`` the
    outputs and ``aux`` the count_ops dictionary plus structure descriptors.
    """
    with np.errstate(all="ignore"):
        env, graph, tgraph, vo, elemental = build_graph(jaxpr, args, dtype)
        vertices = checkify_order(order, jaxpr, vo)
        counts = [eliminate_vertex(v, graph, tgraph, vo) for v in vertices]
        batch = np.asarray(args[0]).shape[0]
        jac, blocks = [], {}
        for oi, outvar in enumerate(jaxpr.outvars):
            row = []
            for ai, a in enumerate(argnums):
                invar = jaxpr.invars[a]
                edge = graph.get(_key(invar), {}).get(_key(outvar))
                if edge is None:
                    row.append(np.zeros((batch,) + outvar.shape + invar.shape, dtype=dtype))
                    blocks[f"{oi},{ai}"] = None
                else:
                    edge = _flush(edge)
                    blocks[f"{oi},{ai}"] = edge.descriptor()
                    row.append(edge.dense().reshape((batch,) + outvar.shape + invar.shape))
            jac.append(row)
    aux = {"num_muls": sum(c[0] for c in counts), "num_adds": sum(c[1] for c in counts),
           "order_counts": [(int(v), int(c[0])) for v, c in zip(vertices, counts)],
           "structure": {"elemental": elemental, "blocks": blocks, "order": [int(v) for v in vertices]}}
    primal = [env[_key(o)] for o in jaxpr.outvars]
    return jac, primal, aux
