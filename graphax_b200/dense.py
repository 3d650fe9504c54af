"""Tensor-valued vertices: reverse-order elimination with dense contractions on tcgen05 (scope row (f)1).

For graphs whose vertices are matrices (config 5: an MLP over a batch, ``argnums`` = the weights) the
per-sample engine does not apply: the Jacobian blocks are [H, I] weight gradients and the products
``J[s,p] += J[s,v] . J[v,p]`` are real tensor contractions.  This module implements the part of that
space the named config needs - a scalar output eliminated in reverse order - at tensor level:

* every product has the dense cotangent edge ``() | vertex-shape`` on its left (reference: ``out_dims = ()``,
  all primal dims dense);
* times an elementwise (diagonal / Kronecker) edge it is a broadcast multiply
  (``_pure_broadcast_mul``, ``sparse/tensor.py:826-866``) -> fused elementwise kernels, compiled per plan;
* times a ``dot_general`` edge (Kronecker x operand, ``primitives.py:349-441``) it is the mixed product
  (``_mixed_mul``, ``tensor.py:1062-1231``: contract the dense pair, carry the Kronecker pair) -> one GEMM
  on the tensor cores (``csrc/gx_gemm.cu``), e.g. ``dW[H,I] = g[B,H]^T x[B,I]``;
* times a broadcast (bias) edge it is a contraction over the broadcast axis -> column sum;
* fill-in on an existing edge is ``_sparse_add`` of two dense edges -> fused into the elementwise kernels.

Precision (``precision=``): ``"bf16"`` rounds the matrix operands of every contraction to bf16, accumulation and
all elementwise arithmetic are fp32 ("bf16-in / fp32-accumulate" of BASELINE.json) - the default when every matrix
argument already IS a bf16 tensor.  ``"fp32"`` (the default otherwise) keeps the reference's float32 semantics on
the same tensor cores: every operand is split into three bf16 terms and the contraction is the sum of the six
leading term products (``gx_gemm_bf16x3_ld``), so results agree with a float32 evaluation to summation-order
noise instead of bf16 operand rounding (3e-2 of max|J|).  Results (weight Jacobians) are fp32 either way.

Anything else (other orders, vector outputs, rank > 2, batch dimensions) raises ``NotImplementedError``.
"""
from __future__ import annotations

import ctypes
from pathlib import Path
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .codegen import _FMT
from .tracer import Jaxpr, Literal, Var, make_jaxpr

Shape2 = Tuple[int, int]


def _norm(shape: Tuple[int, ...]) -> Shape2:
    if len(shape) == 0:
        return (1, 1)
    if len(shape) == 1:
        return (1, shape[0])
    if len(shape) == 2:
        return (shape[0], shape[1])
    raise NotImplementedError("tensor-valued path handles vertices of rank <= 2")


@dataclass
class TNode:
    kind: str                       # input | const | ew | gemm | colsum | sumall
    shape: Shape2
    op: str = ""                    # ew: scalar op name (lowering.OPCODES) or "copy"
    args: Tuple[int, ...] = ()
    value: float = 0.0              # const
    input_index: int = -1
    tx: bool = False                # gemm: out = op(X) . op(Y)^T with op(Z) = Z^T when the flag is set
    ty: bool = False


class TensorDag:
    def __init__(self) -> None:
        self.nodes: List[TNode] = []
        self._memo: Dict[tuple, int] = {}

    def _add(self, key: tuple, node: TNode) -> int:
        if key in self._memo:
            return self._memo[key]
        self.nodes.append(node)
        self._memo[key] = len(self.nodes) - 1
        return len(self.nodes) - 1

    def input(self, idx: int, shape: Shape2) -> int:
        return self._add(("input", idx), TNode("input", shape, input_index=idx))

    def const(self, v: float) -> int:
        return self._add(("const", float(np.float32(v))), TNode("const", (1, 1), value=float(np.float32(v))))

    def ew(self, op: str, *args: int, shape: Optional[Shape2] = None) -> int:
        shapes = [self.nodes[a].shape for a in args]
        if shape is None:
            shape = (max(s[0] for s in shapes), max(s[1] for s in shapes))
        for s in shapes:
            if (s[0] not in (1, shape[0])) or (s[1] not in (1, shape[1])):
                raise TypeError(f"incompatible shapes for elementwise {op}: {shapes}")
        if op == "mul":      # exact identities, as in the per-sample engine
            for i, j in ((0, 1), (1, 0)):
                n = self.nodes[args[i]]
                if n.kind == "const" and n.value == 1.0 and self.nodes[args[j]].shape == shape:
                    return args[j]
        return self._add(("ew", op, args, shape), TNode("ew", shape, op=op, args=tuple(args)))

    def gemm(self, x: int, tx: bool, y: int, ty: bool) -> int:
        xs, ys = self.nodes[x].shape, self.nodes[y].shape
        m, k = (xs[1], xs[0]) if tx else xs
        n, k2 = (ys[1], ys[0]) if ty else ys
        if k != k2:
            raise TypeError(f"gemm: contraction sizes differ: {xs} (t={tx}) vs {ys} (t={ty})")
        return self._add(("gemm", x, tx, y, ty), TNode("gemm", (m, n), args=(x, y), tx=tx, ty=ty))

    def colsum(self, x: int) -> int:
        return self._add(("colsum", x), TNode("colsum", (1, self.nodes[x].shape[1]), args=(x,)))

    def sumall(self, x: int) -> int:
        return self._add(("sumall", x), TNode("sumall", (1, 1), args=(x,)))

    def reduce_to(self, x: int, shape: Shape2) -> int:
        """Contract a cotangent over the axes along which its operand was broadcast."""
        s = self.nodes[x].shape
        if s == shape:
            return x
        if shape == (1, s[1]):
            return self.colsum(x)
        if shape == (1, 1):
            return self.sumall(x)
        raise NotImplementedError(f"reduction of a {s} cotangent to operand shape {shape}")


_UNARY_PARTIALS = {
    # out, x -> partial   (reference primitives.py:150-175)
    "neg": lambda d, out, x: d.const(-1.0),
    "exp": lambda d, out, x: out,
    "log": lambda d, out, x: d.ew("div", d.const(1.0), x),
    "sqrt": lambda d, out, x: d.ew("div", d.const(0.5), out),
    "tanh": lambda d, out, x: d.ew("sub", d.const(1.0), d.ew("mul", out, out)),
    "logistic": lambda d, out, x: d.ew("mul", out, d.ew("sub", d.const(1.0), out)),
    "sin": lambda d, out, x: d.ew("cos", x),
    "cos": lambda d, out, x: d.ew("neg", d.ew("sin", x)),
}


@dataclass
class DensePlan:
    jaxpr: Jaxpr
    dag: TensorDag
    argnums: List[int]
    grads: List[Optional[int]]            # per argnum: node holding the weight Jacobian (None = structurally zero)
    primal: int
    in_shapes: List[Shape2]
    steps: List[dict] = field(default_factory=list)
    stats: Dict[str, object] = field(default_factory=dict)


def build_dense_plan(jaxpr: Jaxpr, order, argnums: Sequence[int]) -> DensePlan:
    if order not in ("rev", "reverse"):
        n = len(jaxpr.eqns)
        if list(order) != list(range(n - 1, 0, -1)):
            raise NotImplementedError("the tensor-valued path implements reverse-order elimination only")
    if len(jaxpr.outvars) != 1 or _norm(jaxpr.outvars[0].shape) != (1, 1):
        raise NotImplementedError("the tensor-valued path needs a single scalar output")
    if getattr(jaxpr, "constvars", None):
        raise NotImplementedError("the tensor-valued path takes every array as an argument: pass the arrays the "
                                  "function closes over as (non-differentiated) arguments")
    dag = TensorDag()
    env: Dict[int, int] = {}
    key = lambda v: -(v.input_index + 1) if v.input_index is not None else v.eqn_index + 1
    in_shapes = [_norm(v.shape) for v in jaxpr.invars]
    for i, v in enumerate(jaxpr.invars):
        env[key(v)] = dag.input(i, in_shapes[i])

    def read(a) -> int:
        if isinstance(a, Literal):
            return dag.const(a.val)
        if key(a) in transposed_of:
            raise NotImplementedError("a transposed matrix is only supported as an operand of a contraction "
                                      "(x @ W.T) on the tensor-valued path")
        return env[key(a)]

    # `W.T` feeding a contraction (x @ W.T, how models usually write a linear layer): the transpose vertex is folded
    # into the contraction's dimension numbers - the kernel reads either operand layout in place - and its Jacobian
    # goes straight to the matrix it transposes
    transposed_of: Dict[int, Var] = {}

    def contraction_operands(eqn):
        """[(underlying variable, contracted axis of it)] for both operands of a dot_general."""
        (ca, cb), (ba, _bb) = eqn.params["dimension_numbers"]
        if len(ca) != 1 or ba or eqn.invars[0].ndim != 2 or eqn.invars[1].ndim != 2:
            raise NotImplementedError("dot_general: exactly one contracting dimension of two matrices")
        out = []
        for a, c in zip(eqn.invars, (ca[0], cb[0])):
            if isinstance(a, Var) and key(a) in transposed_of:
                out.append((transposed_of[key(a)], 1 - c))
            else:
                out.append((a, c))
        return out

    # ---- forward: primal values (core.py:380-403) ------------------------------------------------
    for eqn in jaxpr.eqns:
        p = eqn.primitive
        shape = _norm(eqn.outvar.shape)
        if p == "transpose" and eqn.invars[0].ndim == 2 and tuple(eqn.params["permutation"]) == (1, 0) \
                and isinstance(eqn.invars[0], Var) and key(eqn.invars[0]) not in transposed_of:
            transposed_of[key(eqn.outvar)] = eqn.invars[0]
            continue
        if p == "dot_general":
            (va, ca0), (vb, cb0) = contraction_operands(eqn)
            out = dag.gemm(read(va), ca0 == 0, read(vb), cb0 == 0)
            if dag.nodes[out].shape != shape:
                raise AssertionError(f"shape mismatch lowering {eqn!r}")
            env[key(eqn.outvar)] = out
            continue
        ins = [read(a) for a in eqn.invars]
        if p == "broadcast_in_dim":
            if _norm(eqn.invars[0].shape) != shape:
                raise NotImplementedError("broadcast_in_dim other than rank promotion (n,) -> (1, n)")
            out = ins[0]
        elif p == "reduce_sum":
            axes = tuple(eqn.params["axes"])
            nd = eqn.invars[0].ndim
            if axes == tuple(range(nd)):
                out = dag.sumall(ins[0])
            elif nd == 2 and axes == (0,):
                out = dag.colsum(ins[0])
            else:
                raise NotImplementedError(f"reduce_sum over axes {axes}")
        elif p == "integer_pow":
            y = int(eqn.params["y"])
            if y != 2:
                raise NotImplementedError("integer_pow with y != 2 on the tensor-valued path")
            out = dag.ew("mul", ins[0], ins[0])
        elif p in ("add", "sub", "mul", "div"):
            out = dag.ew(p, ins[0], ins[1], shape=shape)
        elif p in _UNARY_PARTIALS:
            out = dag.ew(p, ins[0])
        else:
            raise NotImplementedError(f"{p} does not have registered elemental partial (tensor-valued path).")
        if dag.nodes[out].shape != shape:
            raise AssertionError(f"shape mismatch lowering {eqn!r}")
        env[key(eqn.outvar)] = out

    # ---- reverse elimination (core.py:155-272 with a dense cotangent as the post edge) -----------
    cot: Dict[int, int] = {key(jaxpr.outvars[0]): dag.const(1.0)}
    n_products = 0

    def accumulate(v: Var, contrib: int) -> None:
        k = key(v)
        contrib = dag.reduce_to(contrib, _norm(v.shape))
        cot[k] = dag.ew("add", contrib, cot[k]) if k in cot else contrib      # product + existing edge

    for eqn in reversed(jaxpr.eqns):
        g = cot.pop(key(eqn.outvar), None)
        if g is None:
            continue                                  # no path from this vertex to the output (or a folded transpose)
        p = eqn.primitive
        if p == "dot_general":
            (va, ca0), (vb, cb0) = contraction_operands(eqn)
            operands, ins = [va, vb], [read(va), read(vb)]
        else:
            operands, ins = list(eqn.invars), [read(a) for a in eqn.invars]
        out = env[key(eqn.outvar)]
        traced = [i for i, a in enumerate(operands) if isinstance(a, Var)]
        oshape = _norm(eqn.outvar.shape)
        gfull = g if dag.nodes[g].shape == oshape else dag.ew("copy", g, shape=oshape)
        for i in traced:
            a = operands[i]
            n_products += 1
            if p == "dot_general":
                # out[M,N] = A[M,K] . B[N,K]^T with A = op_ta(a), B = op_tb(b); gemm(X,tx,Y,ty) = op(X) . op(Y)^T
                ta, tb = ca0 == 0, cb0 == 0
                if i == 0:      # dA = g . B, stored transposed when the operand was
                    eff = dag.gemm(gfull, False, ins[1], not tb) if not ta else dag.gemm(ins[1], not tb, gfull, False)
                else:           # dB = g^T . A  (the batch axis is the contraction: [B,H]^T [B,H'])
                    eff = dag.gemm(gfull, True, ins[0], not ta) if not tb else dag.gemm(ins[0], not ta, gfull, True)
                accumulate(a, eff)
            elif p == "broadcast_in_dim":
                accumulate(a, gfull)
            elif p == "reduce_sum":
                accumulate(a, dag.ew("copy", g, shape=_norm(a.shape)))
            elif p == "integer_pow":
                accumulate(a, dag.ew("mul", gfull, dag.ew("mul", dag.const(2.0), ins[0])))
            elif p == "add":
                accumulate(a, gfull)
            elif p == "sub":
                accumulate(a, gfull if i == 0 else dag.ew("neg", gfull))
            elif p == "mul":
                accumulate(a, dag.ew("mul", gfull, ins[1 - i]))
            elif p == "div":
                x, y = ins
                part = dag.ew("div", dag.const(1.0), y) if i == 0 else \
                    dag.ew("div", dag.ew("neg", x), dag.ew("mul", y, y))
                accumulate(a, dag.ew("mul", gfull, part))
            else:
                accumulate(a, dag.ew("mul", gfull, _UNARY_PARTIALS[p](dag, out, ins[0])))

    grads = [cot.get(key(jaxpr.invars[a])) for a in argnums]
    plan = DensePlan(jaxpr, dag, list(argnums), grads, env[key(jaxpr.outvars[0])], in_shapes)
    plan.stats = {"n_vertices": len(jaxpr.eqns), "n_products": n_products}
    return plan


# ------------------------------------------------------------------------------------------------
# kernel partition + execution
# ------------------------------------------------------------------------------------------------
_EW_FMT = dict(_FMT)
_EW_FMT["copy"] = "({0})"
# tanh on [B, H] tensors is the dominant elementwise cost of an MLP step: libm's tanhf is ~30 instructions, this
# form is 7 (ex2.approx + rcp.approx, absolute error ~2e-7, far below the bf16 rounding of the GEMM operands)
_EW_FMT["tanh"] = "gx_tanh({0})"
_GX_TANH = """__device__ __forceinline__ float gx_tanh(float x) {
  const float e = __expf(2.0f * fabsf(x));                 // ex2.approx; overflows to +inf -> result 1
  const float r = 1.0f - __fdividef(2.0f, e + 1.0f);
  return copysignf(r, x);
}
"""


def _padded(shape: Shape2) -> Shape2:
    """Storage extents of a [rows, cols] tensor: every axis longer than 1 is rounded up to whole 128-element
    tiles, so that kernels may read and write whole tile rows; the true extents go into the TMA tensor maps
    (out-of-range operand elements read as zero) and into the row bounds of reductions."""
    r, c = shape
    return (r if r == 1 else -(-r // 128) * 128, c if c == 1 else -(-c // 128) * 128)


class DenseExecutor:
    """Materialises the tensor DAG with the library's kernels: tcgen05 GEMMs and fused elementwise kernels
    compiled per plan (NVRTC).  torch supplies device memory and the stream only.

    Partition: contractions (GEMM) and full reductions are kernels of their own.  An elementwise value goes to
    HBM only if a contraction reads it (bf16 copy), another kernel must read it (fp32 copy) or it is a result;
    otherwise every consumer recomputes it from its materialised ancestors - a tanh is cheaper than a 16 MB
    round trip.  The contraction of a cotangent with a broadcast (bias) edge, a column sum, is folded into the
    kernel that produces the cotangent (per-thread partial sums + one atomicAdd per column)."""

    def __init__(self, plan: DensePlan, device: int = 0, with_primal: bool = False, precision: str = "bf16"):
        import torch
        from . import runtime
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.plan, self.device, self.rt, self.torch = plan, device, runtime, torch
        # fp32: contractions run as six products of bf16 split terms (gx_gemm_bf16x3_ld) on fp32 copies of their
        # operands; no fused epilogues and a single stream (the split terms of a tensor are shared by its readers)
        self.split = precision == "fp32"
        self.precision = precision
        self.lib = runtime.load_library()
        self.with_primal = with_primal
        dag = plan.dag
        self.roots = [g for g in plan.grads if g is not None] + ([plan.primal] if with_primal else [])
        self.consumers: Dict[int, List[int]] = {}
        reach, stack = set(), list(self.roots)
        while stack:
            n = stack.pop()
            if n in reach:
                continue
            reach.add(n)
            for a in dag.nodes[n].args:
                self.consumers.setdefault(a, []).append(n)
                stack.append(a)
        self.reach = reach
        heavy = lambda n: dag.nodes[n].kind in ("gemm", "colsum", "sumall")
        self.materialised = set()
        for n in sorted(reach):
            node = dag.nodes[n]
            if node.kind in ("input", "gemm", "colsum", "sumall"):
                self.materialised.add(n)
            elif node.kind == "ew" and (n in self.roots or any(heavy(c) for c in self.consumers.get(n, []))):
                self.materialised.add(n)
        # column sums of an elementwise value ride along in the kernel that computes it
        self.fused_colsums: Dict[int, List[int]] = {}
        for n in sorted(reach):
            node = dag.nodes[n]
            if node.kind == "colsum" and dag.nodes[node.args[0]].kind == "ew":
                self.fused_colsums.setdefault(node.args[0], []).append(n)
        fused = {c for cs in self.fused_colsums.values() for c in cs}
        self.needs_bf16 = {a for n in reach if dag.nodes[n].kind == "gemm" for a in dag.nodes[n].args}
        self.order = [n for n in sorted(reach)
                      if n in self.materialised and dag.nodes[n].kind != "input" and n not in fused]
        self.kernels: Dict[int, object] = {}
        self.buffers: Dict[Tuple[int, str], object] = {}
        self.launches = 0
        self._ew_meta: Dict[int, tuple] = {}
        self.input_dtypes: Optional[List[str]] = None     # elementwise kernels are compiled once these are known
        self.fuse = os.environ.get("GX_DENSE_FUSE", "1") != "0" and not self.split
        self.gemm_fused: Dict[int, List[int]] = {}      # gemm node -> elementwise nodes evaluated in its epilogue
        self.fused_into: Dict[int, int] = {}
        if self.fuse:
            self._plan_fusion()
        # which fp32 copies does anybody read?  (an epilogue reads its own accumulator from registers)
        self.f32_read = set(self.roots)
        for n in sorted(reach):
            node = dag.nodes[n]
            if n not in self.materialised or n in fused:
                continue
            if node.kind == "ew":
                self.f32_read.update(l for l in self._leaves_and_expr(n)[0] if l != self.fused_into.get(n))
            elif node.kind in ("colsum", "sumall"):
                self.f32_read.add(node.args[0])
        if self.split:
            self.f32_read.update(self.needs_bf16)
        self.order = [n for n in self.order if n not in self.fused_into]
        self.fused_kernels: Dict[int, object] = {}
        self._fused_meta: Dict[int, tuple] = {}
        # accumulators that kernels add into (column sums, split-K outputs, full reductions) live in one arena
        # that a single memset clears at the start of a step
        self.arena_keys: Dict[Tuple[int, str], Tuple[int, Shape2]] = {}
        off = 0
        def reserve(key, shape):
            nonlocal off
            if key not in self.arena_keys:
                shape = _padded(shape)
                self.arena_keys[key] = (off, shape)
                off += (shape[0] * shape[1] + 3) // 4 * 4
        for n in sorted(reach):
            node = dag.nodes[n]
            if n not in self.materialised:
                continue
            if node.kind == "gemm" and n in self.order and (self._gemm_geometry(n)[3] > 1 or self.split):
                reserve((n, "f32"), node.shape)
            elif node.kind in ("colsum", "sumall"):
                reserve((n, "f32"), node.shape)
                if node.kind == "sumall":
                    reserve((n, "part"), (1, dag.nodes[node.args[0]].shape[1]))
        self.arena_size = off
        self.arena = None
        self.input_stage: Dict[Tuple[int, str], object] = {}
        # one stream by default: the programmatic-dependent-launch chain of the five contractions (55.4 us per MLP
        # step, B200 r02) beats running the two weight-Jacobian contractions on a side stream (57.4 us: the fork /
        # join edges of the captured graph cost more than the overlap gains); GX_DENSE_STREAMS=2 restores the latter
        self.concurrent = os.environ.get("GX_DENSE_STREAMS", "1") == "2" and not self.split
        self._split_bufs: Dict[int, list] = {}
        self._split_done: set = set()
        # kernels that add into the arena, and results that nothing downstream reads (side-stream candidates)
        self.arena_users = set()
        for n in self.order:
            if (n, "f32") in self.arena_keys:
                self.arena_users.add(n)
            for o in [n] + self.gemm_fused.get(n, []):
                if o in self.fused_colsums:
                    self.arena_users.add(n)
        self.side_nodes = {n for n in self.order if dag.nodes[n].kind == "gemm" and n in self.roots
                           and not self.consumers.get(n) and n not in self.gemm_fused}

    # ---- fused elementwise kernels -------------------------------------------------------------
    def _leaves_and_expr(self, root: int):
        dag = self.plan.dag
        leaves: List[int] = []
        lines: List[str] = []
        names: Dict[int, str] = {}

        def visit(n: int) -> str:
            if n in names:
                return names[n]
            node = dag.nodes[n]
            if node.kind == "const":
                bits = int(np.array(node.value, dtype=np.float32).view(np.uint32))
                names[n] = f"__int_as_float(0x{bits:08x})"
                return names[n]
            if n != root and n in self.materialised:
                if n not in leaves:
                    leaves.append(n)
                names[n] = f"l{leaves.index(n)}_@"
                return names[n]
            args = [visit(a) for a in node.args]
            names[n] = f"t{n}_@"
            lines.append(f"      const float t{n}_@ = {_EW_FMT[node.op].format(*args)};")
            return names[n]
        result = visit(root)
        return leaves, lines, result

    def _leaf_dtype(self, leaf: int) -> str:
        node = self.plan.dag.nodes[leaf]
        return self.input_dtypes[node.input_index] if node.kind == "input" else "f32"

    def _compile_ew(self, n: int) -> None:
        """One kernel per materialised elementwise node: every thread owns 4 consecutive columns (128-bit
        loads/stores) and strides over rows; optional outputs: fp32 copy, bf16 copy, column sums."""
        dag = self.plan.dag
        leaves, lines, result = self._leaves_and_expr(n)
        rows, cols = _padded(dag.nodes[n].shape)      # kernels see the storage width; launches pass the true row count
        vec = 4 if cols % 4 == 0 else 1
        params, loads = [], []
        for i, leaf in enumerate(leaves):
            lr, lc = dag.nodes[leaf].shape
            bf = self._leaf_dtype(leaf) == "bf16"
            params.append(f"const {'unsigned short' if bf else 'float'}* __restrict__ p{i}")
            for j in range(vec):
                idx = "0" if (lr, lc) == (1, 1) else (f"c + {j}" if lr == 1 else ("r" if lc == 1 else f"i + {j}"))
                src = f"p{i}[{idx}]"
                loads.append(f"      const float l{i}_{j} = {'gx_bf16_to_f32(' + src + ')' if bf else src};")
        body: List[str] = []
        for j in range(vec):
            body += [ln.replace("_@", f"_{j}") for ln in lines]
            body.append(f"      const float y{j} = {result.replace('_@', f'_{j}')};")
        stores = []
        if vec == 4:
            stores.append("      if (o32) *reinterpret_cast<float4*>(o32 + i) = make_float4(y0, y1, y2, y3);")
            stores.append("      if (o16) *reinterpret_cast<uint2*>(o16 + i) = make_uint2(gx_f32_to_bf16(y0) | ((unsigned)gx_f32_to_bf16(y1) << 16), "
                          "gx_f32_to_bf16(y2) | ((unsigned)gx_f32_to_bf16(y3) << 16));")
            stores.append("      s0 += y0; s1 += y1; s2 += y2; s3 += y3;")
            tail = ["  if (csum) {   // fold the 8 row lanes of the CTA in shared memory: one atomic per column and CTA",
                    "    part[ty][tx] = make_float4(s0, s1, s2, s3);",
                    "    __syncthreads();",
                    "    if (ty == 0 && c < cols) {",
                    "      for (int k = 1; k < 8; ++k) { const float4 q = part[k][tx]; s0 += q.x; s1 += q.y; s2 += q.z; s3 += q.w; }",
                    "      atomicAdd(csum + c, s0); atomicAdd(csum + c + 1, s1); atomicAdd(csum + c + 2, s2); atomicAdd(csum + c + 3, s3);",
                    "    }",
                    "  }"]
        else:
            stores += ["      if (o32) o32[i] = y0;", "      if (o16) o16[i] = gx_f32_to_bf16(y0);", "      s0 += y0;"]
            tail = ["  if (csum && c < cols) atomicAdd(csum + c, s0);"]
        src = "\n".join([
            f"// fused elementwise kernel for tensor node {n}: [{rows} x {cols}] from {len(leaves)} operand(s)",
            "__device__ __forceinline__ float gx_bf16_to_f32(unsigned short h) { return __uint_as_float((unsigned)h << 16); }",
            "__device__ __forceinline__ unsigned short gx_f32_to_bf16(float f) {   // round to nearest even",
            "  unsigned u = __float_as_uint(f);",
            "  if ((u & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((u >> 16) | 0x40u);",
            "  return (unsigned short)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);",
            "}",
            _GX_TANH,
            "extern \"C\" __global__ void gx_ew(" + ", ".join(params + [
                "float* __restrict__ o32", "unsigned short* __restrict__ o16", "float* __restrict__ csum",
                "int rows", "int cols"]) + ") {",
            "  __shared__ float4 part[8][32];",
            "  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 column groups x 8 row lanes",
            f"  const int c = (blockIdx.x * 32 + tx) * {vec};",
            "  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f; (void)s1; (void)s2; (void)s3; (void)part;",
            "  if (c < cols)",
            "  for (int r = blockIdx.y * 8 + ty; r < rows; r += gridDim.y * 8) {",
            "      const long long i = (long long)r * cols + c;",
            *loads, *body, *stores,
            "  }", *tail, "}"]) + "\n"
        h = ctypes.c_void_p()
        rc = self.lib.gx_jit_compile(src.encode(), b"gx_ew", self.device, ctypes.byref(h))
        if rc != 0:
            raise self.rt.GraphaxB200Error(f"gx_jit_compile failed ({rc}): {self.lib.gx_last_error().decode()}")
        self.kernels[n] = h
        self._ew_meta[n] = (leaves, vec)

    # ---- elementwise consumers fused into the contraction's epilogue ------------------------------------
    def _gemm_geometry(self, n: int):
        """(M, N, K, k_splits) of gemm node n: op(X)[M,K] . op(Y)[N,K]^T."""
        dag = self.plan.dag
        node = dag.nodes[n]
        xs, ys = dag.nodes[node.args[0]].shape, dag.nodes[node.args[1]].shape
        m, k = xs[::-1] if node.tx else xs
        nn = ys[1] if node.ty else ys[0]
        tiles = -(-m // 128) * -(-nn // 128)
        splits = 1
        if tiles < 96:
            limit = int(os.environ.get("GX_SPLITK_CTAS", "160"))
            for sp in (16, 8, 4, 2):
                if -(-k // 64) >= 4 * sp and tiles * sp <= limit:      # at least four 64-wide K blocks per split
                    splits = sp
                    break
        return m, nn, k, splits

    def _plan_fusion(self) -> None:
        """An elementwise node moves into a contraction's epilogue when the contraction result is one of its
        operands, has its shape and is not split along K, and every other operand is already in memory when
        the contraction runs (an input, or something materialised earlier)."""
        dag = self.plan.dag
        pos = {n: i for i, n in enumerate(self.order)}
        for n in list(self.order):
            node = dag.nodes[n]
            if node.kind != "ew":
                continue
            leaves = self._leaves_and_expr(n)[0]
            gemms = [l for l in leaves if dag.nodes[l].kind == "gemm" and dag.nodes[l].shape == node.shape]
            if not gemms:
                continue
            g = max(gemms, key=lambda l: pos[l])              # the last one to be computed hosts the epilogue
            if self._gemm_geometry(g)[3] != 1:
                continue
            ok = True
            for l in leaves:
                if l == g:
                    continue
                ln = dag.nodes[l]
                shape_ok = ln.shape in (node.shape, (1, node.shape[1]), (node.shape[0], 1), (1, 1))
                ready = ln.kind == "input" or (l in pos and pos[l] < pos[g])
                if not (shape_ok and ready):
                    ok = False
            if ok:
                self.gemm_fused.setdefault(g, []).append(n)
                self.fused_into[n] = g

    def _compile_fused(self, g: int) -> None:
        """gx_gemm_kernel.cuh + a generated epilogue: per 32-column chunk every lane (= tile row) evaluates the
        fused expressions on its accumulator values four columns at a time, stores fp32 / bf16 copies with
        128 / 64-bit stores and folds column sums across the warp's 32 rows with 31 shuffles."""
        import re
        dag = self.plan.dag
        M, N = dag.nodes[g].shape
        outs = [g] + self.gemm_fused[g]                     # output slot 0 = the contraction result itself
        leaves: List[int] = []                              # global operand list (memory operands only)
        bodies = []
        for n in outs[1:]:
            lv, lines, result = self._leaves_and_expr(n)
            def sub(text: str) -> str:
                def repl(mo):
                    leaf = lv[int(mo.group(1))]
                    if leaf == g:
                        return "acc_@"
                    if leaf not in leaves:
                        leaves.append(leaf)
                    return f"L{leaves.index(leaf)}_@"
                return re.sub(r"\bl(\d+)_@", repl, text)
            bodies.append(([sub(ln) for ln in lines], sub(result)))
        nl, no = max(1, len(leaves)), len(outs)
        tiles_m, tiles_n = -(-M // 128), -(-N // 128)
        deep = tiles_m * tiles_n <= 148                          # single-wave grids: deep operand ring, one CTA per SM
        # CTA pairs (cta_group::2: 256 x 128 per cluster of two CTAs, 24 instead of 32 KB of operands per CTA and
        # k-block) whenever the tile rows pair up; the stage count names the variant (gx_gemm_kernel.cuh)
        pair = tiles_m % 2 == 0 and os.environ.get("GX_GEMM_PAIR", "0") == "1"      # off by default: measured neutral
        stages = (8 if deep else 4) if pair else (6 if deep else 3)
        min_blocks = 1 if deep else 2
        src = ["#define GX_NVRTC", "#define GX_GEMM_FUSED", "#define GX_GEMM_KERNEL_NAME gx_gemm_fused",
               f"#define GX_GEMM_STAGES {stages}", f"#define GX_GEMM_MIN_BLOCKS {min_blocks}"]
        epi = [_GX_TANH,
               f"struct GxEpi {{ const void* leaf[{nl}]; float* o32[{no}]; unsigned short* o16[{no}]; float* csum[{no}]; }};",
               "struct GxEpiState { int unused; };",
               "__device__ __forceinline__ void gx_epi_begin(GxEpiState&) {}",
               "__device__ __forceinline__ void gx_epi_finish(const GxEpi&, GxEpiState&, int) {}",
               ]
        # memory operands: fetched by gx_epi_prefetch (issued while the main loop still runs), consumed by gx_epi_chunk
        pre_fields, pre_loads, chunk_head, chunk_quad = [], [], [], []
        for i, leaf in enumerate(leaves):
            lr, lc = dag.nodes[leaf].shape
            bf = self._leaf_dtype(leaf) == "bf16"
            ptr = f"reinterpret_cast<const {'unsigned short' if bf else 'float'}*>(e.leaf[{i}])"
            if (lr, lc) == (M, N) and not bf:
                # coalesced row-contiguous loads now, through the staging tile to "one row per lane" later
                pre_fields.append(f"float4 M{i}[8];")
                pre_loads.append(f"  gxg::fetch_rows_f32(p.M{i}, {ptr}, N, m_base, M, col0, lane);")
                chunk_head.append(f"  float Lm{i}[32];")
                chunk_head.append(f"  gxg::spread_rows_f32(tile, Lm{i}, p.M{i}, lane);")
                chunk_quad.append(f"    const float L{i}_0 = Lm{i}[4 * q], L{i}_1 = Lm{i}[4 * q + 1], L{i}_2 = Lm{i}[4 * q + 2], "
                                  f"L{i}_3 = Lm{i}[4 * q + 3];")
            elif (lr, lc) == (1, 1) or lc == 1:
                idx = "0" if (lr, lc) == (1, 1) else "(ok ? row : 0)"
                val = f"gxg::bf16_bits_to_f32({ptr}[{idx}])" if bf else f"__ldg({ptr} + {idx})"
                pre_fields.append(f"float S{i};")
                pre_loads.append(f"  p.S{i} = {val};")
                chunk_quad.append(f"    const float L{i}_0 = p.S{i}, L{i}_1 = L{i}_0, L{i}_2 = L{i}_0, L{i}_3 = L{i}_0;")
            else:
                base = "col0 + 4 * q" if lr == 1 else "off + 4 * q"
                if bf:
                    pre_fields.append(f"uint2 R{i}[8];")
                    pre_loads.append(f"  _Pragma(\"unroll\") for (int q = 0; q < 8; ++q) p.R{i}[q] = "
                                     f"__ldg(reinterpret_cast<const uint2*>({ptr} + {base}));")
                    chunk_quad.append(f"    const uint2 R{i} = p.R{i}[q];")
                    chunk_quad.append(f"    const float L{i}_0 = gxg::bf16_bits_to_f32(R{i}.x & 0xffffu), L{i}_1 = gxg::bf16_bits_to_f32(R{i}.x >> 16), "
                                      f"L{i}_2 = gxg::bf16_bits_to_f32(R{i}.y & 0xffffu), L{i}_3 = gxg::bf16_bits_to_f32(R{i}.y >> 16);")
                else:
                    pre_fields.append(f"float4 R{i}[8];")
                    pre_loads.append(f"  _Pragma(\"unroll\") for (int q = 0; q < 8; ++q) p.R{i}[q] = "
                                     f"__ldg(reinterpret_cast<const float4*>({ptr} + {base}));")
                    chunk_quad.append(f"    const float4 R{i} = p.R{i}[q];")
                    chunk_quad.append(f"    const float L{i}_0 = R{i}.x, L{i}_1 = R{i}.y, L{i}_2 = R{i}.z, L{i}_3 = R{i}.w;")
        epi += ["struct GxEpiPre { " + " ".join(pre_fields or ["int unused;"]) + " };",
                "__device__ __forceinline__ void gx_epi_prefetch(const GxEpi& e, GxEpiPre& p, int m_base, int col0, int M, int N,",
                "                                                int lane) {",
                "  const int row = m_base + lane;",
                "  const bool ok = row < M;",
                "  const size_t off = (size_t)(ok ? row : 0) * N + col0;",
                "  (void)off;"] + pre_loads + ["}",
                "__device__ __forceinline__ void gx_epi_chunk(const GxEpi& e, GxEpiState&, const GxEpiPre& p, float* tile, int m_base,",
                "                                             int col0, int M, int N, const float (&acc)[32], int lane) {",
                "  const int row = m_base + lane;",
                "  const bool ok = row < M;"]
        csum_slots = [k for k, n in enumerate(outs) if k > 0 and n in self.fused_colsums]
        for k in range(1, no):
            epi.append(f"  float Y{k}[32];")
        epi += chunk_head
        epi.append("#pragma unroll")
        epi.append("  for (int q = 0; q < 8; ++q) {")
        epi += chunk_quad
        for j in range(4):
            epi.append(f"    const float acc_{j} = acc[4 * q + {j}];")
        emitted = set()                                 # temporaries are named after their node: shared by outputs
        for k, (lines, result) in enumerate(bodies, start=1):
            for j in range(4):
                for ln in lines:
                    ln = ln.replace("_@", f"_{j}").replace("      const", "    const")
                    if ln not in emitted:
                        emitted.add(ln)
                        epi.append(ln)
                epi.append(f"    Y{k}[4 * q + {j}] = {result.replace('_@', f'_{j}')};")
        epi.append("  }")
        # GX_EPI_DEBUG (measurement only, results are wrong): leave parts of the epilogue out to see what they cost
        debug = set(filter(None, os.environ.get("GX_EPI_DEBUG", "").split(",")))
        # row-contiguous stores through the warp's staging tile (gx_gemm_kernel.cuh: store_rows)
        if "nostore" not in debug:
            # only outputs somebody reads get a store: a call whose pointers are null at run time would still keep
            # its 32 values alive (the raw accumulator next to every fused output: 200 bytes of spills at 96 registers)
            stored = lambda o: o in self.f32_read or o in self.needs_bf16
            if stored(g):
                epi.append("  gxg::store_rows(tile, acc, e.o32[0], e.o16[0], N, m_base, M, col0, lane);")
            for k in range(1, no):
                if stored(outs[k]):
                    epi.append(f"  gxg::store_rows(tile, Y{k}, e.o32[{k}], e.o16[{k}], N, m_base, M, col0, lane);")
        else:
            for k in range(1, no):
                epi.append(f"  if (Y{k}[lane] == 12345.678f) e.o32[{k}][0] = Y{k}[0];")
        for k in csum_slots:
            if "nocsum" in debug:
                continue
            if M % 128:                                     # rows of the last tile beyond M stay out of the sums
                epi.append("#pragma unroll")
                epi.append(f"  for (int j = 0; j < 32; ++j) Y{k}[j] = ok ? Y{k}[j] : 0.0f;")
            epi.append(f"  gxg::warp_transpose_reduce(Y{k}, lane);")
            if "noatomic" in debug:
                epi.append(f"  if (Y{k}[0] == 12345.678f) atomicAdd(e.csum[{k}] + col0 + lane, Y{k}[0]);")
            else:
                epi.append(f"  atomicAdd(e.csum[{k}] + col0 + lane, Y{k}[0]);")
        epi.append("}")
        if "notanh" in debug:
            epi = [ln if "__device__" in ln else ln.replace("gx_tanh(", "0.5f * (") for ln in epi]
        header = (Path(__file__).resolve().parent / "csrc" / "gx_gemm_kernel.cuh").read_text().replace("#pragma once", "")
        marker = "#ifndef GX_GEMM_KERNEL_NAME"
        text = "\n".join(src) + "\n" + header.replace(marker, "\n".join(epi) + "\n" + marker, 1)
        h = ctypes.c_void_p()
        rc = self.lib.gx_jit_compile(text.encode(), b"gx_gemm_fused", self.device, ctypes.byref(h))
        if rc != 0:
            raise self.rt.GraphaxB200Error(f"gx_jit_compile(fused gemm) failed ({rc}): {self.lib.gx_last_error().decode()}")
        self.fused_kernels[g] = h
        self._fused_meta[g] = (leaves, outs, csum_slots, stages)

    # ---- execution -------------------------------------------------------------------------------
    def _buf(self, n: int, kind: str):
        key = (n, kind)
        if key not in self.buffers and key in self.arena_keys:
            if self.arena is None:
                self.arena = self.torch.zeros(max(4, self.arena_size), dtype=self.torch.float32, device=f"cuda:{self.device}")
            off, shape = self.arena_keys[key]
            self.buffers[key] = self.arena[off:off + shape[0] * shape[1]].view(shape)
        if key not in self.buffers:
            dt = self.torch.float32 if kind == "f32" else self.torch.bfloat16
            self.buffers[key] = self.torch.zeros(_padded(self.plan.dag.nodes[n].shape), dtype=dt, device=f"cuda:{self.device}")
        return self.buffers[key]

    def _stage_input(self, i: int, t, dtype):
        """An argument in tile-padded storage of the given dtype (the tensor itself when it already is)."""
        shape = tuple(t.shape)
        if _padded(shape) == shape and t.dtype == dtype:
            return t
        key = (i, str(dtype))
        buf = self.input_stage.get(key)
        if buf is None:
            buf = self.input_stage[key] = self.torch.zeros(_padded(shape), dtype=dtype, device=t.device)
        buf[:shape[0], :shape[1]].copy_(t)
        return buf

    def _bf16(self, n: int):
        """bf16 copy of node n as stored ([rows, cols] row-major); the GEMM reads it K-major or MN-major."""
        node = self.plan.dag.nodes[n]
        return self.inputs_bf16[node.input_index] if node.kind == "input" else self._buf(n, "bf16")

    def _split3(self, n: int, stream) -> list:
        """The three bf16 terms of node n's fp32 value (computed once per step, right before its first reader)."""
        bufs = self._split_bufs.get(n)
        src = self._data(n)
        if bufs is None:
            bufs = self._split_bufs[n] = [self.torch.zeros(tuple(src.shape), dtype=self.torch.bfloat16, device=src.device)
                                          for _ in range(3)]
        if n not in self._split_done:
            rc = self.lib.gx_split3_bf16(src.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(),
                                         src.numel(), stream)
            if rc != 0:
                raise self.rt.GraphaxB200Error(self.lib.gx_gemm_last_error().decode())
            self._split_done.add(n)
            self.launches += 1
        return bufs

    def _launch_node(self, n: int) -> None:
        """Enqueue the kernel(s) that materialise node n on the current stream."""
        torch, dag = self.torch, self.plan.dag
        stream = torch.cuda.current_stream(self.device).cuda_stream
        data = self._data
        node = dag.nodes[n]
        if node.kind == "gemm" and n in self.fused_kernels:
            x, y = node.args
            a, b = self._bf16(x), self._bf16(y)
            m, nn, k, _ = self._gemm_geometry(n)
            leaves, outs, csum_slots, stages = self._fused_meta[n]
            ptrs = [data(l).data_ptr() for l in leaves] or [0]
            o32, o16, cs = [], [], []
            for slot, o in enumerate(outs):
                o32.append(self._buf(o, "f32").data_ptr() if o in self.f32_read else 0)
                o16.append(self._buf(o, "bf16").data_ptr() if o in self.needs_bf16 else 0)
                if slot in csum_slots:
                    cs.append(self._buf(self.fused_colsums[o][0], "f32").data_ptr())
                else:
                    cs.append(0)
            block = (ctypes.c_void_p * (len(ptrs) + 3 * len(outs)))(*(ptrs + o32 + o16 + cs))
            rc = self.lib.gx_gemm_fused_launch(self.fused_kernels[n], m, nn, k, a.data_ptr(), a.shape[1], int(node.tx),
                                               b.data_ptr(), b.shape[1], int(node.ty), _padded(node.shape)[1],
                                               ctypes.cast(block, ctypes.c_void_p), stages, stream)
            if rc != 0:
                raise self.rt.GraphaxB200Error(self.lib.gx_last_error().decode())
            self.launches += 1
            for o in outs[1:]:
                for extra in self.fused_colsums.get(o, [])[1:]:
                    self._buf(extra, "f32").copy_(self._buf(self.fused_colsums[o][0], "f32"))
        elif node.kind == "gemm" and self.split:
            x, y = node.args
            a3, b3 = self._split3(x, stream), self._split3(y, stream)
            m, nn, k, splits = self._gemm_geometry(n)
            out = self._buf(n, "f32")          # in the arena: cleared at the start of the step, all six products add
            vp = ctypes.c_void_p
            rc = self.lib.gx_gemm_bf16x3_ld(m, nn, k, (vp * 3)(*[t.data_ptr() for t in a3]), a3[0].shape[1], int(node.tx),
                                            (vp * 3)(*[t.data_ptr() for t in b3]), b3[0].shape[1], int(node.ty),
                                            out.data_ptr(), out.shape[1], splits, stream)
            if rc != 0:
                raise self.rt.GraphaxB200Error(self.lib.gx_gemm_last_error().decode())
            self.launches += 6
        elif node.kind == "gemm":
            x, y = node.args
            a, b = self._bf16(x), self._bf16(y)  # op(X) = X^T is read in place as an MN-major operand
            m, nn, k, splits = self._gemm_geometry(n)
            out = self._buf(n, "f32")          # split-K accumulators live in the arena the step has already cleared
            rc = self.lib.gx_gemm_bf16_ld(m, nn, k, a.data_ptr(), a.shape[1], int(node.tx), b.data_ptr(), b.shape[1],
                                          int(node.ty), out.data_ptr(), out.shape[1], 0, splits, stream)
            if rc != 0:
                raise self.rt.GraphaxB200Error(self.lib.gx_gemm_last_error().decode())
            self.launches += 1
            if n in self.needs_bf16:
                self._buf(n, "bf16").copy_(out)
                self.launches += 1
        elif node.kind in ("colsum", "sumall"):
            src = data(node.args[0])
            out = self._buf(n, "f32")
            r, c = dag.nodes[node.args[0]].shape      # true extents: padding rows / columns stay out of the sums
            ld = src.shape[1]
            if node.kind == "colsum":
                rcs = [self.lib.gx_colsum_f32(src.data_ptr(), out.data_ptr(), r, ld, stream)]
            else:                               # all elements: column sums over the true rows, then the true columns
                part = self._buf(n, "part")
                rcs = [self.lib.gx_colsum_f32(src.data_ptr(), part.data_ptr(), r, ld, stream),
                       self.lib.gx_colsum_f32(part.data_ptr(), out.data_ptr(), c, 1, stream)]
            self.launches += len(rcs)
            if any(rcs):
                raise self.rt.GraphaxB200Error(self.lib.gx_last_error().decode())
        else:
            leaves, vec = self._ew_meta[n]
            r, c = node.shape[0], _padded(node.shape)[1]      # true rows (column sums), padded width (= row stride)
            o32 = self._buf(n, "f32").data_ptr() if n in self.f32_read else 0
            o16 = self._buf(n, "bf16").data_ptr() if (n in self.needs_bf16 and not self.split) else 0
            csum = self._buf(self.fused_colsums[n][0], "f32").data_ptr() if n in self.fused_colsums else 0
            vals = [ctypes.c_void_p(data(l).data_ptr()) for l in leaves]
            vals += [ctypes.c_void_p(o32), ctypes.c_void_p(o16), ctypes.c_void_p(csum), ctypes.c_int(r), ctypes.c_int(c)]
            argv = (ctypes.c_void_p * len(vals))(*[ctypes.cast(ctypes.pointer(v), ctypes.c_void_p) for v in vals])
            gx = (c // vec + 31) // 32
            gy = max(1, min((r + 7) // 8, (148 * 6) // gx))
            rc = self.lib.gx_jit_launch(self.kernels[n], gx, gy, 256, argv, stream)
            if rc != 0:
                raise self.rt.GraphaxB200Error(self.lib.gx_last_error().decode())
            self.launches += 1
            for extra in self.fused_colsums.get(n, [])[1:]:       # same sums requested twice: copy
                self._buf(extra, "f32").copy_(self._buf(self.fused_colsums[n][0], "f32"))

    def _plain_gemm(self, n: int) -> bool:
        """A contraction launched by the library kernel as it is: no fused epilogue, no three-term split, and nobody
        needs a bf16 copy of its result."""
        return (self.plan.dag.nodes[n].kind == "gemm" and n not in self.gemm_fused and not self.split
                and n not in self.needs_bf16)

    def launch_groups(self, host_zero: bool = False) -> List[Tuple[int, ...]]:
        """The step's launches in order: single nodes, or pairs of consecutive plain contractions that do not depend
        on each other and share one launch (``_launch_dual``).  ``host_zero``: the first launch also clears the arena
        and stays single."""
        dag = self.plan.dag
        dual = os.environ.get("GX_DENSE_DUAL", "1") != "0" and not self.concurrent
        groups: List[Tuple[int, ...]] = []
        pos = 0
        while pos < len(self.order):
            n = self.order[pos]
            nxt = self.order[pos + 1] if pos + 1 < len(self.order) else None
            if (dual and nxt is not None and self._plain_gemm(n) and self._plain_gemm(nxt)
                    and n not in dag.nodes[nxt].args and not (host_zero and pos == 0)):
                groups.append((n, nxt))
                pos += 2
            else:
                groups.append((n,))
                pos += 1
        return groups

    def _launch_dual(self, n0: int, n1: int) -> None:
        """Two independent plain contractions in one launch (gx_gemm_bf16_dual_ld): the weight Jacobians of adjacent
        layers are each half a wave of split-K tiles; together they fill the GPU once and pay one pipeline fill, one
        epilogue tail and one hand-over instead of two."""
        dag = self.plan.dag
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        cols = {k: [] for k in ("M", "N", "K", "A", "lda", "amn", "B", "ldb", "bmn", "C", "ldc", "c16", "splits")}
        for n in (n0, n1):
            node = dag.nodes[n]
            a, b = self._bf16(node.args[0]), self._bf16(node.args[1])
            m, nn, k, splits = self._gemm_geometry(n)
            out = self._buf(n, "f32")
            for key, v in zip(cols, (m, nn, k, a.data_ptr(), a.shape[1], int(node.tx), b.data_ptr(), b.shape[1],
                                     int(node.ty), out.data_ptr(), out.shape[1], 0, splits)):
                cols[key].append(v)
        ints = lambda key: (ctypes.c_int * 2)(*cols[key])
        ptrs = lambda key: (ctypes.c_void_p * 2)(*cols[key])
        rc = self.lib.gx_gemm_bf16_dual_ld(ints("M"), ints("N"), ints("K"), ptrs("A"), ints("lda"), ints("amn"), ptrs("B"),
                                           ints("ldb"), ints("bmn"), ptrs("C"), ints("ldc"), ints("c16"), ints("splits"),
                                           stream)
        if rc != 0:
            raise self.rt.GraphaxB200Error(self.lib.gx_gemm_last_error().decode())
        self.launches += 1

    def run(self, inputs: Sequence) -> Tuple[List, object]:
        torch, dag = self.torch, self.plan.dag
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.inputs_f32, self.inputs_bf16 = {}, {}
        self._split_done = set()
        dts = []
        for i, t in enumerate(inputs):
            t = t.reshape(self.plan.in_shapes[i]).contiguous()
            if self.split and t.dtype != torch.float32:
                t = t.float()
            dts.append("bf16" if t.dtype == torch.bfloat16 else "f32")
            if t.dtype == torch.bfloat16:
                self.inputs_bf16[i] = self._stage_input(i, t, torch.bfloat16)
            else:
                self.inputs_f32[i] = self._stage_input(i, t, torch.float32)
        if self.input_dtypes != dts:
            self.input_dtypes = dts                 # elementwise kernels read inputs in their own dtype
            for n in list(self.kernels):
                self.lib.gx_jit_destroy(self.kernels[n])
            self.kernels.clear()
            for n in self.order:
                if dag.nodes[n].kind == "ew":
                    self._compile_ew(n)
            for h in self.fused_kernels.values():
                self.lib.gx_jit_destroy(h)
            self.fused_kernels.clear()
            for g in self.gemm_fused:
                self._compile_fused(g)
        for i in range(len(inputs)):                # GEMM operands must be bf16
            if not self.split and i not in self.inputs_bf16 and any(dag.nodes[n].kind == "input" and dag.nodes[n].input_index == i
                                                 for n in self.needs_bf16):
                self.inputs_bf16[i] = self._stage_input(i, inputs[i].reshape(self.plan.in_shapes[i]), torch.bfloat16)

        def data(n: int):
            node = dag.nodes[n]
            if node.kind == "input":
                return self.inputs_bf16[node.input_index] if self.input_dtypes[node.input_index] == "bf16" \
                    else self.inputs_f32[node.input_index]
            return self._buf(n, "f32")

        self._data = data
        # Two streams: the accumulator memset and the contractions nobody downstream waits for (the weight
        # Jacobians themselves) run beside the critical chain forward -> cotangents; under CUDA-graph capture the
        # events become parallel branches.
        main = torch.cuda.current_stream(self.device)
        # The accumulator arena is cleared by the first contraction of the step itself (its epilogue warps are idle
        # during the main loop) when that contraction does not add into the arena: no memset launch, and every later
        # launch keeps a single, programmatic dependency (a second edge from a memset on the side stream made the
        # first arena user start 1.8 us late: profiles/gemm_trace_r02c_mlp.txt).  Otherwise: memset beside the chain.
        first = self.order[0] if self.order else None
        host_zero = bool(self.arena_size and first is not None and dag.nodes[first].kind == "gemm" and not self.split
                         and not self.concurrent and first not in self.arena_users
                         and os.environ.get("GX_DENSE_MEMSET", "0") != "1")
        use_side = self.concurrent or bool(self.arena_size and not host_zero)
        side = None
        if use_side:
            if getattr(self, "side", None) is None:
                self.side = torch.cuda.Stream(device=self.device)
            side = self.side
            fork = torch.cuda.Event()
            fork.record(main)
            side.wait_event(fork)
        zeroed = None
        if self.arena_size:
            for key in self.arena_keys:
                self._buf(*key)
            if host_zero:
                rc = self.lib.gx_gemm_zero_with_next_launch(self.arena.data_ptr(), self.arena.numel() * 4)
                if rc != 0:
                    raise self.rt.GraphaxB200Error(self.lib.gx_gemm_last_error().decode())
            else:
                with torch.cuda.stream(side):
                    self.arena.zero_()
                    zeroed = torch.cuda.Event()
                    zeroed.record(side)
                self.launches += 1
        for group in self.launch_groups(host_zero):
            n = group[0]
            if len(group) == 2:
                if zeroed is not None and (group[0] in self.arena_users or group[1] in self.arena_users):
                    main.wait_event(zeroed)
                    zeroed = None
                self._launch_dual(*group)
                continue
            if n in self.side_nodes and self.concurrent:
                ready = torch.cuda.Event()
                ready.record(main)
                side.wait_event(ready)
                with torch.cuda.stream(side):
                    self._launch_node(n)
                continue
            if zeroed is not None and n in self.arena_users:
                main.wait_event(zeroed)
                zeroed = None
            self._launch_node(n)
        if use_side:
            join = torch.cuda.Event()
            join.record(side)
            main.wait_event(join)
        grads = []
        for a, g in zip(self.plan.argnums, self.plan.grads):
            shape = self.plan.jaxpr.invars[a].shape
            if g is None:
                grads.append(torch.zeros(shape, device=f"cuda:{self.device}"))
            else:
                r, c = dag.nodes[g].shape
                grads.append(data(g)[:r, :c].float().reshape(shape))
        return grads, (data(self.plan.primal)[0, 0].reshape(()) if self.with_primal else None)


def default_precision(args) -> str:
    """"bf16" when every matrix argument already is a bf16 tensor (nothing is lost by bf16 operands), else "fp32"."""
    import torch
    mats = [a for a in args if getattr(a, "ndim", 0) >= 2]
    return "bf16" if mats and all(isinstance(a, torch.Tensor) and a.dtype == torch.bfloat16 for a in mats) else "fp32"


def dense_jacve(fun, order, argnums, *args, has_aux: bool = False, device: int = 0, precision: Optional[str] = None,
                frontend: str = "auto", _cache={}):
    """``jacve`` for a scalar function of matrices (reverse order): returns the weight Jacobians as fp32
    tensors of the arguments' shapes (tuple over ``argnums``; with ``has_aux`` also the function value).
    ``precision``: "bf16" (operands of the contractions rounded to bf16) or "fp32" (three-term bf16 split, float32
    accuracy); default: "bf16" only if every matrix argument is a bf16 tensor."""
    shapes = tuple(tuple(a.shape) for a in args)
    if not (order in ("rev", "reverse")):
        # Other orders at tensor level would carry edges such as J[z2 | W1] = [B, O | H, I] - O(batch x parameters)
        # values per edge, which is why forward mode over weight matrices is 11x the multiplications already at the
        # reference's 6-10-4 toy (33 732 vs 3 096).  Small problems run on the per-sample engine, which honours any
        # order on matrix-valued vertices (scalarised entries, the reference's structure and count_ops); large ones
        # are refused rather than attempted.
        total = sum(int(np.prod(sh, dtype=np.int64)) if sh else 1 for sh in shapes)
        if total > 4096:
            raise NotImplementedError("the tensor-valued path eliminates in reverse order only; other orders are "
                                      f"supported up to 4096 input scalars (this call has {total})")
        from .api import jacve
        out = jacve(fun, order, argnums, has_aux=has_aux, device=device, frontend=frontend)(*args)
        primal, grads = out if has_aux else (None, out)
        grads = grads if isinstance(grads, tuple) and len(argnums) > 1 else (grads if isinstance(grads, tuple) else (grads,))
        return (primal, tuple(grads)) if has_aux else tuple(grads)
    precision = precision or default_precision(args)
    key = (fun, str(order), tuple(argnums), shapes, device, has_aux, precision)
    if key not in _cache:
        plan = build_dense_plan(make_jaxpr(fun, shapes, frontend), order, argnums)
        _cache[key] = DenseExecutor(plan, device, with_primal=has_aux, precision=precision)
    ex = _cache[key]
    grads, primal = ex.run(args)
    return (primal, tuple(grads)) if has_aux else tuple(grads)
