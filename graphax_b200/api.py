"""``jacve`` - the reference's operator API on top of the B200 engine.

Mirrors ``graphax.jacve(fun, order, argnums=(0,), has_aux=False, count_ops=False,
sparse_representation=False)`` (reference ``core.py:27-93``): same argument
meaning, same result pytree (a tuple per output of a tuple per argnum; a single
output with several arguments is unwrapped, ``core.py:79-80``), same exceptions
for bad orders (``ValueError``, ``core.py:302,310``) and unknown primitives
(``NotImplementedError``, ``core.py:398``).

What changes is when the work happens.  The reference re-walks the jaxpr and
emits hundreds of jnp ops on every trace; here the first call for a given
(shapes, order, argnums) lowers the elimination once into a plan, uploads it,
and every call after that is one kernel launch.  Batching follows the scalable
form ``jax.vmap(jacve(f, ...))``: ``graphax_b200.vmap(jacfun)``
or ``jacfun.batched`` take arguments with a leading batch axis.

Arrays are ``torch`` CUDA tensors (float32); Python numbers, lists and NumPy
arrays are uploaded.  There is no CPU execution path.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import wraps
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .plan import Plan, build_plan
from .tracer import Jaxpr, Literal, make_jaxpr

EliminationOrder = Union[Sequence[int], str]


def tree_allclose(tree1, tree2, equal_nan: bool = False) -> bool:
    """Reference ``core.py:17-20``: allclose(atol=1e-5, rtol=1e-4) over a pytree."""
    import torch

    def leaves(t):
        if isinstance(t, (tuple, list)):
            for x in t:
                yield from leaves(x)
        else:
            yield t
    l1, l2 = list(leaves(tree1)), list(leaves(tree2))
    if len(l1) != len(l2):
        return False
    for a, b in zip(l1, l2):
        a = torch.as_tensor(a, dtype=torch.float32).cpu()
        b = torch.as_tensor(b, dtype=torch.float32).cpu()
        if a.shape != b.shape or not torch.allclose(a, b, atol=1e-5, rtol=1e-4, equal_nan=equal_nan):
            return False
    return True


@dataclass
class SparseJacobian:
    """``sparse_representation=True`` result: the structure descriptor graphax's
    SparseTensor would carry (``sparse/tensor.py:44-85``) plus the dense block."""
    out_dims: list
    primal_dims: list
    val_shape: Optional[list]
    dense: Any


def sparse_tensor_zeros_like(st: SparseJacobian) -> SparseJacobian:
    """Reference ``sparse/tensor.py:190-201`` (exported from ``graphax/__init__.py:4``): the same structure with an
    all-zero value."""
    import torch
    dense = torch.zeros_like(st.dense) if isinstance(st.dense, torch.Tensor) else np.zeros_like(st.dense)
    return SparseJacobian([dict(d) if isinstance(d, dict) else d for d in st.out_dims],
                          [dict(d) if isinstance(d, dict) else d for d in st.primal_dims],
                          None if st.val_shape is None else list(st.val_shape), dense)


def get_output_vertices(fun: Callable, arg_shapes: Sequence[Tuple[int, ...]], frontend: str = "auto") -> set:
    """Reference ``utils.py:8-27``: the output vertices that feed other vertices (``vo_vertices``, core.py:335-337,
    404-409) - outputs that may, and in a custom order must, be eliminated.  Takes the function and its per-sample
    argument shapes where the reference takes a jaxpr."""
    from .elimination import build_graph
    return set(build_graph(make_jaxpr(fun, tuple(tuple(s) for s in arg_shapes), frontend)).vo_vertices)


def get_valid_vertices(fun: Callable, arg_shapes: Sequence[Tuple[int, ...]], frontend: str = "auto") -> List[int]:
    """Reference ``utils.py:30-36``: the vertex ids a custom elimination order has to contain (every equation that is
    not a pure output), in ascending order = the ``"fwd"`` order (core.py:275-311)."""
    from .elimination import build_graph, eliminable_vertices
    return eliminable_vertices(build_graph(make_jaxpr(fun, tuple(tuple(s) for s in arg_shapes), frontend)))


class JacveFunction:
    """The callable ``jacve`` returns. ``__call__`` = one sample, ``batched`` = leading batch axis."""

    def __init__(self, fun: Callable, order: EliminationOrder, argnums: Sequence[int], has_aux: bool,
                 count_ops: bool, sparse_representation: bool, order_numbering: str, device: Optional[int],
                 jit: bool, contract_sums: bool = False, frontend: str = "auto", rounding: Optional[str] = None,
                 precision: Optional[str] = None):
        from .plan import DEFAULT_ROUNDING
        if precision not in (None, "bf16", "fp32"):
            raise ValueError("precision must be 'bf16', 'fp32' or None")
        self.precision = precision
        self.fun = fun
        self.frontend = frontend
        self.rounding = ("fma" if contract_sums else DEFAULT_ROUNDING) if rounding is None else rounding
        if self.rounding not in ("fma", "reference"):
            raise ValueError("rounding must be 'fma' or 'reference'")
        if contract_sums and self.rounding == "reference":
            raise ValueError("contract_sums changes the association of the edge sums: use rounding='fma' with it")
        self.contract_sums = bool(contract_sums)
        self.order = order if isinstance(order, str) else [int(o) for o in order]
        self.argnums = tuple(int(a) for a in (argnums if isinstance(argnums, (tuple, list, range)) else (argnums,)))
        self.has_aux = has_aux
        self.count_ops = count_ops
        self.sparse_representation = sparse_representation
        if order_numbering not in ("per_sample", "vmap"):
            raise ValueError("order_numbering must be 'per_sample' or 'vmap'")
        self.order_numbering = order_numbering
        self.device = device
        self.jit = jit
        self._plans: Dict[Tuple, Plan] = {}
        self._device_plans: Dict[Tuple, Any] = {}
        self._elementwise_cache: Dict[Tuple, Any] = {}
        self._package_geometry: Dict[Tuple, Any] = {}
        wraps(fun)(self)

    # ---- host side: trace + lower once per shape signature -------------------
    def lower(self, arg_shapes: Sequence[Tuple[int, ...]]) -> Plan:
        """Trace ``fun`` on per-sample shapes and build the elimination plan (no GPU needed)."""
        key = tuple(tuple(s) for s in arg_shapes)
        plan = self._plans.get(key)
        if plan is None:
            jaxpr = make_jaxpr(self.fun, key, self.frontend)
            order = self.order
            if not isinstance(order, str) and self.order_numbering == "vmap":
                order = jaxpr.translate_vmap_order(order)
            plan = build_plan(jaxpr, order, self.argnums, with_primal=True, contract=self.contract_sums,
                              rounding=self.rounding)
            plan.jaxpr = jaxpr
            self._plans[key] = plan
        return plan

    def _device_plan(self, plan: Plan, device: int):
        from .runtime import DevicePlan
        key = (plan.key, device)
        dp = self._device_plans.get(key)
        if dp is None:
            dp = DevicePlan(plan, device, jit=self.jit)
            self._device_plans[key] = dp
        return dp

    # ---- device side ---------------------------------------------------------
    def _prepare(self, args: Sequence[Any], batched: bool):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("graphax_b200 needs a CUDA device: there is no CPU fallback for jacve")
        device = self.device
        if device is None:
            dev_args = [a.device.index for a in args if isinstance(a, torch.Tensor) and a.is_cuda]
            device = dev_args[0] if dev_args else torch.cuda.current_device()
        tensors = []
        for a in args:
            if not isinstance(a, torch.Tensor):
                # device arrays of other libraries (jax.Array, cupy) come in through DLPack without a copy
                if hasattr(a, "__dlpack__") and not isinstance(a, np.ndarray):
                    try:
                        a = torch.from_dlpack(a)
                    except Exception:
                        a = torch.as_tensor(np.asarray(a, dtype=np.float32))
                else:
                    a = torch.as_tensor(np.asarray(a, dtype=np.float32))
            tensors.append(a.to(device=f"cuda:{device}", dtype=torch.float32).contiguous())
        if batched:
            batch = tensors[0].shape[0]
            if any(t.ndim < 1 or t.shape[0] != batch for t in tensors):
                raise ValueError("batched call: every argument needs the same leading batch axis")
            shapes = [tuple(t.shape[1:]) for t in tensors]
        else:
            batch = 1
            shapes = [tuple(t.shape) for t in tensors]
        return tensors, shapes, batch, device

    def _run_dense(self, args: Sequence[Any]):
        """Matrix-valued arguments: tensor-level reverse elimination with tcgen05 contractions (dense.py)."""
        import torch
        from .dense import dense_jacve
        if not torch.cuda.is_available():
            raise RuntimeError("graphax_b200 needs a CUDA device: there is no CPU fallback for jacve")
        if self.count_ops or self.sparse_representation:
            raise NotImplementedError("count_ops / sparse_representation for tensor-valued vertices")
        device = self.device if self.device is not None else torch.cuda.current_device()
        tensors = [a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=np.float32)) for a in args]
        tensors = [t.to(f"cuda:{device}") for t in tensors]
        out = dense_jacve(self.fun, self.order, self.argnums, *tensors, has_aux=self.has_aux, device=device,
                          precision=self.precision, frontend=self.frontend)
        if self.has_aux:
            primal, grads = out
            return primal, (grads if len(self.argnums) > 1 else grads[0])
        return out if len(self.argnums) > 1 else out[0]

    _ELEMENTWISE = {"add", "sub", "mul", "div", "atan2", "pow", "max", "min", "integer_pow", "square", "neg", "abs",
                    "sqrt", "sin", "cos", "tan", "exp", "log", "log1p", "tanh", "sinh", "cosh", "asin", "acos", "atan",
                    "asinh", "acosh", "atanh", "erf", "logistic"}

    def _elementwise_scalar_jaxpr(self, shapes):
        """If ``fun`` is elementwise over arguments of one common shape (every equation maps that shape to
        itself; other operands are literals), its Jacobian blocks are diagonal - in the reference every edge
        is a pair of SparseDimensions sharing one val (primitives.py:42-107) - and their values are the
        Jacobians of the scalar function, one per element.  Returns the scalar function's jaxpr, else None."""
        key = tuple(tuple(s) for s in shapes)
        if key in self._elementwise_cache:
            return self._elementwise_cache[key]
        self._elementwise_cache[key] = result = self._detect_elementwise(key)
        return result

    def _detect_elementwise(self, shapes):
        common = shapes[0]
        if not common or any(tuple(s) != tuple(common) for s in shapes):
            return None
        try:
            full = make_jaxpr(self.fun, shapes, self.frontend)
            scalar = make_jaxpr(self.fun, [()] * len(shapes), self.frontend)
        except Exception:
            return None
        if len(full.eqns) != len(scalar.eqns) or len(full.outvars) != len(scalar.outvars):
            return None
        if full.constvars or scalar.constvars:
            return None                 # closed-over arrays are per-element data: not a function of scalars alone
        for e, f in zip(full.eqns, scalar.eqns):
            if e.primitive != f.primitive or e.primitive not in self._ELEMENTWISE or e.outvar.shape != tuple(common):
                return None
            if any(hasattr(a, "shape") and not isinstance(a, Literal) and a.shape != tuple(common) for a in e.invars):
                return None
        if any(o.shape != tuple(common) for o in full.outvars):
            return None
        return scalar

    def _run_elementwise(self, args: Sequence[Any]):
        """Elementwise functions of large tensors (reference test_Simple, 50x50 arguments): the elements are
        the batch of the per-sample engine; the diagonal values are scattered into ``out.shape + in.shape``
        blocks (or returned as the SparseTensor-like descriptor with ``sparse_representation=True``)."""
        import torch
        tensors, shapes, _, device = self._prepare(args, batched=False)
        common = tuple(shapes[0])
        n = int(np.prod(common, dtype=np.int64))
        plan = self.lower([()] * len(shapes))
        dp = self._device_plan(plan, device)
        with torch.cuda.device(device):
            jac = torch.empty((n, plan.n_rows, plan.n_cols), dtype=torch.float32, device=f"cuda:{device}")
            primal = torch.empty((n, plan.n_rows), dtype=torch.float32, device=f"cuda:{device}") if self.has_aux else None
            dp.launch(n, [t.reshape(n).data_ptr() for t in tensors], jac.data_ptr(),
                      primal.data_ptr() if primal is not None else None, torch.cuda.current_stream(device).cuda_stream)
            nd = len(common)
            if not self.sparse_representation:       # one zero fill and one strided copy for all blocks
                dense = torch.zeros((plan.n_rows, plan.n_cols, n, n), dtype=torch.float32, device=f"cuda:{device}")
                dense.diagonal(dim1=2, dim2=3).copy_(jac.permute(1, 2, 0))
            blocks = []
            for oi in range(plan.n_rows):
                row = []
                for ai in range(len(self.argnums)):
                    if self.sparse_representation:
                        dims_o = [["sparse", d, common[d], d, nd + d] for d in range(nd)]
                        dims_p = [["sparse", nd + d, common[d], d, d] for d in range(nd)]
                        row.append(SparseJacobian(dims_o, dims_p, list(common), jac[:, oi, ai].reshape(common)))
                    else:
                        row.append(dense[oi, ai].reshape(common + common))
                blocks.append(row)
        grads_tree, single = self._tree(blocks, len(shapes))
        if self.count_ops:
            # every scalar multiplication / addition of the per-element plan happens once per element
            # (get_num_muls / get_num_adds of diagonal edges count the val size, tensor.py:1414-1493)
            return grads_tree, {"num_muls": plan.stats["num_muls"] * n, "num_adds": plan.stats["num_adds"] * n,
                                "order_counts": [tuple(int(c) * n if i else c for i, c in enumerate(t))
                                                 for t in plan.stats["order_counts"]]}
        if self.has_aux:
            outs = [primal[:, oi].reshape(common) for oi in range(plan.n_rows)]
            return (outs[0] if single else tuple(outs)), grads_tree
        return grads_tree

    def _run(self, args: Sequence[Any], batched: bool):
        import torch
        # (the reference's own timing scripts run RoeFlux_1d / RobotArm_6DOF elementwise on (600,) / (1000,) arrays,
        # performance_measurements.py:105-154)
        if not batched and any(getattr(a, "ndim", 0) >= 1 and int(np.prod(a.shape)) >= 64 for a in args):
            shapes = [tuple(getattr(a, "shape", ())) for a in args]
            if self._elementwise_scalar_jaxpr(shapes) is not None:
                return self._run_elementwise(args)
        if not batched and any(getattr(a, "ndim", 0) >= 2 and int(np.prod(a.shape)) >= 4096 for a in args):
            return self._run_dense(args)        # big matrices: tensor-level path; small ones are scalarised per sample
        tensors, shapes, batch, device = self._prepare(args, batched)
        plan = self.lower(shapes)
        dp = self._device_plan(plan, device)
        with torch.cuda.device(device):
            jac = torch.empty((batch, plan.n_rows, plan.n_cols), dtype=torch.float32, device=f"cuda:{device}")
            primal = torch.empty((batch, plan.n_rows), dtype=torch.float32, device=f"cuda:{device}") \
                if self.has_aux else None
            stream = torch.cuda.current_stream(device).cuda_stream
            dp.launch(batch, [t.data_ptr() for t in tensors], jac.data_ptr(),
                      primal.data_ptr() if primal is not None else None, stream)
        return self._package(plan, jac, primal, shapes, batch, batched)

    def _package(self, plan: Plan, jac, primal, shapes, batch: int, batched: bool):
        """Slice the packed tile into the reference's pytree (core.py:555-568, 76-92)."""
        jaxpr: Jaxpr = plan.jaxpr
        lead = (batch,) if batched else ()
        out_shapes = [v.shape for v in jaxpr.outvars]
        # every block is one strided view of the packed tile: row range of the output x column range of the argument,
        # split into the output's and the argument's own dimensions (geometry cached per plan and argument shapes: a
        # host-array call spends more time on this packaging than on anything else it does in Python)
        key = (id(plan), tuple(map(tuple, shapes)))
        geom = self._package_geometry.get(key)
        if geom is None:
            rows = [int(v) for v in np.concatenate([[0], np.cumsum(plan.out_sizes)])]
            cols = [int(v) for v in np.concatenate([[0], np.cumsum([plan.in_sizes[a] for a in self.argnums])])]

            def dense_strides(shape, unit):
                st, acc = [], unit
                for d in reversed(shape):
                    st.append(acc)
                    acc *= int(d)
                return tuple(reversed(st))
            blocks = {}
            for oi, oshape in enumerate(out_shapes):
                for ai, a in enumerate(self.argnums):
                    blocks[oi, ai] = (tuple(oshape) + tuple(shapes[a]),
                                      dense_strides(oshape, plan.n_cols) + dense_strides(shapes[a], 1),
                                      rows[oi] * plan.n_cols + cols[ai])
            geom = self._package_geometry[key] = (rows, cols, blocks, plan)     # the plan stays alive with its id
        rows, cols, blocks, _ = geom
        tile = plan.n_rows * plan.n_cols
        is_numpy = isinstance(jac, np.ndarray)              # host-side checks hand the packed tile over as an array
        base = 0 if is_numpy else jac.storage_offset()
        grads = []
        for oi, oshape in enumerate(out_shapes):
            row = []
            for ai, a in enumerate(self.argnums):
                shape, strides, off = blocks[oi, ai]
                if is_numpy:
                    block = (jac if batched else jac[0])[..., rows[oi]:rows[oi + 1], cols[ai]:cols[ai + 1]].reshape(lead + shape)
                else:
                    block = jac.as_strided(lead + shape, ((tile,) if batched else ()) + strides, base + off)
                if self.sparse_representation:
                    d = plan.structure["blocks"][f"{oi},{ai}"]
                    block = None if d is None else SparseJacobian(d["out_dims"], d["primal_dims"], d["val_shape"], block)
                row.append(block)
            grads.append(tuple(row) if len(self.argnums) > 1 else row[0])
        # core.py:76-92: always a tuple over the outputs, except that a single output of a function of
        # several arguments is unwrapped (so 1 output / 1 argument gives a 1-tuple, like the reference)
        single = len(out_shapes) == 1 and len(jaxpr.invars) > 1
        grads_tree = grads[0] if single else tuple(grads)
        if self.count_ops:
            if plan.stats["num_muls"] is None:
                # an edge outside the reference's structure rules (the reference itself raises there, e.g. a
                # concatenation of a replicated dimension): report the structurally non-zero scalar products instead
                aux = {"num_muls": plan.stats["scalar_muls"], "num_adds": plan.stats["scalar_adds"],
                       "order_counts": list(plan.stats["scalar_order_counts"]), "model": "scalar products"}
                return grads_tree, aux
            aux = {"num_muls": plan.stats["num_muls"], "num_adds": plan.stats["num_adds"],
                   "order_counts": list(plan.stats["order_counts"])}
            return grads_tree, aux
        if self.has_aux:
            pt = primal if batched else primal[0]
            outs = [pt[..., rows[oi]:rows[oi + 1]].reshape(lead + tuple(s)) for oi, s in enumerate(out_shapes)]
            primal_tree = outs[0] if single else tuple(outs)
            return primal_tree, grads_tree
        return grads_tree

    def _run_host(self, args: Sequence[Any], out=None):
        """Batched call on HOST arrays: pinned staging, chunked H2D -> kernel -> D2H (gx_jacve_host).
        Results come back as (pinned) CPU tensors, like a jitted function called with NumPy arrays.
        ``out`` may hold a reusable ``[batch, n_rows, n_cols]`` float32 CPU tensor (pinned for full copy
        bandwidth): page-locking a fresh 200 MB result on every call costs ~20 ms, 5x the transfer."""
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("graphax_b200 needs a CUDA device: there is no CPU fallback for jacve")
        device = self.device if self.device is not None else torch.cuda.current_device()
        tensors = [a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
                   for a in args]
        tensors = [t.to(torch.float32).contiguous() for t in tensors]
        batch = tensors[0].shape[0]
        if any(t.ndim < 1 or t.shape[0] != batch for t in tensors):
            raise ValueError("batched call: every argument needs the same leading batch axis")
        shapes = [tuple(t.shape[1:]) for t in tensors]
        plan = self.lower(shapes)
        dp = self._device_plan(plan, device)
        pinned = all(t.is_pinned() for t in tensors)
        if out is not None:
            if (not isinstance(out, torch.Tensor) or out.is_cuda or out.dtype != torch.float32
                    or tuple(out.shape) != (batch, plan.n_rows, plan.n_cols) or not out.is_contiguous()):
                raise ValueError(f"out must be a contiguous float32 CPU tensor of shape "
                                 f"{(batch, plan.n_rows, plan.n_cols)}")
            jac, pinned = out, pinned and out.is_pinned()
        else:
            jac = torch.empty((batch, plan.n_rows, plan.n_cols), dtype=torch.float32, pin_memory=pinned)
        primal = torch.empty((batch, plan.n_rows), dtype=torch.float32, pin_memory=pinned) if self.has_aux else None
        dp.run_host(batch, [t.data_ptr() for t in tensors], jac.data_ptr(),
                    primal.data_ptr() if primal is not None else None, pinned)
        return self._package(plan, jac, primal, shapes, batch, True)

    def _tree(self, blocks, n_invars: int):
        """core.py:76-92: tuple over outputs of tuple over argnums; a single output of a function of several
        arguments is unwrapped, a single argnum gives the bare block."""
        grads = [tuple(row) if len(self.argnums) > 1 else row[0] for row in blocks]
        single = len(blocks) == 1 and n_invars > 1
        return (grads[0] if single else tuple(grads)), single

    def _run_traced(self, args: Sequence[Any]):
        """Nested use (reference economics.py:21-22, BlackScholes_Jacobian): called on traced values while an
        outer function is being lowered.  The inner elimination plan is replayed operation by operation on the
        outer trace - every plan operation becomes one outer equation on rank-0 values - so the outer jacve
        differentiates through the inner Jacobian accumulation (second-order derivatives).  The inner plan uses
        the reference's division formulas (``division="ieee"``): one equation per division to differentiate."""
        from .tracer import Tracer, jnp as tjnp
        if self.sparse_representation:
            raise NotImplementedError("sparse_representation inside a traced function")
        shapes = [tuple(a.shape) if isinstance(a, Tracer) else tuple(np.shape(a)) for a in args]
        jaxpr = make_jaxpr(self.fun, shapes, self.frontend)
        order = self.order
        if not isinstance(order, str) and self.order_numbering == "vmap":
            order = jaxpr.translate_vmap_order(order)
        plan = build_plan(jaxpr, order, self.argnums, with_primal=True, division="ieee")
        comps: List[Any] = []
        for a, shape in zip(args, shapes):
            size = int(np.prod(shape, dtype=np.int64)) if shape else 1
            if not shape:
                comps.append(a if isinstance(a, Tracer) else float(a))
            elif isinstance(a, Tracer):
                flat = a.reshape(size)
                comps.extend(flat[i] for i in range(size))
            else:
                comps.extend(float(v) for v in np.asarray(a, dtype=np.float32).reshape(-1))
        dag = plan.dag
        unary = {"neg": lambda x: -x, "abs": tjnp.abs, "sqrt": tjnp.sqrt, "sin": tjnp.sin, "cos": tjnp.cos,
                 "tan": tjnp.tan, "exp": tjnp.exp, "log": tjnp.log, "log1p": tjnp.log1p, "tanh": tjnp.tanh,
                 "sinh": tjnp.sinh, "cosh": tjnp.cosh, "asin": tjnp.arcsin, "acos": tjnp.arccos,
                 "atan": tjnp.arctan, "asinh": tjnp.arcsinh, "acosh": tjnp.arccosh, "atanh": tjnp.arctanh,
                 "erf": tjnp.erf, "logistic": tjnp.sigmoid}
        binary = {"add": lambda x, y: x + y, "sub": lambda x, y: x - y, "mul": lambda x, y: x * y,
                  "div": lambda x, y: x / y, "atan2": tjnp.arctan2, "pow": tjnp.power, "max": tjnp.maximum,
                  "min": tjnp.minimum, "gt": tjnp.greater, "lt": tjnp.less, "eq": tjnp.equal}
        vals: Dict[int, Any] = {}

        def value(node: int):
            kind = dag.op[node]
            if kind == "const":
                return float(dag.cval[node])
            if kind == "input":
                return comps[dag.args[node][0]]
            return vals[node]

        tile: List[Any] = [0.0] * (plan.n_rows * plan.n_cols)
        primal: List[Any] = [0.0] * plan.n_rows
        for op in plan.ssa:
            xs = [value(a) for a in op.args]
            if op.op == "store_jac":
                tile[op.dst] = xs[0]
            elif op.op == "store_primal":
                primal[op.dst] = xs[0]
            elif not any(isinstance(x, Tracer) for x in xs):
                raise NotImplementedError(f"nested jacve: constant sub-expression ({op.op}) left to the device")
            elif op.op == "fma":
                vals[op.dst] = xs[0] * xs[1] + xs[2]
            elif op.op == "select":
                which = xs[0] if isinstance(xs[0], Tracer) else tjnp.zeros(()) + xs[0]
                vals[op.dst] = tjnp.where(which > 0.5, xs[1], xs[2])
            elif op.op in unary:
                vals[op.dst] = unary[op.op](xs[0])
            else:
                vals[op.dst] = binary[op.op](xs[0], xs[1])

        def assemble(entries: List[Any], shape: Tuple[int, ...]):
            if not shape:
                return entries[0]
            parts = [e.reshape(1) if isinstance(e, Tracer) else tjnp.zeros((1,)) + e for e in entries]
            return tjnp.concatenate(parts).reshape(shape) if len(parts) > 1 else parts[0].reshape(shape)

        rows = np.concatenate([[0], np.cumsum(plan.out_sizes)]).astype(int)
        cols = np.concatenate([[0], np.cumsum([plan.in_sizes[a] for a in self.argnums])]).astype(int)
        blocks = []
        for oi, o in enumerate(jaxpr.outvars):
            row = []
            for ai, a in enumerate(self.argnums):
                entries = [tile[r * plan.n_cols + c] for r in range(rows[oi], rows[oi + 1])
                           for c in range(cols[ai], cols[ai + 1])]
                row.append(assemble(entries, tuple(o.shape) + tuple(shapes[a])))
            blocks.append(row)
        grads_tree, single = self._tree(blocks, len(jaxpr.invars))
        if self.count_ops:
            if plan.stats["num_muls"] is None:
                raise NotImplementedError("count_ops needs tensor-level edge structure (rank <= 1 vertices)")
            return grads_tree, {"num_muls": plan.stats["num_muls"], "num_adds": plan.stats["num_adds"]}
        if self.has_aux:
            outs = [assemble(primal[rows[oi]:rows[oi + 1]], tuple(o.shape)) for oi, o in enumerate(jaxpr.outvars)]
            return (outs[0] if single else tuple(outs)), grads_tree
        return grads_tree

    def __call__(self, *args):
        from .tracer import Tracer
        if any(isinstance(a, Tracer) for a in args):
            return self._run_traced(args)
        return self._run(args, batched=False)

    def batched(self, *args, out=None):
        """Leading batch axis on every argument. CUDA tensors stay on the device (one launch on the
        current stream); host arrays (NumPy / CPU tensors) go through the pipelined host path, where
        ``out=`` can supply a reusable packed result buffer."""
        import torch
        on_host = all(isinstance(a, np.ndarray) or (isinstance(a, torch.Tensor) and not a.is_cuda) for a in args)
        if on_host and len(args) > 0:
            return self._run_host(args, out=out)
        if out is not None:
            raise ValueError("out= is only supported for host (CPU) arguments")
        return self._run(args, batched=True)


def jacve(fun: Callable, order: EliminationOrder, argnums: Sequence[int] = (0,), has_aux: bool = False,
          count_ops: bool = False, sparse_representation: bool = False, *, order_numbering: str = "per_sample",
          device: Optional[int] = None, jit: bool = True, contract_sums: bool = False,
          frontend: str = "auto", rounding: Optional[str] = None, precision: Optional[str] = None) -> JacveFunction:
    """Jacobian of ``fun`` w.r.t. ``argnums`` by vertex elimination in the given ``order``.

    ``order``: ``"fwd"``/``"forward"``, ``"rev"``/``"reverse"`` or a sequence of 1-based vertex ids
    (positions of the equations of the traced function, as in the reference).  ``order_numbering="vmap"``
    accepts ids written against ``jacve(jax.vmap(fun))`` (reference ``example_test.py:150-166``).
    ``rounding``: ``"reference"`` evaluates the reference's own sequence of rounded float32 operations (separate
    multiply and add per ``J[s,p] += J[s,v] J[v,p]``, IEEE divisions): bit-identical to the unfused reference
    algorithm; ``"fma"`` contracts product-and-add pairs into FMAs and divides through one reciprocal (fewer
    instructions, a few ulp per operation away from the reference).  Default: ``plan.DEFAULT_ROUNDING``.
    ``precision`` applies to the tensor-valued path only (non-batched calls with matrix arguments of >= 4096
    elements, ``dense.py``): ``"fp32"`` contracts float32 operands to float32 accuracy (three-term bf16 split on the
    tensor cores), ``"bf16"`` rounds the operands of every contraction to bf16 (BASELINE config 5; error up to 3e-2
    of max|J|).  ``None``: "bf16" only when every matrix argument already is a bf16 tensor.  That path supports
    ``order="rev"``, a scalar output, ``has_aux``; ``count_ops`` / ``sparse_representation`` raise there.
    ``contract_sums=True`` trades the reference's association of the edge sums for 6-12 % fewer instructions
    (``plan._contract_sums``); off by default so results follow the reference's summation order.
    ``frontend``: ``"native"`` traces ``fun`` with the engine's own ``jnp`` stand-in, ``"jax"`` with
    ``jax.make_jaxpr`` (functions written against the real ``jax.numpy``; needs JAX), ``"auto"`` tries both.
    """
    return JacveFunction(fun, order, argnums, has_aux, count_ops, sparse_representation, order_numbering,
                         device, jit, contract_sums, frontend, rounding, precision)


def vmap(jacfun: JacveFunction, in_axes: int = 0) -> Callable:
    """``jax.vmap(jacve(f, ...))`` for the engine: the batch becomes the CUDA grid."""
    if not isinstance(jacfun, JacveFunction):
        raise TypeError("graphax_b200.vmap only batches functions returned by graphax_b200.jacve")
    if in_axes == 0:
        return jacfun.batched
    if not isinstance(in_axes, (tuple, list)) or any(a not in (0, None) for a in in_axes):
        raise NotImplementedError("in_axes must be 0 or a tuple of 0 / None (leading batch axis or shared argument)")
    axes = tuple(in_axes)

    def batched(*args, out=None):
        """Shared (in_axes None) arguments - e.g. the weights in ``vmap(f, in_axes=(0, None, ...))`` of the
        reference's vmap_NeuralNetwork test - are broadcast along the batch: every sample's Jacobian block
        with respect to them has shape ``[batch, *out, *arg.shape]``."""
        import torch
        if len(args) != len(axes):
            raise ValueError(f"in_axes has {len(axes)} entries for {len(args)} arguments")
        lead = [a for a, ax in zip(args, axes) if ax == 0]
        if not lead:
            raise ValueError("at least one argument must carry the batch axis")
        batch = lead[0].shape[0]
        full = []
        for a, ax in zip(args, axes):
            if ax is None:
                t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=np.float32))
                if isinstance(lead[0], torch.Tensor):
                    t = t.to(lead[0].device)
                a = t.expand((batch,) + tuple(t.shape)).contiguous()
                if not isinstance(lead[0], torch.Tensor):
                    a = a.numpy()
            full.append(a)
        return jacfun.batched(*full, out=out)
    return batched
