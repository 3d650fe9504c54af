"""ORACLE - test infrastructure, not product code.

Reverse-order vertex elimination at tensor level for graphs with matrix-valued vertices (config 5), restated
with NumPy.  The cotangent edge ``() | vertex-shape`` is a dense array; each product with an in-edge follows
the reference's product classes: elementwise edges -> broadcast multiply (``sparse/tensor.py:826-866``),
``dot_general`` edges (``primitives.py:349-441``: Kronecker x operand) -> ``_mixed_mul`` = ``lax.dot_general``
over the dense pair (``tensor.py:1062-1231``), broadcast edges -> sum over the broadcast axis
(``primitives.py:763-883``), fill-in -> addition (``tensor.py:1267-1411``).

PINNED AGAINST A RUN OF THE REFERENCE'S OWN CODE (tests/test_dense.py::test_dense_oracle_matches_reference_run):
the unmodified reference sources on the NumPy stand-in for JAX (tests/golden/make_reference_goldens.py,
``run_mlp_batched``: the batched MLP of the reference's test_vmap_NeuralNetwork at widths 6 -> 10 -> 4, batch 12)
give the same loss bit for bit and weight Jacobians within 3e-7 of max|J| in forward AND reverse mode; also pinned
against torch.autograd in fp64 (1e-12).  XLA's own arithmetic stays unpinned (no JAX in the image), and so does
bf16: no reference test uses it.  ``bf16_operands`` emulates the engine's precision policy: operands of every contraction
are rounded to bfloat16, accumulation stays in the working dtype.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def _round_bf16(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16 precision, result kept in the input dtype."""
    x = np.ascontiguousarray(a, dtype=np.float32)
    u = x.view(np.uint32)
    rounded = ((u + (0x7FFF + ((u >> 16) & 1))) & 0xFFFF0000).astype(np.uint32)
    return rounded.view(np.float32).astype(a.dtype)


def jacve_rev(jaxpr, argnums: Sequence[int], args: Sequence[np.ndarray], dtype=np.float64, bf16_operands: bool = False):
    from graphax_b200.tracer import Literal, Var
    q = _round_bf16 if bf16_operands else (lambda a: a)
    key = lambda v: -(v.input_index + 1) if v.input_index is not None else v.eqn_index + 1
    env: Dict[int, np.ndarray] = {key(v): np.asarray(a, dtype=dtype) for v, a in zip(jaxpr.invars, args)}
    read = lambda a: np.asarray(a.val, dtype=dtype) if isinstance(a, Literal) else env[key(a)]
    # forward sweep: primal values and elemental partials (core.py:380-410)
    for eqn in jaxpr.eqns:
        p, x = eqn.primitive, [read(a) for a in eqn.invars]
        if p == "dot_general":
            (ca, cb), _ = eqn.params["dimension_numbers"]
            out = np.tensordot(q(x[0]), q(x[1]), axes=(ca, cb)).astype(dtype)
        elif p == "broadcast_in_dim":
            out = x[0].reshape(eqn.params["shape"])
        elif p == "reduce_sum":
            out = x[0].sum(axis=tuple(eqn.params["axes"]), dtype=dtype)
        elif p == "integer_pow":
            out = x[0] ** int(eqn.params["y"])
        elif p in ("add", "sub", "mul", "div"):
            out = {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide}[p](x[0], x[1])
        else:
            out = {"tanh": np.tanh, "exp": np.exp, "log": np.log, "sqrt": np.sqrt, "sin": np.sin, "cos": np.cos,
                   "neg": np.negative, "logistic": lambda v: 1. / (1. + np.exp(-v))}[p](x[0])
        env[key(eqn.outvar)] = np.asarray(out, dtype=dtype)

    cot: Dict[int, np.ndarray] = {key(jaxpr.outvars[0]): np.ones((), dtype=dtype)}

    def accumulate(v, contrib):
        shape = v.shape
        contrib = np.asarray(contrib, dtype=dtype)
        while contrib.ndim > len(shape):
            contrib = contrib.sum(axis=0)
        for ax, (cs, s) in enumerate(zip(contrib.shape, shape)):
            if cs != s:
                contrib = contrib.sum(axis=ax, keepdims=True)
        k = key(v)
        cot[k] = contrib + cot[k] if k in cot else contrib

    for eqn in reversed(jaxpr.eqns):
        g = cot.pop(key(eqn.outvar), None)
        if g is None:
            continue
        p, x, out = eqn.primitive, [read(a) for a in eqn.invars], env[key(eqn.outvar)]
        g = np.broadcast_to(g, out.shape)
        for i, a in enumerate(eqn.invars):
            if not isinstance(a, Var):
                continue
            if p == "dot_general":
                (ca, cb), _ = eqn.params["dimension_numbers"]
                if i == 0:      # d/da: contract g with b over b's free axis
                    fb = 1 - cb[0]
                    c = np.tensordot(q(g), q(x[1]), axes=((1,), (fb,)))          # [M, K]
                    accumulate(a, c if ca[0] == 1 else c.T)
                else:
                    fa = 1 - ca[0]
                    c = np.tensordot(q(g), q(x[0]), axes=((0,), (fa,)))          # [N, K]
                    accumulate(a, c if cb[0] == 1 else c.T)
            elif p in ("broadcast_in_dim", "add"):
                accumulate(a, g if p == "add" else g.reshape(eqn.params["shape"]))
            elif p == "reduce_sum":
                accumulate(a, np.broadcast_to(g, a.shape))
            elif p == "sub":
                accumulate(a, g if i == 0 else -g)
            elif p == "mul":
                accumulate(a, g * x[1 - i])
            elif p == "div":
                accumulate(a, g * (1. / x[1] if i == 0 else -x[0] / (x[1] * x[1])))
            elif p == "integer_pow":
                y = int(eqn.params["y"])
                accumulate(a, g * (y * x[0] ** (y - 1)))
            else:
                part = {"tanh": lambda: 1. - out * out, "exp": lambda: out, "log": lambda: 1. / x[0],
                        "sqrt": lambda: .5 / out, "sin": lambda: np.cos(x[0]), "cos": lambda: -np.sin(x[0]),
                        "neg": lambda: -np.ones_like(out), "logistic": lambda: out * (1. - out)}[p]()
                accumulate(a, g * part)
    grads = [cot.get(key(jaxpr.invars[a]), np.zeros(jaxpr.invars[a].shape, dtype=dtype)) for a in argnums]
    return grads, env[key(jaxpr.outvars[0])]
