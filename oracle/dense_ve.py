"""ORACLE - test infrastructure, not product code.

Vertex elimination with fully densified edge Jacobians, any rank, batch-leading NumPy.  This is what the
reference computes once its SparseTensor compression is stripped away: ``_build_graph`` (core.py:314-412)
stores dJ[dst]/dJ[src] per equation, ``_eliminate_vertex`` (core.py:155-272) forms
``J[s,p] (+)= J[s,v] @ J[v,p]`` for every successor/predecessor pair in dict order.  Used to check the engine on
graphs whose vertices have rank > 1 (the reference's tests/core/primitive_test.py cases, its 4-8-4
NeuralNetwork test), where ``oracle/ve_numpy.py`` (rank <= 1, structure-carrying) does not apply.

Elemental Jacobians of operations that are linear in an operand (dot_general, transpose, reshape, slice,
concatenate, reduce_sum, broadcast, add/sub) are obtained by pushing the identity basis through the NumPy
operation itself; nonlinear elementwise partials use the reference's formulas (primitives.py:150-239,
275-346).  PARITY UNPINNED by a runnable reference; pinned with torch.autograd (tests/test_rank_n.py).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .ve_numpy import _UNARY_RULES, _ipow, checkify_order


def _apply(eqn, vals: List[np.ndarray], batched: List[bool]):
    """Evaluate the primitive with NumPy on arrays whose axis 0 is a batch (or basis) axis when batched[i]."""
    p, params = eqn.primitive, eqn.params
    lead = 1 if any(batched) else 0
    x = [v if b or not lead else v[None] for v, b in zip(vals, batched)]
    if p == "stop_gradient":
        return x[0]
    if p == "convert_element_type":
        return x[0]
    if p == "select_n":       # lax.select_n(which, case0, case1)
        nd = max(a.ndim for a in x)
        w, c0, c1 = [v.reshape(v.shape[:1] + (1,) * (nd - v.ndim) + v.shape[1:]) if lead else v for v in x]
        return np.where(w != 0, c1, c0)
    if p in ("gt", "lt", "eq", "ge", "le", "ne"):
        nd = max(a.ndim for a in x)
        a, b = [v.reshape(v.shape[:1] + (1,) * (nd - v.ndim) + v.shape[1:]) if lead else v for v in x]
        return {"gt": np.greater, "lt": np.less, "eq": np.equal, "ge": np.greater_equal, "le": np.less_equal,
                "ne": np.not_equal}[p](a, b).astype(np.result_type(a, b))
    if p in ("add", "sub", "mul", "div", "atan2", "pow", "max", "min"):
        nd = max(a.ndim for a in x)
        a, b = [v.reshape(v.shape[:1] + (1,) * (nd - v.ndim) + v.shape[1:]) if v.ndim > 1 or lead else v for v in x]
        if lead:      # align per-sample ranks (rank-0 operands broadcast against rank-n ones)
            a = a.reshape(a.shape[:1] + (1,) * (nd - a.ndim) + a.shape[1:])
            b = b.reshape(b.shape[:1] + (1,) * (nd - b.ndim) + b.shape[1:])
        return {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide, "atan2": np.arctan2,
                "pow": np.power, "max": np.maximum, "min": np.minimum}[p](a, b)
    if p == "dot_general":
        (ca, cb), (ba, bb) = params["dimension_numbers"]
        a, b = x
        letters = iter("abcdefghijklmnopqrstuvw")
        la, lb = [None] * (a.ndim - 1), [None] * (b.ndim - 1)
        for i, j in zip(ba, bb):
            la[i] = lb[j] = next(letters)
        for i, j in zip(ca, cb):
            la[i] = lb[j] = next(letters)
        for l in (la, lb):
            for i in range(len(l)):
                if l[i] is None:
                    l[i] = next(letters)
        out = [la[i] for i in ba] + [c for i, c in enumerate(la) if i not in ba and i not in ca] + \
              [c for i, c in enumerate(lb) if i not in bb and i not in cb]
        return np.einsum(f"z{''.join(la)},z{''.join(lb)}->z{''.join(out)}",
                         np.broadcast_to(a, (max(a.shape[0], b.shape[0]),) + a.shape[1:]),
                         np.broadcast_to(b, (max(a.shape[0], b.shape[0]),) + b.shape[1:]))
    (a,) = x[:1]
    if p == "transpose":
        return np.transpose(a, (0,) + tuple(q + 1 for q in params["permutation"]))
    if p in ("reshape", "squeeze"):
        return a.reshape(a.shape[:1] + tuple(eqn.outvar.shape))
    if p == "slice":
        return a[(slice(None),) + tuple(slice(s, l) for s, l in zip(params["start_indices"], params["limit_indices"]))]
    if p == "broadcast_in_dim":
        shape, bdims = tuple(params["shape"]), tuple(params["broadcast_dimensions"])
        view = [1] * len(shape)
        for d, ax in zip(a.shape[1:], bdims):
            view[ax] = d
        return np.broadcast_to(a.reshape(a.shape[:1] + tuple(view)), a.shape[:1] + shape)
    if p == "reduce_sum":
        return a.sum(axis=tuple(q + 1 for q in params["axes"]))
    if p == "reduce_max":
        return a.max(axis=tuple(q + 1 for q in params["axes"]))
    if p == "reduce_min":
        return a.min(axis=tuple(q + 1 for q in params["axes"]))
    if p == "concatenate":
        n = max(v.shape[0] for v in x)
        return np.concatenate([np.broadcast_to(v, (n,) + v.shape[1:]) for v in x], axis=int(params["dimension"]) + 1)
    if p == "integer_pow":
        return _ipow(a, int(params["y"]))
    if p == "square":
        return a * a
    if p == "erf":
        import math
        return np.vectorize(math.erf, otypes=[a.dtype])(a)
    return _UNARY_RULES[p][0](a)


_LINEAR_IN = {"dot_general": (0, 1), "transpose": (0,), "reshape": (0,), "squeeze": (0,), "slice": (0,),
              "broadcast_in_dim": (0,), "reduce_sum": (0,), "concatenate": None, "add": (0, 1), "sub": (0, 1)}


def _elemental_jacobians(eqn, vals, out, is_var, dtype):
    """dense [B, n_out, n_in] per traced operand."""
    B = out.shape[0]
    n_out = int(np.prod(out.shape[1:], dtype=np.int64)) if out.ndim > 1 else 1
    p = eqn.primitive
    jacs = []
    for i, a in enumerate(eqn.invars):
        if not is_var[i]:
            continue
        n_in = a.size
        if p in _LINEAR_IN and (_LINEAR_IN[p] is None or i in _LINEAR_IN[p]):
            # push the identity basis of operand i through the operation, other operands held at their values
            cols = []
            basis = np.eye(n_in, dtype=dtype).reshape((n_in,) + tuple(a.shape))
            for b in range(B):
                args = [basis if j == i else (np.zeros_like(v[b]) if (p in ("add", "sub", "concatenate") and is_var[j] is not None and j != i)
                                              else v[b]) for j, v in enumerate(vals)]
                batched = [j == i for j in range(len(vals))]
                if p in ("add", "sub", "concatenate"):
                    args = [basis if j == i else np.zeros_like(v[b]) for j, v in enumerate(vals)]
                res = _apply(eqn, args, batched)
                cols.append(np.asarray(res, dtype=dtype).reshape(n_in, n_out).T)
            jacs.append(np.stack(cols))
            continue
        x = [v for v in vals]
        bshape = out.shape
        def full(t):
            t = np.asarray(t, dtype=dtype)
            if t.ndim == 0:
                t = np.broadcast_to(t, (B,))
            t = t.reshape(t.shape[:1] + (1,) * (len(bshape) - t.ndim) + t.shape[1:])
            return np.broadcast_to(t, bshape)
        if p == "mul":
            part = full(x[1 - i])
        elif p == "div":
            part = full(1. / full(x[1])) if i == 0 else -full(x[0]) / (full(x[1]) * full(x[1]))
        elif p == "atan2":
            a2 = full(x[0]) ** 2 + full(x[1]) ** 2
            part = full(x[1]) / a2 if i == 0 else -full(x[0]) / a2
        elif p == "pow":
            part = full(x[1]) * np.power(full(x[0]), full(x[1]) - 1.) if i == 0 else np.log(full(x[0])) * out
        elif p in ("max", "min"):
            first = full(x[0]) > full(x[1]) if p == "max" else full(x[0]) < full(x[1])
            part = np.where(first, 1., 0.) if i == 0 else np.where(first, 0., 1.)
        elif p in ("gt", "lt", "eq", "ge", "le", "ne"):
            part = np.zeros(bshape, dtype=dtype)
        elif p == "select_n":       # the selected case passes its derivative, the predicate none
            w = full(x[0]) != 0
            part = np.zeros(bshape, dtype=dtype) if i == 0 else np.where(w, 0., 1.) if i == 1 else np.where(w, 1., 0.)
        elif p == "convert_element_type":
            part = np.ones(bshape, dtype=dtype)
        elif p == "integer_pow":
            y = int(eqn.params["y"])
            part = y * _ipow(x[0], y - 1)
        elif p == "square":
            part = 2. * x[0]
        elif p in ("reduce_max", "reduce_min"):
            axes = tuple(q + 1 for q in eqn.params["axes"])
            hit = (x[0] == np.expand_dims(out, axes)).astype(dtype)
            part_in = hit / hit.sum(axis=axes, keepdims=True)               # [B, *in_shape]
            onehot = np.zeros((n_in, n_out), dtype=dtype)
            idx_out = np.broadcast_to(np.expand_dims(np.arange(n_out).reshape(out.shape[1:]), tuple(q for q in eqn.params["axes"])),
                                      a.shape).reshape(-1)
            onehot[np.arange(n_in), idx_out] = 1.
            jacs.append(part_in.reshape(B, 1, n_in) * onehot.T[None])
            continue
        elif p == "erf":
            part = 2. * np.exp(-(x[0] * x[0])) / np.sqrt(np.pi)
        else:
            part = _UNARY_RULES[p][1](out, x[0])
        part = np.asarray(part, dtype=dtype).reshape(B, n_out)
        # which operand component does output component o read?  (broadcast of the operand to the out shape)
        sel = np.broadcast_to(np.arange(n_in).reshape((1,) * (len(bshape) - 1 - a.ndim) + tuple(a.shape)), bshape[1:]).reshape(-1)
        onehot = np.zeros((n_out, n_in), dtype=dtype)
        onehot[np.arange(n_out), sel] = 1.
        jacs.append(part[:, :, None] * onehot[None])
    return jacs


def jacve(jaxpr, order, argnums: Sequence[int], args: Sequence[np.ndarray], dtype=np.float64):
    """Same contract as ve_numpy.jacve: (jac[o][a] of shape [B, *out, *in], primal list)."""
    from graphax_b200.tracer import Literal, Var
    key = lambda v: -(v.input_index + 1) if v.input_index is not None else v.eqn_index + 1
    B = np.asarray(args[0]).shape[0]
    env: Dict[int, np.ndarray] = {key(v): np.asarray(a, dtype=dtype).reshape((B,) + tuple(v.shape))
                                  for v, a in zip(jaxpr.invars, args)}
    graph: Dict[int, Dict[int, np.ndarray]] = {key(v): {} for v in jaxpr.invars}
    tgraph: Dict[int, Dict[int, np.ndarray]] = {key(v): {} for v in jaxpr.invars}
    outkeys = {key(v) for v in jaxpr.outvars}
    vo = set()
    with np.errstate(all="ignore"):
        for eqn in jaxpr.eqns:
            v = key(eqn.outvar)
            graph.setdefault(v, {})
            tgraph.setdefault(v, {})
            is_var = [isinstance(a, Var) for a in eqn.invars]
            for a in eqn.invars:
                if isinstance(a, Var) and a.eqn_index is not None and key(a) in outkeys:
                    vo.add(key(a))
            vals = [env[key(a)] if isinstance(a, Var) else np.full((B,), a.val, dtype=dtype) for a in eqn.invars]
            out = np.asarray(_apply(eqn, vals, [True] * len(vals)), dtype=dtype).reshape((B,) + tuple(eqn.outvar.shape))
            env[v] = out
            if not any(is_var) or eqn.primitive == "stop_gradient":       # primitives.py:458-462: no edge
                continue
            for a, J in zip([a for a in eqn.invars if isinstance(a, Var)], _elemental_jacobians(eqn, vals, out, is_var, dtype)):
                src = key(a)
                J = J + graph[src][v] if v in graph[src] else J
                graph[src][v] = J
                tgraph[v][src] = J
        for vertex in checkify_order(order, jaxpr, vo):
            for s in list(graph[vertex].keys()):
                post = graph[vertex][s]
                for p in list(tgraph[vertex].keys()):
                    prod = np.matmul(post, tgraph[vertex][p])
                    if s in graph[p]:
                        prod = prod + graph[p][s]
                    graph[p][s] = prod
                    tgraph[s][p] = prod
            if vertex not in vo:
                for p in tgraph[vertex]:
                    del graph[p][vertex]
            for s in graph[vertex]:
                del tgraph[s][vertex]
            del graph[vertex]
            if vertex not in vo:
                del tgraph[vertex]
    jac = []
    for o in jaxpr.outvars:
        row = []
        for a in argnums:
            inv = jaxpr.invars[a]
            J = graph.get(key(inv), {}).get(key(o))
            shape = (B,) + tuple(o.shape) + tuple(inv.shape)
            row.append(np.zeros(shape, dtype=dtype) if J is None else J.reshape(shape))
        jac.append(row)
    return jac, [env[key(o)] for o in jaxpr.outvars]
