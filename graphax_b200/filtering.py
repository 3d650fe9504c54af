"""``filter_jacve`` - the reference's adapter for models that are pytrees (``equinox_bindings.py:44-135``).

``graphax.filter_jacve(loss_fn, order="rev")(model, x, y)`` differentiates ``loss_fn`` with respect to the
floating-point array leaves of its FIRST argument (an equinox ``Module``: a dataclass pytree of arrays and static
fields) and returns a pytree shaped like that argument, with ``None`` where a leaf is not differentiable
(reference ``tests/eqx_test.py``).  ``argnums`` may not be passed (``equinox_bindings.py:55-61``): to differentiate
several objects, collect them in a tuple and pass that as the first argument.

equinox is not a dependency of this engine (nor installed here), so the pytree protocol is a small one of its own:
dicts (keys in sorted order, as ``jax.tree_util`` does), lists, tuples, dataclass instances (fields in declaration
order - what an equinox Module is) and ``None``; everything else is a leaf.  Leaves that are floating-point arrays
(torch tensors or NumPy arrays) of the first argument are differentiated; array leaves of the other arguments are
passed through as (non-differentiated) arrays; any other leaf - ints, strings, callables such as an activation
function - is static and closed over, like ``equinox.partition(x, is_inexact_array)`` does.

The flattened function goes through ``jacve``, so large weight matrices take the tensor-valued path (reverse order,
tcgen05 contractions) and small ones the per-sample engine, and every option of ``jacve`` applies.
"""
from __future__ import annotations

import copy
import dataclasses
import functools as ft
from typing import Any, Callable, List, Tuple

import numpy as np

_SENTINEL = object()


# ---- a minimal pytree protocol ------------------------------------------------------------------------------------
def tree_flatten(tree: Any) -> Tuple[List[Any], Any]:
    """(leaves, treedef).  ``treedef`` is a nested description that ``tree_unflatten`` understands."""
    leaves: List[Any] = []

    def walk(node):
        if node is None:
            return ("none",)
        if isinstance(node, dict):
            keys = sorted(node)
            return ("dict", type(node), tuple(keys), tuple(walk(node[k]) for k in keys))
        if isinstance(node, (list, tuple)) and not hasattr(node, "_fields"):
            return ("seq", type(node), tuple(walk(v) for v in node))
        if dataclasses.is_dataclass(node) and not isinstance(node, type):
            names = tuple(f.name for f in dataclasses.fields(node))
            return ("dataclass", node, names, tuple(walk(getattr(node, n)) for n in names))
        leaves.append(node)
        return ("leaf",)
    return leaves, walk(tree)


def tree_unflatten(treedef: Any, leaves: List[Any]) -> Any:
    it = iter(leaves)

    def build(d):
        tag = d[0]
        if tag == "none":
            return None
        if tag == "leaf":
            return next(it)
        if tag == "dict":
            return d[1]((k, build(c)) for k, c in zip(d[2], d[3])) if d[1] is dict else {k: build(c) for k, c in zip(d[2], d[3])}
        if tag == "seq":
            vals = [build(c) for c in d[2]]
            return vals if d[1] is list else tuple(vals)
        proto, names, children = d[1], d[2], d[3]
        new = copy.copy(proto)                        # no __init__ / __post_init__: fields are overwritten one by one
        for n, c in zip(names, children):
            object.__setattr__(new, n, build(c))      # works for frozen dataclasses (equinox Modules are frozen)
        return new
    return build(treedef)


def is_array(x: Any) -> bool:
    import torch
    return isinstance(x, (np.ndarray, torch.Tensor))


def is_inexact_array(x: Any) -> bool:
    """``equinox.is_inexact_array``: a floating-point array."""
    import torch
    if isinstance(x, torch.Tensor):
        return x.is_floating_point()
    return isinstance(x, np.ndarray) and np.issubdtype(x.dtype, np.floating)


def _static_key(x: Any):
    try:
        hash(x)
        return x
    except TypeError:
        return id(x)


# ---- filter_jacve -----------------------------------------------------------------------------------------------
def filter_jacve(fun: Callable = _SENTINEL, **gradkwargs) -> Callable:
    """``filter_jacve(fun, order=..., count_ops=False, sparse_representation=False, ...)(model, *args, **kwargs)``.

    Returns the Jacobian of ``fun`` with respect to every floating-point array leaf of ``model`` as a pytree shaped
    like ``model`` (``None`` at the other leaves); with several outputs, a tuple of such pytrees.  ``count_ops=True``
    returns ``(pytree, counts)`` like ``jacve``."""
    if fun is _SENTINEL:
        return ft.partial(filter_jacve, **gradkwargs)
    if gradkwargs.pop("argnums", None) is not None:
        raise ValueError("`argnums` should not be passed. If you need to differentiate multiple objects then collect "
                         "them into a tuple and pass that as the first argument.")
    if "order" not in gradkwargs:
        raise TypeError("filter_jacve needs an elimination order (order=...)")
    cache = {}

    @ft.wraps(fun)
    def wrapped(x, *args, **kwargs):
        from .api import jacve
        x_leaves, x_def = tree_flatten(x)
        a_leaves, a_def = tree_flatten(args)
        diff = [i for i, l in enumerate(x_leaves) if is_inexact_array(l)]
        passed = [i for i, l in enumerate(a_leaves) if is_array(l)]
        if not diff:
            raise ValueError("the first argument has no floating-point array leaf to differentiate")

        def flat(*arrays):
            xl, al = list(x_leaves), list(a_leaves)
            for i, a in zip(diff, arrays[:len(diff)]):
                xl[i] = a
            for i, a in zip(passed, arrays[len(diff):]):
                al[i] = a
            return fun(tree_unflatten(x_def, xl), *tree_unflatten(a_def, al), **kwargs)

        static = tuple(_static_key(l) for i, l in enumerate(x_leaves) if i not in diff) + \
            tuple(_static_key(l) for i, l in enumerate(a_leaves) if i not in passed)
        key = (repr(_shape_of_def(x_def)), repr(_shape_of_def(a_def)), static, tuple(sorted(kwargs)))
        entry = cache.get(key)
        if entry is None:
            from .tracer import make_jaxpr
            jf = jacve(flat, argnums=tuple(range(len(diff))), **gradkwargs)
            shapes = [tuple(x_leaves[i].shape) for i in diff] + [tuple(a_leaves[i].shape) for i in passed]
            n_out = len(make_jaxpr(flat, shapes, gradkwargs.get("frontend", "auto")).outvars)
            entry = cache[key] = (jf, n_out)
        jf, n_out = entry
        out = jf(*[x_leaves[i] for i in diff], *[a_leaves[i] for i in passed])
        aux = None
        if gradkwargs.get("count_ops") or gradkwargs.get("has_aux"):
            first, second = out
            out, aux = (first, second) if gradkwargs.get("count_ops") else (second, first)

        def like_model(blocks):
            blocks = blocks if isinstance(blocks, tuple) else (blocks,)
            leaves = [None] * len(x_leaves)
            for i, b in zip(diff, blocks):
                leaves[i] = b
            return tree_unflatten(x_def, leaves)
        # jacve: a tuple per output of a tuple per argnum; a single output of a function of several arguments is
        # unwrapped (core.py:76-92), a single argnum gives the bare block
        if n_out == 1:
            row = out if len(diff) + len(passed) > 1 or not isinstance(out, tuple) else out[0]
            result = like_model(row)
        else:
            result = tuple(like_model(o) for o in out)
        if aux is None:
            return result
        return (result, aux) if gradkwargs.get("count_ops") else (aux, result)
    return wrapped


def _shape_of_def(d):
    """The treedef without the dataclass prototype objects (for cache keys)."""
    if d[0] == "dataclass":
        return ("dataclass", type(d[1]).__name__, d[2], tuple(_shape_of_def(c) for c in d[3]))
    if d[0] == "dict":
        return ("dict", d[2], tuple(_shape_of_def(c) for c in d[3]))
    if d[0] == "seq":
        return ("seq", d[1].__name__, tuple(_shape_of_def(c) for c in d[2]))
    return d
