"""`jax.jacrev` / `jax.jacfwd` / `jax.grad` of the stand-in: an INDEPENDENT autodiff for the reference's own tests,
which compare `graphax.jacve` against `jax.jacrev`.  The function is traced with the stand-in's `make_jaxpr`, the
equations are evaluated with torch in float64 and differentiated by `torch.autograd`.  Test infrastructure only."""
from __future__ import annotations

import numpy as np

from . import _impl as J


def _torch_eval(jaxpr, consts, args):
    import torch
    env = {}
    for v, a in zip(jaxpr.invars, args):
        env[v] = a
    for v, c in zip(jaxpr.constvars, consts):
        env[v] = torch.tensor(np.asarray(J._raw(c)), dtype=torch.float64)

    def read(a):
        if isinstance(a, J.Literal):
            return torch.tensor(np.asarray(a.val, dtype=np.float64))
        return env[a]
    un = {"neg": torch.neg, "abs": torch.abs, "sqrt": torch.sqrt, "sin": torch.sin, "cos": torch.cos, "tan": torch.tan,
          "exp": torch.exp, "log": torch.log, "tanh": torch.tanh, "log1p": torch.log1p, "sinh": torch.sinh,
          "cosh": torch.cosh, "asin": torch.asin, "acos": torch.acos, "atan": torch.atan, "asinh": torch.asinh,
          "acosh": torch.acosh, "atanh": torch.atanh, "erf": torch.erf, "logistic": torch.sigmoid,
          "square": torch.square, "device_put": lambda x: x, "convert_element_type": lambda x: x}
    bi = {"add": torch.add, "add_any": torch.add, "sub": torch.sub, "mul": torch.mul, "div": torch.div, "atan2": torch.atan2,
          "pow": torch.pow, "max": torch.maximum, "min": torch.minimum}
    for e in jaxpr.eqns:
        x = [read(a) for a in e.invars]
        p, prm = e.primitive.name, e.params
        if p in bi:
            r = bi[p](x[0], x[1])
        elif p in ("gt", "lt", "eq"):
            r = {"gt": torch.gt, "lt": torch.lt, "eq": torch.eq}[p](x[0], x[1]).to(torch.float64)
        elif p == "select_n":
            r = torch.where(x[0] != 0, x[2], x[1])
        elif p == "integer_pow":
            r = x[0] ** prm["y"]
        elif p in un:
            r = un[p](x[0])
        elif p == "stop_gradient":
            r = x[0].detach()
        elif p == "reduce_sum":
            r = x[0].sum(dim=tuple(prm["axes"])) if x[0].ndim else x[0]
        elif p == "reduce_max":
            r = torch.amax(x[0], dim=tuple(prm["axes"]))
        elif p == "reduce_min":
            r = torch.amin(x[0], dim=tuple(prm["axes"]))
        elif p == "dot_general":
            (ca, cb), (ba, bb) = prm["dimension_numbers"]
            letters = iter("abcdefghijklmnopqrstuvw")
            la, lb = [None] * x[0].ndim, [None] * x[1].ndim
            for i, j in list(zip(ba, bb)) + list(zip(ca, cb)):
                la[i] = lb[j] = next(letters)
            la = [c or next(letters) for c in la]
            lb = [c or next(letters) for c in lb]
            out = [la[i] for i in ba] + [c for i, c in enumerate(la) if i not in ba and i not in ca] + \
                  [c for i, c in enumerate(lb) if i not in bb and i not in cb]
            r = torch.einsum(f"{''.join(la)},{''.join(lb)}->{''.join(out)}", x[0], x[1])
        elif p == "transpose":
            r = x[0].permute(*prm["permutation"])
        elif p in ("reshape", "squeeze"):
            r = x[0].reshape(e.outvars[0].aval.shape)
        elif p == "slice":
            r = x[0][tuple(slice(a, b) for a, b in zip(prm["start_indices"], prm["limit_indices"]))]
        elif p == "concatenate":
            r = torch.cat(x, dim=prm["dimension"])
        elif p == "broadcast_in_dim":
            view = [1] * len(prm["shape"])
            for d, ax in zip(x[0].shape, prm["broadcast_dimensions"]):
                view[ax] = d
            r = x[0].reshape(view).expand(tuple(prm["shape"])).clone()
        else:
            raise NotImplementedError(f"jaxlite autodiff: {p}")
        env[e.outvars[0]] = r
    return [read(o) for o in jaxpr.outvars]


def jacrev(fun, argnums=0, has_aux=False, **_):
    def jacfun(*args):
        import torch
        args = [a if J._is_value(a) else J.Array(a) for a in args]
        closed = J.make_jaxpr(fun)(*args)
        nums = (argnums,) if isinstance(argnums, int) else tuple(argnums)
        targs = [torch.tensor(np.asarray(J._raw(a), dtype=np.float64), requires_grad=(i in nums)) for i, a in enumerate(args)]
        outs = _torch_eval(closed.jaxpr, closed.consts, targs)
        leaves = []
        for o in outs:
            flat = o.reshape(-1)
            per_arg = []
            for i in nums:
                rows = []
                for r in range(flat.numel()):
                    g, = torch.autograd.grad(flat[r], targs[i], retain_graph=True, allow_unused=True) if flat.requires_grad \
                        else (None,)
                    rows.append(torch.zeros_like(targs[i]) if g is None else g)
                jac = torch.stack(rows).reshape(tuple(o.shape) + tuple(targs[i].shape)) if rows else \
                    torch.zeros(tuple(o.shape) + tuple(targs[i].shape))
                per_arg.append(J.Array(jac.numpy().astype(np.float32)))
            leaves.append(per_arg[0] if isinstance(argnums, int) else tuple(per_arg))
        # the output pytree of `fun`, every leaf replaced by its Jacobian(s)
        probe = fun(*args)
        _, treedef = J.tree_flatten(probe)
        return J.tree_unflatten(treedef, leaves)
    return jacfun


jacfwd = jacrev


def grad(fun, argnums=0, **_):
    return jacrev(fun, argnums)
