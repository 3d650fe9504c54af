#!/usr/bin/env python
"""A generic-autodiff baseline on one B200: torch.func.vmap(torch.func.jacrev(f)) over the same synthetic batch.

SURVEY.md 8(d) asks for the reference's XLA-on-one-B200 run beside the engine's number; JAX is not in this image,
so the stand-in is PyTorch's functional autodiff of the same traced function (eager: one kernel per primitive and
per cotangent, every intermediate round-tripping HBM - the execution model the reference's emitted jnp ops have
without XLA fusion).  A reported baseline, never part of the product path."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def torch_function(jaxpr):
    """The traced function as a torch callable on per-sample tensors."""
    import torch
    from graphax_b200.tracer import Literal
    un = {"neg": torch.neg, "abs": torch.abs, "sqrt": torch.sqrt, "sin": torch.sin, "cos": torch.cos,
          "tan": torch.tan, "exp": torch.exp, "log": torch.log, "tanh": torch.tanh, "log1p": torch.log1p,
          "sinh": torch.sinh, "cosh": torch.cosh, "asin": torch.asin, "acos": torch.acos, "atan": torch.atan,
          "erf": torch.erf, "logistic": torch.sigmoid, "square": torch.square}

    def f(*args):
        env = {id(v): a for v, a in zip(jaxpr.invars, args)}
        read = lambda a: a.val if isinstance(a, Literal) else env[id(a)]
        for e in jaxpr.eqns:
            x = [read(a) for a in e.invars]
            p, prm = e.primitive, e.params
            if p == "add": r = x[0] + x[1]
            elif p == "sub": r = x[0] - x[1]
            elif p == "mul": r = x[0] * x[1]
            elif p == "div": r = x[0] / x[1]
            elif p == "atan2": r = torch.atan2(x[0], x[1])
            elif p == "pow": r = x[0] ** x[1]
            elif p == "integer_pow": r = x[0] ** prm["y"]
            elif p in un: r = un[p](x[0])
            elif p == "reduce_sum": r = x[0].sum(dim=tuple(prm["axes"])) if x[0].ndim else x[0]
            elif p == "dot_general": r = (x[0] * x[1]).sum() if x[0].ndim == 1 else x[0] @ x[1]
            elif p == "slice":
                r = x[0][tuple(slice(a, b) for a, b in zip(prm["start_indices"], prm["limit_indices"]))]
            elif p in ("squeeze", "reshape"): r = x[0].reshape(e.outvar.shape)
            elif p == "concatenate":
                parts = [t if torch.is_tensor(t) else torch.full(v.shape, float(t), dtype=args[0].dtype, device=args[0].device)
                         for t, v in zip(x, e.invars)]
                r = torch.cat(parts, dim=prm["dimension"])
            elif p == "broadcast_in_dim":
                src = x[0] if torch.is_tensor(x[0]) else torch.tensor(float(x[0]), dtype=args[0].dtype, device=args[0].device)
                view = [1] * len(prm["shape"])
                for d, ax in zip(src.shape, prm["broadcast_dimensions"]):
                    view[ax] = d
                r = src.reshape(view).expand(prm["shape"])
            else:
                raise NotImplementedError(p)
            env[id(e.outvar)] = r
        return tuple(env[id(o)] for o in jaxpr.outvars)
    return f


def measure(workload, batch: int, reps: int = 3, device: str = "cuda"):
    """Jacobians/s of vmap(jacrev(f)) (eager) on `device` over `batch` synthetic samples."""
    import torch
    f = torch_function(workload.jaxpr())
    xs = [torch.from_numpy(x).to(device) for x in workload.inputs(batch)]
    jac = torch.func.vmap(torch.func.jacrev(f, argnums=tuple(range(len(xs)))))
    out = jac(*xs)
    if device == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = jac(*xs)
    if device == "cuda":
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return batch / dt, out


if __name__ == "__main__":
    import numpy as np
    import torch
    from graphax_b200.configs import WORKLOADS
    w = WORKLOADS["roe3d"]
    v, out = measure(w, 1 << 16, device="cuda" if torch.cuda.is_available() else "cpu")
    print(f"torch vmap(jacrev) eager: {v / 1e6:.2f} M Jacobians/s")
