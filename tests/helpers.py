"""Shared test helpers (the only place besides smoke()/bench.py that touches oracle/)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import List, Sequence

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def pack_blocks(jac, batch: int, in_sizes: Sequence[int]) -> np.ndarray:
    """oracle / API block pytree -> packed [B, rows, cols] tile."""
    rows = []
    for row in jac:
        row = row if isinstance(row, (tuple, list)) else [row]
        blocks = []
        for b, n_in in zip(row, in_sizes):
            b = b.detach().cpu().numpy() if hasattr(b, "detach") else np.asarray(b)
            blocks.append(b.reshape(batch, -1, n_in))
        rows.append(np.concatenate(blocks, axis=2))
    return np.concatenate(rows, axis=1)


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def parity_bound(ref64: np.ndarray) -> np.ndarray:
    """|dJ| <= 1e-5 * max|J_sample| + 1e-5 * |J|  (BASELINE.md section 5)."""
    mx = np.abs(ref64).max(axis=(1, 2), keepdims=True)
    return 1e-5 * mx + 1e-5 * np.abs(ref64)


def load_random_g():
    from graphax_b200.tracer import jnp
    d = json.loads((GOLDEN / "random_g.json").read_text())

    def g(*xs):
        env = dict(zip(d["args"], xs))
        for dst, fn, args in d["ops"]:
            env[dst] = jnp.ones(()) if fn == "ones" else getattr(jnp, fn)(*[env[a] for a in args])
        return tuple(env[o] for o in d["outputs"])
    return g, d


def run_engine(plan, inputs: List[np.ndarray], kernel: str = "auto", device: int = 0, with_primal: bool = True,
               samples_per_thread: int = 0):
    """Launch through the C ABI on device buffers; returns (jac [B,r,c], primal [B,r], kernel name)."""
    import torch
    from graphax_b200 import runtime
    dp = runtime.DevicePlan(plan, device, jit=False)
    if kernel == "interpreter":
        dp.select_kernel(runtime.KERNEL_INTERPRETER)
    elif kernel == "jit":
        dp.specialize(samples_per_thread=samples_per_thread)
        dp.select_kernel(runtime.KERNEL_JIT)
    elif kernel == "aot":
        dp.select_kernel(runtime.KERNEL_AOT)
    batch = len(inputs[0])
    dev = [torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda(device) for x in inputs]
    jac = torch.full((batch, plan.n_rows, plan.n_cols), float("nan"), device=f"cuda:{device}")
    primal = torch.full((batch, plan.n_rows), float("nan"), device=f"cuda:{device}") if with_primal else None
    dp.launch(batch, [t.data_ptr() for t in dev], jac.data_ptr(),
              primal.data_ptr() if primal is not None else None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    name = dp.kernel
    out = jac.cpu().numpy(), (primal.cpu().numpy() if primal is not None else None), name
    dp.close()
    return out


def torch_eval_jaxpr(jaxpr, args):
    """Primal evaluation of a traced function with torch ops (float64): the independent autodiff the tests
    differentiate, in the role jax.jacrev / jax.jacfwd play in the reference's tests."""
    import torch
    from graphax_b200.tracer import Literal
    env = {}
    read = lambda a: torch.tensor(a.val, dtype=torch.float64) if isinstance(a, Literal) else env[id(a)]
    for v, a in zip(jaxpr.invars, args):
        env[id(v)] = a
    un = {"neg": torch.neg, "abs": torch.abs, "sqrt": torch.sqrt, "sin": torch.sin, "cos": torch.cos,
          "tan": torch.tan, "exp": torch.exp, "log": torch.log, "tanh": torch.tanh, "log1p": torch.log1p,
          "sinh": torch.sinh, "cosh": torch.cosh, "asin": torch.asin, "acos": torch.acos, "atan": torch.atan,
          "asinh": torch.asinh, "acosh": torch.acosh, "atanh": torch.atanh, "erf": torch.erf,
          "logistic": torch.sigmoid, "square": torch.square}
    for e in jaxpr.eqns:
        x = [read(a) for a in e.invars]
        p, prm = e.primitive, e.params
        if p in ("add", "sub", "mul", "div"):
            r = {"add": torch.add, "sub": torch.sub, "mul": torch.mul, "div": torch.div}[p](x[0], x[1])
        elif p == "atan2": r = torch.atan2(x[0], x[1])
        elif p == "pow": r = torch.pow(x[0], x[1])
        elif p == "max": r = torch.maximum(x[0], x[1])
        elif p == "min": r = torch.minimum(x[0], x[1])
        elif p == "integer_pow": r = x[0] ** prm["y"]
        elif p in ("gt", "lt", "eq", "ge", "le", "ne"):
            r = {"gt": torch.gt, "lt": torch.lt, "eq": torch.eq, "ge": torch.ge, "le": torch.le, "ne": torch.ne}[p](x[0], x[1]).to(torch.float64)
        elif p == "select_n": r = torch.where(x[0] != 0, x[2], x[1])
        elif p == "convert_element_type": r = x[0]
        elif p in un: r = un[p](x[0])
        elif p == "stop_gradient": r = x[0].detach()
        elif p == "reduce_sum": r = x[0].sum(dim=tuple(prm["axes"])) if x[0].ndim else x[0]
        elif p == "reduce_max": r = torch.amax(x[0], dim=tuple(prm["axes"]))
        elif p == "reduce_min": r = torch.amin(x[0], dim=tuple(prm["axes"]))
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
        elif p == "transpose": r = x[0].permute(*prm["permutation"])
        elif p in ("reshape", "squeeze"): r = x[0].reshape(e.outvar.shape)
        elif p == "slice": r = x[0][tuple(slice(a, b) for a, b in zip(prm["start_indices"], prm["limit_indices"]))]
        elif p == "concatenate": r = torch.cat([t.expand(v.shape) if hasattr(v, "shape") else t for t, v in zip(x, e.invars)], dim=prm["dimension"])
        elif p == "broadcast_in_dim":
            view = [1] * len(prm["shape"])
            for d, ax in zip(x[0].shape, prm["broadcast_dimensions"]):
                view[ax] = d
            r = x[0].reshape(view).expand(prm["shape"]).clone()
        else:
            raise NotImplementedError(p)
        env[id(e.outvar)] = r
    return tuple(env[id(o)] for o in jaxpr.outvars)


def jaxpr_from_record(rec, consts=()):
    """Rebuild an engine jaxpr from the equations recorded by tests/golden/make_reference_goldens.py (`dump_jaxpr`):
    the function exactly as the reference traced it, without needing its source.  ``consts``: the values of the arrays
    the function closes over (`const_shapes` in the record), numbered after the arguments."""
    from graphax_b200.tracer import Eqn, Jaxpr, Literal, Var
    env = {-(i + 1): Var(tuple(s), input_index=i) for i, s in enumerate(rec["in_shapes"])}
    n_in = len(rec["in_shapes"])
    const_shapes = rec.get("const_shapes", [])
    assert len(const_shapes) == len(consts), "pass the recorded constants"
    for i, shape in enumerate(const_shapes):
        env[-(n_in + i + 1)] = Var(tuple(shape), input_index=n_in + i, name=f"const{i}")
    eqns = []
    for n, e in enumerate(rec["eqns"], start=1):
        ins = tuple(Literal(float(v)) if kind == "lit" else env[v] for kind, v in e["operands"])
        out = Var(tuple(e["shape"]), eqn_index=n - 1)
        env[n] = out
        tup = lambda v: tuple(tup(x) for x in v) if isinstance(v, list) else v
        params = {k: tup(v) for k, v in e["params"].items() if k not in ("precision", "preferred_element_type", "sharding")}
        eqns.append(Eqn(e["primitive"], ins, out, params))
    return Jaxpr([env[-(i + 1)] for i in range(n_in)], eqns, [env[k] for k in rec["outvars"]],
                 [0] * len(eqns), True, [env[-(n_in + i + 1)] for i in range(len(const_shapes))],
                 [np.asarray(c, dtype=np.float32) for c in consts])
