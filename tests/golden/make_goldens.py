"""Golden Jacobians in fp64 from an independent autodiff (torch.autograd), at the reference's test points.

graphax's tests assert ``jacve(...) == jax.jacrev(...)`` (tests/examples/example_test.py:18-27); JAX is
not available, so the same role is played by torch's reverse-mode autodiff applied to a plain primal
evaluation of the traced function (no graphax-style partial rules involved).  The printed values agree
with SURVEY.md Appendix D, which the survey produced from separately written torch restatements.

Run in the build container:  python tests/golden/make_goldens.py  ->  tests/golden/jacobians_fp64.json
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from graphax_b200.configs import get_workload  # noqa: E402
from graphax_b200.tracer import Literal  # noqa: E402


def evaluate(jaxpr, args):
    """Primal evaluation of a traced function with torch ops (float64)."""
    env = {}

    def read(a):
        return torch.tensor(a.val, dtype=torch.float64) if isinstance(a, Literal) else env[id(a)]
    for v, a in zip(jaxpr.invars, args):
        env[id(v)] = a
    un = {"neg": torch.neg, "abs": torch.abs, "sqrt": torch.sqrt, "sin": torch.sin, "cos": torch.cos,
          "tan": torch.tan, "exp": torch.exp, "log": torch.log, "tanh": torch.tanh, "log1p": torch.log1p,
          "sinh": torch.sinh, "cosh": torch.cosh, "asin": torch.asin, "acos": torch.acos, "atan": torch.atan,
          "asinh": torch.asinh, "acosh": torch.acosh, "atanh": torch.atanh, "erf": torch.erf,
          "logistic": torch.sigmoid, "square": torch.square}
    # mixed rank-0 / rank-1 operands broadcast like lax does
    for e in jaxpr.eqns:
        x = [read(a) for a in e.invars]
        p = e.primitive
        if p == "add": r = x[0] + x[1]
        elif p == "sub": r = x[0] - x[1]
        elif p == "mul": r = x[0] * x[1]
        elif p == "div": r = x[0] / x[1]
        elif p == "atan2": r = torch.atan2(x[0], x[1])
        elif p == "pow": r = torch.pow(x[0], x[1])
        elif p == "max": r = torch.maximum(x[0], x[1])
        elif p == "min": r = torch.minimum(x[0], x[1])
        elif p == "integer_pow": r = x[0] ** e.params["y"]
        elif p in un: r = un[p](x[0])
        elif p == "reduce_sum": r = x[0].sum()
        elif p == "dot_general": r = (x[0] * x[1]).sum()
        elif p == "slice": r = x[0][e.params["start_indices"][0]:e.params["limit_indices"][0]]
        elif p == "concatenate": r = torch.cat(x)
        elif p == "broadcast_in_dim": r = x[0].expand(e.params["shape"]).clone()
        else: raise NotImplementedError(p)
        env[id(e.outvar)] = r
    return tuple(env[id(o)] for o in jaxpr.outvars)


def main():
    out = {}
    for name in ("roe3d", "roe1d", "robotarm", "readme", "simple", "lighthouse", "hole", "helmholtz", "zoo"):
        w = get_workload(name)
        jaxpr = w.jaxpr()
        flat = torch.tensor(w.base, dtype=torch.float64)

        def f(v):
            args, off = [], 0
            for shape, size in zip(w.arg_shapes, w.sizes):
                args.append(v[off:off + size].reshape(shape))
                off += size
            return torch.cat([o.reshape(-1) for o in evaluate(jaxpr, args)])
        J = torch.autograd.functional.jacobian(f, flat)
        out[name] = {"inputs": w.base, "primal": f(flat).tolist(), "jacobian": J.tolist(),
                     "source": "torch.autograd fp64 on the traced primal; test points of the reference's "
                               "tests/examples/example_test.py (:78, :106, :123-128) and README.rst:42"}
        print(name, J.shape)
        torch.set_printoptions(precision=9, linewidth=200)
        print(J)
    (Path(__file__).with_name("jacobians_fp64.json")).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
