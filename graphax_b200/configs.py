"""Named workloads of BASELINE.json: function, per-sample shapes, synthetic inputs, orders.

Synthetic inputs follow SURVEY.md 8(d): the reference's own test point times
``1 + jitter * U(-1, 1)`` with a counter-based generator keyed by (seed, global
sample index, lane), so a shard's data does not depend on how many GPUs share
the batch.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import lru_cache
from typing import Callable, Dict, List, Sequence, Tuple, Union

import numpy as np

from .examples import ElementalZoo, Helmholtz, Hole, Lighthouse, RobotArm_6DOF, RoeFlux_1d, RoeFlux_3d, Simple, readme_f
from .examples import orders as ref_orders
from .plan import Plan, build_plan
from .tracer import Jaxpr, make_jaxpr

SEED = 1234


def uniform_pm1(seed: int, start: int, count: int, lanes: int) -> np.ndarray:
    """U(-1, 1) as float64 [count, lanes]; value depends only on (seed, sample index, lane)."""
    with np.errstate(over="ignore"):
        idx = (np.arange(start, start + count, dtype=np.uint64)[:, None] * np.uint64(lanes)
               + np.arange(lanes, dtype=np.uint64)[None, :])
        z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed) * np.uint64(0xD1B54A32D192ED03)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -52) - 1.0


@dataclass
class Workload:
    name: str
    fun: Callable
    arg_shapes: List[Tuple[int, ...]]
    base: List[float]                    # flattened base point (reference test inputs)
    jitter: float
    batch: int                           # batch of the BASELINE.json config
    orders: Dict[str, Union[str, Sequence[int]]] = field(default_factory=dict)
    vmap_orders: Dict[str, Sequence[int]] = field(default_factory=dict)   # written in the 146-numbering

    @property
    def argnums(self) -> List[int]:
        return list(range(len(self.arg_shapes)))

    @property
    def sizes(self) -> List[int]:
        return [int(np.prod(s, dtype=np.int64)) if s else 1 for s in self.arg_shapes]

    def jaxpr(self) -> Jaxpr:
        return _jaxpr(self.name)

    def order(self, key: str):
        if key in self.orders:
            return self.orders[key]
        return self.jaxpr().translate_vmap_order(self.vmap_orders[key])

    def order_names(self) -> List[str]:
        return list(self.orders) + list(self.vmap_orders)

    def plan(self, order_key: str, with_primal: bool = True, division: str = None, rounding: str = None) -> Plan:
        if rounding is None:
            from .plan import DEFAULT_ROUNDING
            rounding = "fma" if division is not None else DEFAULT_ROUNDING
        if division is None:
            division = "ieee" if rounding == "reference" else "reciprocal"
        return _plan(self.name, order_key, with_primal, division, rounding)

    def inputs(self, batch: int, start: int = 0, seed: int = SEED) -> List[np.ndarray]:
        """float32 arrays [batch, *shape] (rank-0 arguments -> [batch])."""
        n = sum(self.sizes)
        u = uniform_pm1(seed, start, batch, n)
        flat = (np.asarray(self.base, dtype=np.float64)[None, :] * (1.0 + self.jitter * u)).astype(np.float32)
        out, off = [], 0
        for shape, size in zip(self.arg_shapes, self.sizes):
            out.append(np.ascontiguousarray(flat[:, off:off + size].reshape((batch,) + tuple(shape))))
            off += size
        return out


WORKLOADS: Dict[str, Workload] = {
    # config 1: README call  jacve(f, order=[1,3,2,4], argnums=(0,1,2,3))(1.,1.,2.,7.)
    "readme": Workload("readme", readme_f, [()] * 4, [1., 1., 2., 7.], 0.01, 1 << 10,
                       {"readme": [1, 3, 2, 4], "fwd": "fwd", "rev": "rev"}),
    # config 2: RoeFlux_1d, order 'rev', batch 2^16 (test point example_test.py:106)
    "roe1d": Workload("roe1d", RoeFlux_1d, [()] * 6, [.01, .02, .02, .01, .03, .03], 0.01, 1 << 16,
                      {"rev": "rev", "fwd": "fwd", "alphagrad": ref_orders.ROEFLUX_1D_ALPHAGRAD,
                       "markowitz": ref_orders.ROEFLUX_1D_MARKOWITZ}),
    # config 3 (headline): RoeFlux_3d, batch 2^20 (test point example_test.py:123-128)
    "roe3d": Workload("roe3d", RoeFlux_3d, [(1,), (3,), (1,), (1,), (3,), (1,)],
                      [.1, .1, .2, .3, .5, .2, .2, .2, .4, .6], 0.01, 1 << 20,
                      {"rev": "rev", "fwd": "fwd"},
                      {"mixed": ref_orders.ROEFLUX_3D_VMAP_MIXED, "mixed2": ref_orders.ROEFLUX_3D_VMAP_MIXED2}),
    # config 4: RobotArm_6DOF, batch 2^22 (test point example_test.py:78)
    "robotarm": Workload("robotarm", RobotArm_6DOF, [()] * 6, [.02] * 6, 0.5, 1 << 22,
                         {"rev": "rev", "fwd": "fwd", "alphagrad": ref_orders.ROBOTARM_ALPHAGRAD,
                          "markowitz": ref_orders.ROBOTARM_MARKOWITZ}),
}


# Further functions of the reference's tests (tests/examples/example_test.py:43-75); no ahead-of-time
# kernels: they exercise the run-time (NVRTC) specialisation path.
EXTRA_WORKLOADS: Dict[str, Workload] = {
    "simple": Workload("simple", Simple, [()] * 2, [5., 7.], 0.01, 1 << 10, {"fwd": "fwd", "rev": "rev"}),
    "lighthouse": Workload("lighthouse", Lighthouse, [()] * 4, [.02] * 4, 0.01, 1 << 10,
                           {"fwd": "fwd", "rev": "rev"}),
    "hole": Workload("hole", Hole, [()] * 4, [.3, .5, .7, .2], 0.01, 1 << 10, {"fwd": "fwd", "rev": "rev"}),
    "helmholtz": Workload("helmholtz", Helmholtz, [(4,)], [0.05, 0.15, 0.25, 0.35], 0.01, 1 << 10,
                          {"fwd": "fwd", "rev": "rev", "custom": [2, 5, 4, 3, 1]}),
    "zoo": Workload("zoo", ElementalZoo, [(), (), (3,)], [0.4, 1.7, 0.3, 1.1, 2.2], 0.01, 1 << 10,
                    {"fwd": "fwd", "rev": "rev"}),
}


def get_workload(name: str) -> Workload:
    return WORKLOADS[name] if name in WORKLOADS else EXTRA_WORKLOADS[name]


@lru_cache(maxsize=None)
def _jaxpr(name: str) -> Jaxpr:
    w = get_workload(name)
    return make_jaxpr(w.fun, w.arg_shapes)


@lru_cache(maxsize=None)
def _plan(name: str, order_key: str, with_primal: bool, division: str, rounding: str = "fma") -> Plan:
    w = get_workload(name)
    return build_plan(w.jaxpr(), w.order(order_key), w.argnums, with_primal=with_primal, division=division,
                      rounding=rounding)


def named_plans() -> List[Tuple[str, Plan]]:
    """Every (workload, order) pair that gets an ahead-of-time kernel."""
    out = []
    for w in WORKLOADS.values():
        for key in w.order_names():
            for division in ("reciprocal", "ieee"):
                out.append((f"{w.name}:{key}:{division}", w.plan(key, division=division)))
            out.append((f"{w.name}:{key}:reference", w.plan(key, rounding="reference")))
    return out
