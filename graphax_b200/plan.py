"""Elimination plan: the flat, static program one Jacobian evaluation runs.

``build_plan`` traces nothing and computes nothing: it takes the symbolic result
of ``elimination.accumulate`` and linearises the expression DAG into

* an SSA op list (used by the CUDA source emitter, ``codegen.py``), and
* a binary blob with physical slots (consumed by ``gx_plan_create`` for the
  generic persistent interpreter kernel and by the oracle's C replay).

Blob layout (little endian, version 1)::

    u32[16] header: magic 'GXPL', version, total_bytes, n_in_arrays, n_out_arrays,
                    n_in_scalars, n_cols, n_rows, n_ops, n_slots, n_consts, flags,
                    reserved[4]
    u32[n_in_arrays]  in_sizes      scalars per input array, in argument order
    u32[n_out_arrays] out_sizes     scalars per output array
    u32[n_in_arrays]  in_cols       first Jacobian column of the array, 0xffffffff if not in argnums
    f32[n_consts]     consts
    op[n_ops]         {u16 opcode, dst, a, b, c, 0}
        operand = kind<<14 | index, kind 0 slot / 1 const / 2 input scalar
        store_jac:    dst = row*n_cols + col, a = value
        store_primal: dst = row,              a = value

The Jacobian tile is ``[n_rows, n_cols]`` row-major per sample: rows are the
output scalars in output order, columns the scalars of the ``argnums`` inputs in
``argnums`` order.  Every tile position is written by exactly one ``store_jac``
(structural zeros store the constant 0).
"""
from __future__ import annotations

import struct as _struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .elimination import Accumulated, Graph, accumulate
from .lowering import OP, Dag
from .tracer import Jaxpr, Var

# Where plans round unless the caller says otherwise (see build_plan): "reference" = the reference's own sequence of
# rounded operations (bit-identical results), "fma" = contracted (fewer instructions, a few ulp per operation away).
# An explicit ``division=`` or ``contract=True`` selects "fma": both are options of that mode.
DEFAULT_ROUNDING = "reference"
# How the plan's operations are ordered (build_plan): "minpressure" = fewest live values over alternating bottom-up /
# top-down greedy passes.  GX_SCHEDULE overrides the default (experiments; it must then be the same at build time of
# the ahead-of-time kernels and at run time, since the order is part of the plan and hence of its key).
DEFAULT_SCHEDULE = "minpressure"
MAGIC = 0x4C505847
VERSION = 1
FLAG_HAS_PRIMAL = 1
K_SLOT, K_CONST, K_INPUT = 0, 1, 2
MAX_INDEX = (1 << 14) - 1


def fnv1a64(data: bytes) -> int:
    h = 0xCBF29CE484222325
    for b in data:
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


@dataclass
class SsaOp:
    op: str
    dst: int                  # SSA value id (== DAG node id) or tile/row index for stores
    args: Tuple[int, ...]     # DAG node ids


@dataclass
class Plan:
    in_sizes: List[int]
    out_sizes: List[int]
    argnums: List[int]
    n_rows: int
    n_cols: int
    order: List[int]
    dag: Dag
    ssa: List[SsaOp]
    tile: List[Optional[int]]           # per tile position: DAG node or None (structural zero)
    primal_nodes: List[int]
    blob: bytes = b""
    n_slots: int = 0
    stats: Dict[str, object] = field(default_factory=dict)
    structure: Dict[str, object] = field(default_factory=dict)
    division: str = "reciprocal"
    rounding: str = "fma"

    @property
    def key(self) -> str:
        """FNV-1a 64 of the blob (same function as gx_hash_bytes): names AOT/JIT kernels.  Cached per blob
        object: the pure-Python hash of a 16 KB blob costs ~1 ms, which the host call path must not pay."""
        cached = self.__dict__.get("_key_cache")
        if cached is None or cached[0] is not self.blob:
            cached = (self.blob, f"{fnv1a64(self.blob):016x}")
            self.__dict__["_key_cache"] = cached
        return cached[1]

    @property
    def n_in_scalars(self) -> int:
        return sum(self.in_sizes)

    @property
    def in_cols(self) -> List[int]:
        cols, c = [], 0
        first = {}
        for a in self.argnums:
            first[a] = c
            c += self.in_sizes[a]
        return [first.get(i, 0xFFFFFFFF) for i in range(len(self.in_sizes))]

    @property
    def bytes_per_jacobian(self) -> int:
        """Compulsory HBM traffic: read every input once, write the dense tile once
        (240 B for RoeFlux_3d)."""
        return 4 * (self.n_in_scalars + self.n_rows * self.n_cols)


def _linearise(dag: Dag, first_elim: int, roots: Sequence[int]) -> List[int]:
    """Topological order of the nodes reachable from ``roots``: elimination nodes in
    the order the elimination created them, each preceded by whatever it still
    needs (partials and primal values are emitted right before their first use)."""
    reachable = set()
    stack = [r for r in roots]
    while stack:
        n = stack.pop()
        if n in reachable:
            continue
        reachable.add(n)
        stack.extend(dag.args[n] if dag.op[n] not in ("const", "input") else ())
    emitted, out = set(), []

    def visit(root: int) -> None:
        work = [(root, 0)]
        while work:
            n, i = work.pop()
            if n in emitted:
                continue
            args = dag.args[n] if dag.op[n] not in ("const", "input") else ()
            if i < len(args):
                work.append((n, i + 1))
                if args[i] not in emitted:
                    work.append((args[i], 0))
            else:
                emitted.add(n)
                if dag.op[n] not in ("const", "input"):
                    out.append(n)

    for n in range(first_elim, len(dag.op)):
        if n in reachable:
            visit(n)
    for r in roots:
        visit(r)
    return out


def _contract_sums(dag: Dag, roots: Sequence[int], max_terms: int = 1 << 30) -> Dict[int, int]:
    """Fold additions into the multiply chains that feed them: ``c + (a0*b0 + a1*b1 + ..)`` becomes
    ``fma(ak, bk, .. fma(a0, b0, c))`` and ``c - (..)`` the same with negated ``a``s, whenever the chain's
    intermediate values have no other reader.  One instruction less per accumulation into an existing edge
    (core.py:262-272 adds the product to the old edge value); the sum is still one rounded operation per
    node, only its association changes (old edge first instead of last), which stays inside the stated
    fp32 tolerance and is what the plan, its replay and every kernel then agree on bit for bit.
    Returns old node -> new node for every reachable node."""
    reach, stack = set(), list(roots)
    while stack:
        n = stack.pop()
        if n not in reach:
            reach.add(n)
            if dag.op[n] not in ("const", "input"):
                stack.extend(dag.args[n])
    uses: Dict[int, int] = {}
    for n in reach:
        if dag.op[n] not in ("const", "input"):
            for a in dag.args[n]:
                uses[a] = uses.get(a, 0) + 1
    for r in roots:
        uses[r] = uses.get(r, 0) + 1

    def chain(t: int) -> Optional[List[Tuple[int, int]]]:
        """[(a0,b0), (a1,b1), ..] if t is a privately owned mul/fma chain."""
        terms = []
        while True:
            if uses.get(t, 0) != 1:
                return None
            if dag.op[t] == "mul":
                terms.append(dag.args[t])
                return terms[::-1]
            if dag.op[t] != "fma":
                return None
            a, b, c = dag.args[t]
            terms.append((a, b))
            t = c

    new: Dict[int, int] = {}
    for n in sorted(reach):                       # node ids are topological (operands are created first)
        op = dag.op[n]
        if op in ("const", "input"):
            new[n] = n
            continue
        args = dag.args[n]
        done = False
        if op in ("add", "sub"):
            x, y = args
            for t, other, negate_terms, negate_other in ((y, x, op == "sub", False), (x, y, False, op == "sub")):
                terms = chain(t)
                if terms is None or len(terms) > max_terms:
                    continue
                acc = new[other]
                if negate_other:
                    acc = dag.neg(acc)
                for a, b in terms:
                    a, b = new[a], new[b]
                    if negate_terms:
                        if dag.op[b] in ("neg", "const") and dag.op[a] not in ("neg", "const"):
                            a, b = b, a
                        a = dag.neg(a)
                    acc = dag.fma(a, b, acc)
                new[n] = acc
                done = True
                break
        if not done:
            new[n] = dag.emit(op, *[new[a] for a in args])
    return new


def _reschedule_for_pressure(dag: Dag, sched: List[int], roots: Sequence[int]) -> List[int]:
    """Greedy list scheduling of the straight-line program for few live values.

    Any topological order of the DAG computes the same bits (every node is one rounded operation
    on fixed operands), so the order is free.  Among the ready operations pick the one that ends
    the most live ranges (operands whose last consumer it is) net of the value it creates; ties go
    to the original position.  On the RoeFlux_3d plans this cuts the peak number of live values by
    13-28 %, which is what decides registers per thread, spills and resident warps per SM."""
    import heapq
    nodeset = set(sched)
    pos = {n: i for i, n in enumerate(sched)}
    args = {n: [a for a in dict.fromkeys(dag.args[n]) if a in nodeset] for n in sched}
    users: Dict[int, List[int]] = {n: [] for n in sched}
    for n in sched:
        for a in args[n]:
            users[a].append(n)
    remaining = {n: len(users[n]) for n in sched}
    indeg = {n: len(args[n]) for n in sched}

    def key(n: int):
        freed = sum(1 for a in args[n] if remaining[a] == 1)
        creates = 1 if users[n] else 0          # values only read by a store die immediately
        return (-(freed - creates), pos[n])      # heap minimum = best (freed - creates), then original position

    # a ready operation's score only changes when one of its operands drops to a single remaining reader: the heap
    # holds (key, node) entries, stale ones (key no longer current, or node already emitted) are skipped when popped
    current: Dict[int, tuple] = {}
    heap: List[tuple] = []
    for n in sched:
        if indeg[n] == 0:
            current[n] = key(n)
            heap.append((current[n], n))
    heapq.heapify(heap)
    done = set()
    out: List[int] = []
    while heap:
        k, best = heapq.heappop(heap)
        if best in done or current.get(best) != k:
            continue
        done.add(best)
        out.append(best)
        for a in args[best]:
            remaining[a] -= 1
            if remaining[a] == 1:
                for u in users[a]:
                    if u not in done and u in current:       # its last reader is already ready: it now frees `a`
                        current[u] = key(u)
                        heapq.heappush(heap, (current[u], u))
        for u in users[best]:
            indeg[u] -= 1
            if indeg[u] == 0:
                current[u] = key(u)
                heapq.heappush(heap, (current[u], u))
    if len(out) != len(sched):
        raise AssertionError("scheduler lost operations")
    return out


def _reschedule_bottom_up(dag: Dag, sched: List[int]) -> List[int]:
    """The mirror image of ``_reschedule_for_pressure``: build the order from its END.  An operation may be placed
    once all its readers are; placing it closes its own live range and opens one for every operand that no later
    operation reads.  Pick the one that opens the fewest net of the one it closes, ties to the latest original
    position.  Where the top-down pass computes a value as soon as that frees its operands (and may then hold the
    result for a thousand operations), this pass computes it right before its first reader."""
    import heapq
    nodeset = set(sched)
    pos = {n: i for i, n in enumerate(sched)}
    args = {n: [a for a in dict.fromkeys(dag.args[n]) if a in nodeset] for n in sched}
    users: Dict[int, List[int]] = {n: [] for n in sched}
    for n in sched:
        for a in args[n]:
            users[a].append(n)
    unplaced = {n: len(users[n]) for n in sched}
    live = set()                                   # read by an operation already placed, not yet defined

    def key(n: int):
        opens = sum(1 for a in args[n] if a not in live)
        return (opens - (1 if n in live else 0), -pos[n])

    current: Dict[int, tuple] = {}
    heap: List[tuple] = []
    for n in sched:
        if unplaced[n] == 0:
            current[n] = key(n)
            heap.append((current[n], n))
    heapq.heapify(heap)
    done = set()
    out: List[int] = []
    while heap:
        k, best = heapq.heappop(heap)
        if best in done or current.get(best) != k:
            continue
        done.add(best)
        out.append(best)
        live.discard(best)
        for a in args[best]:
            if a not in live:
                live.add(a)
                for u in users[a]:                 # its other readers no longer open this range
                    if u not in done and u in current:
                        current[u] = key(u)
                        heapq.heappush(heap, (current[u], u))
            unplaced[a] -= 1
            if unplaced[a] == 0:
                current[a] = key(a)
                heapq.heappush(heap, (current[a], a))
    if len(out) != len(sched):
        raise AssertionError("scheduler lost operations")
    return out[::-1]


def peak_live_values(dag: Dag, sched: Sequence[int]) -> int:
    """Largest number of simultaneously live values of a schedule (inputs count from the start to their last
    reader, constants never): what the register allocator of the specialised kernel has to fit."""
    last: Dict[int, int] = {}
    for i, n in enumerate(sched):
        for a in dag.args[n]:
            if dag.op[a] != "const":
                last[a] = i
    ends: Dict[int, int] = {}
    for a, i in last.items():
        ends[i] = ends.get(i, 0) + 1
    live = sum(1 for a in last if dag.op[a] == "input")
    peak = live
    for i, n in enumerate(sched):
        live += 1 if n in last else 0
        peak = max(peak, live)
        live -= ends.get(i, 0)
    return peak


def _reschedule_min_pressure(dag: Dag, sched: List[int], roots: Sequence[int], rounds: int = 4) -> List[int]:
    """Alternate the bottom-up and the top-down greedy pass, each starting from the other's result (their tie
    breaks follow the incoming order), and keep the schedule with the fewest live values.  RoeFlux_3d, reference
    rounding: 194 -> 159 live values for the mixed order after two passes.  Fewer live values are not only fewer
    spills: the register allocator avoids same-bank operand pairs (a two-cycle dispatch for FMUL / FADD, DESIGN
    3.3) only while it has registers to choose from - 386 such pairs at 194 live values, 166 at 159, in the same
    168 registers."""
    best = _reschedule_for_pressure(dag, sched, roots)
    best_peak = peak_live_values(dag, best)
    cur = sched
    for _ in range(rounds):
        improved = False
        for step in (_reschedule_bottom_up, lambda d, s: _reschedule_for_pressure(d, s, roots)):
            cur = step(dag, cur)
            peak = peak_live_values(dag, cur)
            if peak < best_peak:
                best, best_peak, improved = cur, peak, True
        if not improved:
            break
    return best


def _random_topological_order(dag: Dag, sched: List[int], seed: int, sigma: float = 8.0) -> List[int]:
    """``sched`` with every operation's position jittered by a Gaussian of ``sigma`` positions, made topological
    again (priority = jittered position)."""
    import heapq
    import random
    rng = random.Random(seed)
    nodeset = set(sched)
    args = {n: [a for a in dict.fromkeys(dag.args[n]) if a in nodeset] for n in sched}
    users: Dict[int, List[int]] = {n: [] for n in sched}
    for n in sched:
        for a in args[n]:
            users[a].append(n)
    prio = {n: i + rng.gauss(0.0, sigma) for i, n in enumerate(sched)}
    indeg = {n: len(args[n]) for n in sched}
    heap = [(prio[n], n) for n in sched if indeg[n] == 0]
    heapq.heapify(heap)
    out: List[int] = []
    while heap:
        _, n = heapq.heappop(heap)
        out.append(n)
        for u in users[n]:
            indeg[u] -= 1
            if indeg[u] == 0:
                heapq.heappush(heap, (prio[u], u))
    if len(out) != len(sched):
        raise AssertionError("scheduler lost operations")
    return out


def _reschedule_for_ilp(dag: Dag, sched: List[int], limit: int, window: int) -> List[int]:
    """List scheduling for instruction-level parallelism under a cap on live values.

    A warp issues a dependent FFMA only every ~4 cycles, and with 168 registers per thread there are
    just 3 warps per scheduler to hide that, so the straight-line body must interleave independent
    chains itself.  Among the ready operations prefer one whose operands were all produced at least
    ``window`` operations ago, then the one with the longest path to the end of the program; once
    ``limit`` values are live, fall back to the pressure rule (end as many live ranges as possible)."""
    nodeset = set(sched)
    pos = {n: i for i, n in enumerate(sched)}
    args = {n: [a for a in dict.fromkeys(dag.args[n]) if a in nodeset] for n in sched}
    users: Dict[int, List[int]] = {n: [] for n in sched}
    for n in sched:
        for a in args[n]:
            users[a].append(n)
    height: Dict[int, int] = {}
    for n in reversed(sched):
        lat = 8 if dag.op[n] in ("div", "sqrt") or dag.op[n] not in ("add", "sub", "mul", "fma", "neg") else 1
        height[n] = lat + max((height[u] for u in users[n]), default=0)
    remaining = {n: len(users[n]) for n in sched}
    indeg = {n: len(args[n]) for n in sched}
    ready = [n for n in sched if indeg[n] == 0]
    issued: Dict[int, int] = {}
    out: List[int] = []
    live = 0
    while ready:
        now = len(out)
        best, best_score = None, None
        for n in ready:
            freed = sum(1 for a in args[n] if remaining[a] == 1)
            creates = 1 if users[n] else 0
            rested = all(now - issued[a] >= window for a in args[n])
            if live < limit:
                score = (rested, height[n], freed - creates, -pos[n])
            else:
                score = (freed - creates, rested, height[n], -pos[n])
            if best_score is None or score > best_score:
                best, best_score = n, score
        ready.remove(best)
        issued[best] = now
        out.append(best)
        live += 1 if users[best] else 0
        for a in args[best]:
            remaining[a] -= 1
            if remaining[a] == 0:
                live -= 1
        for u in users[best]:
            indeg[u] -= 1
            if indeg[u] == 0:
                ready.append(u)
    if len(out) != len(sched):
        raise AssertionError("scheduler lost operations")
    return out


def build_plan(jaxpr: Jaxpr, order, argnums: Sequence[int], with_primal: bool = True,
               division: Optional[str] = None, schedule: Optional[str] = None, contract: bool = False,
               rounding: Optional[str] = None) -> Plan:
    """Lower (jaxpr, order, argnums) to a plan.

    ``rounding`` fixes where the plan rounds:

    * ``"reference"`` - the reference's own sequence of rounded operations: every ``J[s,p] += J[s,v] J[v,p]`` is a
      rounded product followed by a rounded sum (``post_val * pre_val`` then ``+=``, core.py:210, 262-272; no FMA
      contraction), contractions sum their products in ascending index order, and ``x/y`` with its partials is the
      three IEEE divisions of primitives.py:197-199.  Results are bit-identical to the unfused float32 evaluation
      of the reference algorithm (``oracle/ve_numpy.py``) up to the sign of exact zeros.
    * ``"fma"`` - product-and-add pairs are contracted into FMAs (one rounding instead of two) and, by default,
      ``x/y`` is lowered through one reciprocal (``division="reciprocal"``).  Fewer operations; agrees with the
      reference to a few ulp per operation, which an ill-conditioned elimination order can amplify.

    ``division`` ("reciprocal" | "ieee", ``lowering.lower_eqn``) defaults to "reciprocal" for ``rounding="fma"`` and
    must be "ieee" for ``rounding="reference"``.  ``contract=True`` additionally folds additions into the multiply
    chains feeding them (``_contract_sums``): 6-12 % fewer instructions, but a different association of the sums
    than the reference's, so it is opt-in (``rounding="fma"`` only)."""
    if schedule is None:
        import os
        schedule = os.environ.get("GX_SCHEDULE") or DEFAULT_SCHEDULE
    if rounding is None:
        rounding = "fma" if (contract or division is not None) else DEFAULT_ROUNDING
    if rounding not in ("fma", "reference"):
        raise ValueError("rounding must be 'fma' or 'reference'")
    if division is None:
        division = "ieee" if rounding == "reference" else "reciprocal"
    if rounding == "reference" and (division != "ieee" or contract):
        raise ValueError("rounding='reference' implies division='ieee' and contract=False")
    argnums = [int(a) for a in argnums]
    for a in argnums:
        if not 0 <= a < len(jaxpr.invars):
            raise ValueError(f"argnum {a} out of range for a function of {len(jaxpr.invars)} arguments")
    acc = accumulate(jaxpr, order, argnums, division, fuse=(rounding == "fma"))
    g, dag = acc.graph, acc.graph.dag

    in_sizes = [v.size for v in jaxpr.invars]
    out_sizes = [v.size for v in jaxpr.outvars]
    if sum(in_sizes) > 4096 or sum(out_sizes) * sum(in_sizes[a] for a in argnums) > 16384:
        raise NotImplementedError("per-sample plans are scalarised: use the tensor-valued path (dense.py) for "
                                  "arguments with thousands of elements")
    n_rows, n_cols = sum(out_sizes), sum(in_sizes[a] for a in argnums)
    row0 = np.concatenate([[0], np.cumsum(out_sizes)]).astype(int)
    col0 = np.concatenate([[0], np.cumsum([in_sizes[a] for a in argnums])]).astype(int)

    tile: List[Optional[int]] = [None] * (n_rows * n_cols)
    for (oi, ai), (_, entries) in acc.blocks.items():
        for (i, j), node in entries.items():
            if not (0 <= i < out_sizes[oi] and 0 <= j < in_sizes[argnums[ai]]):
                raise AssertionError("edge entry outside its block")
            tile[(row0[oi] + i) * n_cols + col0[ai] + j] = node
    primal_nodes = [n for o in jaxpr.outvars for n in g.env[Graph.key(o)].nodes] if with_primal else []

    if contract:
        new = _contract_sums(dag, [n for n in tile if n is not None] + primal_nodes)
        tile = [None if n is None else new[n] for n in tile]
        primal_nodes = [new[n] for n in primal_nodes]
    roots = [n for n in tile if n is not None] + primal_nodes
    sched = _linearise(dag, g.first_elim_node, roots)
    if schedule == "pressure":
        sched = _reschedule_for_pressure(dag, sched, roots)
    elif schedule == "minpressure":
        sched = _reschedule_min_pressure(dag, sched, roots)
    elif schedule.startswith("minpressure:"):
        # "minpressure:<seed>[:<sigma>]": the same passes from a jittered copy of the order - their tie breaks follow the incoming
        # order, so every seed gives another schedule of about the same pressure (tools/tune.py --schedules: how much
        # of the kernel time is the compiler's luck with one particular order)
        parts = schedule.split(":")
        sched = _reschedule_min_pressure(dag, _random_topological_order(dag, sched, int(parts[1]), float(parts[2]) if len(parts) > 2 else 8.0), roots)
    elif schedule.startswith("ilp"):
        # "ilp:<live limit>:<window>"
        parts = schedule.split(":")
        sched = _reschedule_for_ilp(dag, sched, int(parts[1]) if len(parts) > 1 else 120,
                                    int(parts[2]) if len(parts) > 2 else 4)
    elif schedule != "elimination":
        raise ValueError("schedule must be 'pressure', 'minpressure[:seed]', 'ilp[:limit[:window]]' or 'elimination'")

    # stores directly after the value they read; constant / input / zero stores at the end
    stores_after: Dict[int, List[SsaOp]] = {}
    tail: List[SsaOp] = []
    zero = dag.const(0.0)
    for pos, node in enumerate(tile):
        op = SsaOp("store_jac", pos, (zero if node is None else node,))
        if node is None or dag.op[node] in ("const", "input"):
            tail.append(op)
        else:
            stores_after.setdefault(node, []).append(op)
    for r, node in enumerate(primal_nodes):
        op = SsaOp("store_primal", r, (node,))
        if dag.op[node] in ("const", "input"):
            tail.append(op)
        else:
            stores_after.setdefault(node, []).append(op)
    ssa: List[SsaOp] = []
    for n in sched:
        ssa.append(SsaOp(dag.op[n], n, dag.args[n]))
        ssa.extend(stores_after.get(n, ()))
    ssa.extend(tail)

    plan = Plan(in_sizes, out_sizes, argnums, n_rows, n_cols, list(acc.order), dag, ssa, tile, primal_nodes)
    plan.division = division
    plan.rounding = rounding
    _allocate_and_serialise(plan, with_primal)
    _collect_stats(plan, acc)
    return plan


def _allocate_and_serialise(plan: Plan, with_primal: bool) -> None:
    dag = plan.dag
    last_use: Dict[int, int] = {}
    for idx, op in enumerate(plan.ssa):
        for a in op.args:
            last_use[a] = idx
    consts: Dict[int, int] = {}
    slot_of: Dict[int, int] = {}
    free: List[int] = []
    n_slots = 0

    def operand(node: int) -> int:
        kind = dag.op[node]
        if kind == "const":
            idx = consts.setdefault(node, len(consts))
            return (K_CONST << 14) | idx
        if kind == "input":
            return (K_INPUT << 14) | dag.args[node][0]
        return (K_SLOT << 14) | slot_of[node]

    records = []
    for idx, op in enumerate(plan.ssa):
        enc = [operand(a) for a in op.args] + [0] * (3 - len(op.args))
        # operands whose last use is here release their slot before dst is allocated,
        # so dst may reuse one of them (the interpreter reads all operands first)
        for a in set(op.args):
            if dag.op[a] not in ("const", "input") and last_use[a] == idx:
                free.append(slot_of.pop(a))
        if op.op in ("store_jac", "store_primal"):
            dst = op.dst
        else:
            if op.dst not in last_use:
                raise AssertionError("dead value in plan")
            if free:
                free.sort(reverse=True)
                dst = free.pop()
            else:
                dst = n_slots
                n_slots += 1
            slot_of[op.dst] = dst
        records.append((OP[op.op], dst, enc[0], enc[1], enc[2]))
    if n_slots > MAX_INDEX or len(consts) > MAX_INDEX or plan.n_in_scalars > MAX_INDEX \
            or plan.n_rows * plan.n_cols > 0xFFFF:
        raise NotImplementedError("plan exceeds the 14-bit operand encoding")

    const_vals = np.zeros(len(consts), dtype=np.float32)
    for node, i in consts.items():
        const_vals[i] = dag.cval[node]
    flags = FLAG_HAS_PRIMAL if with_primal else 0
    body = b"".join([
        np.asarray(plan.in_sizes, dtype="<u4").tobytes(),
        np.asarray(plan.out_sizes, dtype="<u4").tobytes(),
        np.asarray(plan.in_cols, dtype="<u4").tobytes(),
        const_vals.astype("<f4").tobytes(),
        np.asarray([r + (0,) for r in records], dtype="<u2").reshape(-1, 6).tobytes() if records else b"",
    ])
    header = [MAGIC, VERSION, 64 + len(body), len(plan.in_sizes), len(plan.out_sizes), plan.n_in_scalars,
              plan.n_cols, plan.n_rows, len(records), n_slots, len(consts), flags, 0, 0, 0, 0]
    plan.blob = _struct.pack("<16I", *header) + body
    plan.n_slots = n_slots


def _collect_stats(plan: Plan, acc: Accumulated) -> None:
    hist: Dict[str, int] = {}
    for op in plan.ssa:
        hist[op.op] = hist.get(op.op, 0) + 1
    classes: Dict[str, int] = {}
    for s in acc.steps:
        for k, v in s.classes.items():
            classes[k] = classes.get(k, 0) + v
    plan.stats = {
        "n_vertices": len(acc.graph.jaxpr.eqns),
        "n_eliminated": len(acc.order),
        "n_ops": len(plan.ssa),
        "n_slots": plan.n_slots,
        "op_histogram": hist,
        # reference count_ops semantics; None when a tensor-level descriptor is unavailable (rank > 1 vertices)
        "num_muls": acc.num_muls if acc.structure_complete else None,
        "num_adds": acc.num_adds if acc.structure_complete else None,
        "order_counts": [(s.vertex, s.num_muls) for s in acc.steps] if acc.structure_complete else None,
        # structurally non-zero scalar products / additions: what the kernels execute before x1 / +0 are dropped;
        # defined at any rank (for vertices of rank <= 1 see num_muls / num_adds, the reference's own cost model)
        "scalar_muls": acc.scalar_muls, "scalar_adds": acc.scalar_adds,
        "scalar_order_counts": [(s.vertex, s.scalar_muls) for s in acc.steps],
        "pairs": sum(s.pairs for s in acc.steps),
        "max_pairs_per_step": max((s.pairs for s in acc.steps), default=0),
        "product_classes": classes,
        "bytes_per_jacobian": plan.bytes_per_jacobian,
    }
    plan.structure = {
        "elemental": [[src, dst, None if s is None else s.descriptor()] for src, dst, s in acc.graph.elemental],
        "blocks": {f"{oi},{ai}": (None if s is None else s.descriptor())
                   for (oi, ai), (s, _) in sorted(acc.blocks.items())},
        "order": list(acc.order),
    }
