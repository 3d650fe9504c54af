"""Pairing of independent plan operations of ONE sample into f32x2 instructions (sm_100: FMUL2 / FADD2).

The reference rounding has no fused multiply-add: ``J[s,p] += J[s,v] * J[v,p]`` is a rounded product and a rounded
sum (core.py:210, 262-272), so a RoeFlux_3d sample is ~1500 FMUL / FADD, one issue slot each, and the kernel is
bound by issue slots, not by the FP32 pipe.  sm_100 has packed single-precision instructions that
apply the SAME correctly rounded operation to the two halves of 64-bit register pairs: ``mul.rn.f32x2`` and
``add.rn.f32x2`` give bit for bit what two ``mul.rn.f32`` / ``add.rn.f32`` give, in one issue slot.

The elimination is full of isomorphic, independent operation pairs: the products of one ``J[s,v]`` with the
entries ``J[v,p0]``, ``J[v,p1]`` of neighbouring columns, and the two sums they go into.  This module finds them
(superword-level parallelism, bottom-up from the 64-bit tile stores and top-down from the pairs found):

* a *pack* is an ordered pair of values that live in the halves of one aligned register pair; every value
  belongs to at most one pack, so that nothing ever has to be copied between pairs;
* two operations can become one packed instruction when they have the same kind (``mul``/``mul``, or any two of
  ``add``/``sub``: a subtraction is an addition with a negated operand half), neither depends on the other, and
  for each operand position the two operands are the same value (a broadcast operand), already a pack in that
  orientation, or both still free (they then become a pack themselves, and the search recurses into them);
* the contracted graph (packed pairs as single nodes) must stay acyclic; pairs that close a cycle are dropped.

Nothing here touches the plan: the result is an emission order for codegen.emit_body, and each half of a packed
instruction is the plan operation it replaces, rounded the same way - kernels stay bit-identical to the plan replay.
"""
from __future__ import annotations

import heapq
import sys
from typing import Dict, List, Optional, Tuple

from .plan import Plan, SsaOp

PACKABLE = ("mul", "add", "sub")


def _strip(dag, n: int) -> Tuple[int, bool]:
    neg = False
    while dag.op[n] == "neg":
        n, neg = dag.args[n][0], not neg
    return n, neg


class Pairing:
    """packed: list of (op0, op1, ((x0, x1), (y0, y1))) - two plan operations and their aligned operand pairs;
    order: emission list of ("op", SsaOp) / ("pack", index into packed)."""

    def __init__(self) -> None:
        self.packed: List[Tuple[SsaOp, SsaOp, Tuple[Tuple[int, int], Tuple[int, int]]]] = []
        self.order: List[Tuple[str, object]] = []
        self.mate: Dict[int, Tuple[int, int]] = {}


def pair_operations(plan: Plan) -> Pairing:
    dag = plan.dag
    ops = [o for o in plan.ssa if o.op not in ("store_jac", "store_primal")]
    by_dst = {o.dst: o for o in ops}
    idx = {o.dst: i for i, o in enumerate(ops)}
    anc: List[int] = []                       # bit sets of (transitive) operand operations
    for o in ops:
        s = 0
        for a in o.args:
            if a in idx:
                s |= anc[idx[a]] | (1 << idx[a])
        anc.append(s)

    def independent(a: int, b: int) -> bool:
        if a not in idx or b not in idx:
            return True
        return not (anc[idx[a]] >> idx[b]) & 1 and not (anc[idx[b]] >> idx[a]) & 1

    mate: Dict[int, Tuple[int, int]] = {}     # value -> (other half, own position 0 / 1)
    chosen: List[Tuple[int, int, tuple]] = []

    old_limit = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old_limit, 20000))

    def try_pack(v0: int, v1: int, tent: dict) -> bool:
        """Can (v0, v1) be the two halves of one register pair?  ``tent`` collects tentative mates / packed ops."""
        v0, v1 = _strip(dag, v0)[0], _strip(dag, v1)[0]
        if v0 == v1:
            return True                                        # one register, read as both halves
        k0, k1 = dag.op[v0], dag.op[v1]
        if k0 == "const" or k1 == "const":
            # an immediate pair costs a MOV per half and use: only the same magnitude in both halves is free
            return k0 == k1 and abs(float(dag.cval[v0])) == abs(float(dag.cval[v1]))
        cur0, cur1 = tent.get(v0, mate.get(v0)), tent.get(v1, mate.get(v1))
        if cur0 is not None or cur1 is not None:
            return cur0 == (v1, 0) and cur1 == (v0, 1)
        if not independent(v0, v1):
            return False
        tent[v0], tent[v1] = (v1, 0), (v0, 1)
        same = k0 == k1 and k0 in PACKABLE
        if same or {k0, k1} == {"add", "sub"}:
            a0, b0 = dag.args[v0]
            a1, b1 = dag.args[v1]
            alignments = [((a0, a1), (b0, b1))]
            if same and k0 in ("mul", "add"):
                alignments.append(((a0, b1), (b0, a1)))
            for x, y in alignments:
                t2 = dict(tent)
                t2["ops"] = list(tent.get("ops", ()))
                if try_pack(x[0], x[1], t2) and try_pack(y[0], y[1], t2):
                    t2["ops"].append((v0, v1, (x, y)))
                    tent.clear()
                    tent.update(t2)
                    return True
        return True        # a leaf pack: two scalar results written into the halves of one register pair

    def commit(tent: dict) -> int:
        new_ops = tent.pop("ops", [])
        mate.update(tent)
        chosen.extend(new_ops)
        return len(new_ops)

    # bottom-up from the tile stores of neighbouring columns (they leave as one 64-bit shared-memory store)
    stores = {o.dst: o.args[0] for o in plan.ssa if o.op == "store_jac"}
    for k in sorted(stores):
        if k % 2 == 0 and k + 1 in stores and (k // plan.n_cols) == ((k + 1) // plan.n_cols):
            tent: dict = {}
            if try_pack(stores[k], stores[k + 1], tent):
                commit(tent)
    # top-down: two free operations of the same kind that read the two halves of a pack
    progress = True
    while progress:
        progress = False
        readers: Dict[int, List[int]] = {}
        for o in ops:
            if o.dst in mate or o.op not in PACKABLE:
                continue
            for a in o.args:
                a = _strip(dag, a)[0]
                if a in mate:
                    readers.setdefault(a, []).append(o.dst)
        for v0, (v1, p) in list(mate.items()):
            if p != 0:
                continue
            for u0 in readers.get(v0, ()):
                if u0 in mate:
                    continue
                for u1 in readers.get(v1, ()):
                    if u1 in mate or u1 == u0:
                        continue
                    k0, k1 = dag.op[u0], dag.op[u1]
                    if not (k0 == k1 or {k0, k1} == {"add", "sub"}):
                        continue
                    tent = {}
                    if try_pack(u0, u1, tent) and tent.get("ops"):
                        commit(tent)
                        progress = True
                        break
    sys.setrecursionlimit(old_limit)

    # ---- ptxas (12.9, -O3) contracts ``mul.rn.f32x2`` feeding ``add.rn.f32x2`` into FFMA2 although both carry a
    # rounding modifier (--fmad=false does not stop it; tools/microbench/f32x2_contraction.cu).  Scalar producers
    # or consumers are left alone, so a packed sum must never read the pair a packed product wrote: of every such
    # (product pack, sum pack) conflict one side goes back to two scalar instructions, the one in most conflicts
    # first (a greedy vertex cover of the bipartite conflict graph).
    while True:
        mul_packs = {(d0, d1): i for i, (d0, d1, _) in enumerate(chosen) if dag.op[d0] == "mul"}
        degree: Dict[int, int] = {}
        for i, (d0, d1, (x, y)) in enumerate(chosen):
            if dag.op[d0] == "mul":
                continue
            for a0, a1 in (x, y):
                j = mul_packs.get((_strip(dag, a0)[0], _strip(dag, a1)[0]))
                if j is not None:
                    degree[i] = degree.get(i, 0) + 1
                    degree[j] = degree.get(j, 0) + 1
        if not degree:
            break
        drop = max(degree, key=lambda i: (degree[i], dag.op[chosen[i][0]] == "mul"))
        chosen = chosen[:drop] + chosen[drop + 1:]

    # ---- emission order: topological order of the contracted graph, original positions as priorities ----
    pos = {o.dst: i for i, o in enumerate(plan.ssa) if o.op not in ("store_jac", "store_primal")}
    while True:
        node_of: Dict[int, int] = {}                     # value -> contracted node (index into `nodes`)
        nodes: List[Tuple[int, ...]] = []
        for d0, d1, _ in chosen:
            node_of[d0] = node_of[d1] = len(nodes)
            nodes.append((d0, d1))
        for o in ops:
            if o.dst not in node_of:
                node_of[o.dst] = len(nodes)
                nodes.append((o.dst,))
        preds: List[set] = [set() for _ in nodes]
        for i, members in enumerate(nodes):
            for d in members:
                for a in by_dst[d].args:
                    if a in node_of and node_of[a] != i:
                        preds[i].add(node_of[a])
        succs: List[List[int]] = [[] for _ in nodes]
        indeg = [len(p) for p in preds]
        for i, p in enumerate(preds):
            for j in p:
                succs[j].append(i)
        key = [min(pos[d] for d in members) for members in nodes]
        heap = [(key[i], i) for i in range(len(nodes)) if indeg[i] == 0]
        heapq.heapify(heap)
        out: List[int] = []
        while heap:
            _, i = heapq.heappop(heap)
            out.append(i)
            for j in succs[i]:
                indeg[j] -= 1
                if indeg[j] == 0:
                    heapq.heappush(heap, (key[j], j))
        if len(out) == len(nodes):
            break
        # a cycle through packed pairs: dissolve the stuck pair whose halves lie furthest apart, try again
        stuck = [i for i in range(len(nodes)) if indeg[i] > 0 and len(nodes[i]) == 2]
        if not stuck:
            raise AssertionError("cycle without a packed pair")
        worst = max(stuck, key=lambda i: abs(pos[nodes[i][0]] - pos[nodes[i][1]]))
        chosen = [c for c in chosen if (c[0], c[1]) != nodes[worst]]

    result = Pairing()
    result.mate = mate
    pack_index: Dict[Tuple[int, int], int] = {}
    for d0, d1, operands in chosen:
        pack_index[(d0, d1)] = len(result.packed)
        result.packed.append((by_dst[d0], by_dst[d1], operands))
    stores_after: Dict[int, List[SsaOp]] = {}
    tail: List[SsaOp] = []
    last_value: Optional[int] = None
    for o in plan.ssa:
        if o.op in ("store_jac", "store_primal"):
            src = o.args[0]
            if src in node_of:
                stores_after.setdefault(src, []).append(o)
            else:
                tail.append(o)
    for i in out:
        members = nodes[i]
        if len(members) == 2:
            result.order.append(("pack", pack_index[members]))
        else:
            result.order.append(("op", by_dst[members[0]]))
        for d in members:
            for st in stores_after.get(d, ()):
                result.order.append(("op", st))
    for st in tail:
        result.order.append(("op", st))
    return result
