"""Graph builder and symbolic vertex elimination.

``build_graph`` mirrors the reference's ``_build_graph`` (``core.py:314-412``):
one pass over the equations that numbers vertices, evaluates the primal and
stores one edge per traced operand in ``graph[src][dst]`` /
``transpose[dst][src]`` (insertion-ordered dicts, like the reference's).
``eliminate`` mirrors ``_eliminate_vertex`` (``core.py:155-272``): for every
successor s (outer loop) and predecessor p (inner loop) of the vertex,
``J[s,p] (+)= J[s,v] . J[v,p]``, then the vertex's edges are deleted.

What differs from the reference is *when* the arithmetic happens: nothing is
computed here.  Every edge carries (i) the tensor-level structure descriptor
(``structure.Struct``) that ``count_ops`` and the sparsity classes are defined
on, and (ii) its structurally non-zero scalar entries as nodes of the expression
DAG.  Products append ``mul``/``fma``/``add`` nodes; the finished DAG is
linearised into the flat plan (``plan.py``).

Accumulation order (stated because fp32 results depend on it): within one
product the contraction index runs ascending, ``t = a0*b0; t = fma(ak, bk, t)``;
an existing edge is added last, ``new = t + old``, contracted to a single
``fma(a0, b0, old)`` when the product has one term.  Factors that are
structurally +-1 (add/sub/neg/slice/concatenate/reduce_sum edges) are not
multiplied: ``x*1`` and ``fma(x, 1, c) = x + c`` are exact identities.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Dict, List, Optional, Sequence, Set, Tuple, Union

from . import structure as st
from . import structure_nd as nd
import numpy as np

from .lowering import Dag, Entries, Vec, lower_eqn
from .tracer import Jaxpr, Literal, Var

Order = Union[str, Sequence[int]]


@dataclass
class Edge:
    struct: Optional[st.Struct]   # tensor-level descriptor; None when a vertex of rank > 1 or a shape op other than
                                  # slice / concatenate is involved (values stay exact, count_ops is unavailable)
    entries: Entries              # (i, j) -> DAG node; i over dst components, j over src components


@dataclass
class Graph:
    jaxpr: Jaxpr
    dag: Dag
    env: Dict[int, Vec]                                  # vertex key -> lowered primal value
    fwd: Dict[int, Dict[int, Edge]]                      # graph[src][dst]
    bwd: Dict[int, Dict[int, Edge]]                      # transpose_graph[dst][src]
    vo_vertices: Set[int]                                # outputs that are consumed later (core.py:387-390)
    elemental: List[Tuple[int, int, Optional[st.Struct]]] = field(default_factory=list)
    first_elim_node: int = 0

    @staticmethod
    def key(var: Var) -> int:
        """Inputs are -(k+1), equation results are their 1-based vertex id."""
        return -(var.input_index + 1) if var.input_index is not None else var.eqn_index + 1


def build_graph(jaxpr: Jaxpr, division: str = "ieee", fuse: bool = True) -> Graph:
    dag = Dag(fuse=fuse)
    env: Dict[int, Vec] = {}
    k = 0
    for var in jaxpr.invars:
        env[Graph.key(var)] = Vec(var.shape, [dag.input(k + i) for i in range(var.size)])
        k += var.size
    # closed-over arrays (core.py:374): sources of edges like the inputs, their values constants of the plan
    for var, val in zip(jaxpr.constvars, jaxpr.consts):
        env[Graph.key(var)] = Vec(var.shape, [dag.const(x) for x in np.asarray(val, dtype=np.float32).reshape(-1)])
    g = Graph(jaxpr, dag, env, {}, {}, set())
    for var in list(jaxpr.invars) + list(jaxpr.constvars):
        g.fwd[Graph.key(var)] = {}
        g.bwd[Graph.key(var)] = {}
    out_keys = {Graph.key(v) for v in jaxpr.outvars if isinstance(v, Var)}

    def read(atom) -> Vec:
        if isinstance(atom, Literal):
            return Vec((), [dag.const(atom.val)], is_literal=True)
        return env[Graph.key(atom)]

    for eqn in jaxpr.eqns:
        v = Graph.key(eqn.outvar)
        g.fwd.setdefault(v, {})
        g.bwd.setdefault(v, {})
        for a in eqn.invars:
            if isinstance(a, Var) and Graph.key(a) in out_keys and a.eqn_index is not None:
                g.vo_vertices.add(Graph.key(a))
        out, edges = lower_eqn(dag, eqn, read, division)
        if out.shape != eqn.outvar.shape:
            raise AssertionError(f"shape mismatch lowering {eqn!r}: {out.shape}")
        env[v] = out
        n_traced = sum(isinstance(a, Var) for a in eqn.invars)
        if len(edges) != n_traced:
            continue                                      # core.py:409: no edge at all
        for src_var, struct, entries in edges:
            src = Graph.key(src_var)
            if v in g.fwd[src]:
                # same variable used twice by one equation (x*x): the reference overwrites the
                # first edge (core.py:367-371); this engine accumulates
                old = g.fwd[src][v]
                merged = dict(old.entries)
                for ij, n in entries.items():
                    merged[ij] = dag.add(n, merged[ij]) if ij in merged else n
                both = struct is not None and old.struct is not None
                alg, pair = _algebra(struct, old.struct) if both else (st, [])
                edge = Edge(_try(alg.add, *pair) if both else None, merged)
            else:
                edge = Edge(struct, entries)
            g.fwd[src][v] = edge
            g.bwd[v][src] = edge
            g.elemental.append((src, v, edge.struct))
    g.first_elim_node = len(dag.op)
    return g


def eliminable_vertices(g: Graph) -> List[int]:
    out_keys = {Graph.key(v) for v in g.jaxpr.outvars if isinstance(v, Var)}
    return [i for i in range(1, len(g.jaxpr.eqns) + 1) if i not in out_keys or i in g.vo_vertices]


def checkify_order(order: Order, g: Graph) -> List[int]:
    """``_checkify_order`` (core.py:275-311), with strict validation of custom
    orders: unknown ids, output vertices and duplicates raise ``ValueError``
    instead of failing later with ``IndexError``/``KeyError``."""
    valid = eliminable_vertices(g)
    if isinstance(order, str):
        if order in ("forward", "fwd"):
            return valid
        if order in ("reverse", "rev"):
            return valid[::-1]
        raise ValueError(f"{order} is not a valid order identifier!")
    order = [int(o) for o in order]
    missing = set(valid).difference(order)
    if missing:
        raise ValueError(f"Supplied order is missing vertices {missing}!")
    extra = [o for o in order if o not in set(valid)]
    if extra:
        raise ValueError(f"Supplied order contains vertices that cannot be eliminated: {sorted(set(extra))}!")
    if len(set(order)) != len(order):
        raise ValueError("Supplied order eliminates a vertex more than once!")
    return order


def _try(fn, *structs):
    """Structure algebra outside what the reference's SparseTensor rules cover (a value-less shape-op edge that
    reaches an output, the same variable concatenated twice - the reference itself raises or returns a wrongly shaped
    block there): the scalar entries stay exact, the descriptor is dropped (``None`` = not tracked)."""
    try:
        return fn(*structs)
    except (AssertionError, NotImplementedError, IndexError, AttributeError):
        return None


def _strip(s):
    return replace(s, pre=()) if s.pre else s


def _algebra(*structs):
    """The structure algebra for these edges: ``structure.py`` while every vertex involved has rank <= 1,
    ``structure_nd.py`` as soon as one descriptor is an N-dimensional one (rank <= 1 descriptors are converted)."""
    if any(isinstance(s, nd.NStruct) for s in structs):
        return nd, [nd.from_struct(s) for s in structs]
    return st, list(structs)


def _product_entries(dag: Dag, post: Entries, pre: Entries, existing: Optional[Entries]) -> Entries:
    by_k_post: Dict[int, List[Tuple[int, int]]] = {}
    for (i, k), n in post.items():
        by_k_post.setdefault(k, []).append((i, n))
    terms: Dict[Tuple[int, int], List[Tuple[int, int, int]]] = {}
    for (k, j), m in pre.items():
        for i, n in by_k_post.get(k, ()):
            terms.setdefault((i, j), []).append((k, n, m))
    out: Entries = {}
    for ij in sorted(terms):
        ts = sorted(terms[ij])
        old = existing.get(ij) if existing is not None else None
        if old is not None and len(ts) == 1:
            out[ij] = dag.fma(ts[0][1], ts[0][2], old)
            continue
        acc = dag.mul(ts[0][1], ts[0][2])
        for _, n, m in ts[1:]:
            acc = dag.fma(n, m, acc)
        out[ij] = dag.add(acc, old) if old is not None else acc
    if existing is not None:
        for ij, n in existing.items():
            out.setdefault(ij, n)
    return out


def _count_terms(post: Entries, pre: Entries, existing: Optional[Entries]) -> Tuple[int, int]:
    """(scalar products, scalar additions onto an existing edge) of one J[s,p] (+)= J[s,v] J[v,p]."""
    by_k: Dict[int, List[int]] = {}
    for (i, k) in post:
        by_k.setdefault(k, []).append(i)
    n_mul, hit = 0, set()
    for (k, j) in pre:
        rows = by_k.get(k)
        if rows:
            n_mul += len(rows)
            if existing is not None:
                hit.update((i, j) for i in rows)
    return n_mul, (sum(1 for ij in hit if ij in existing) if existing is not None else 0)


@dataclass
class StepStats:
    vertex: int
    num_muls: int = 0
    num_adds: int = 0
    scalar_muls: int = 0          # structurally non-zero scalar products / additions: defined at any rank
    scalar_adds: int = 0
    pairs: int = 0
    classes: Dict[str, int] = field(default_factory=dict)
    structure_complete: bool = True



def _structured_product(g: Graph, stats: StepStats, post: Edge, pre: Edge, old: Optional[Edge], p: int, s: int) -> None:
    """One J[s,p] (+)= J[s,v] J[v,p] with the reference's SparseTensor structure rules (core.py:198-262)."""
    dag = g.dag
    alg, converted = _algebra(post.struct, pre.struct, *([old.struct] if old is not None else []))
    post_s, pre_s = converted[0], converted[1]
    pre_pending = pre_s.pre
    if post_s.pre and pre_s.has_val:
        # unload_pre_transforms: shape-op Jacobians pending on the outgoing edge act on the incoming one
        pre_s = _strip(pre_s)
        for t in post_s.pre:
            pre_s = alg.apply_transform(t, pre_s)
    if pre_s.has_val and post_s.has_val:
        lhs, rhs = _strip(post_s), _strip(pre_s)
        cls = alg.mul_class(lhs, rhs)
        key = f"{post.struct.sparsity_class()}x{pre.struct.sparsity_class()}:{cls}"
        stats.classes[key] = stats.classes.get(key, 0) + 1
        out_s = alg.mul(lhs, rhs)
        stats.num_muls += alg.num_muls(lhs, rhs)
        pending = pre_pending
    elif pre_s.has_val:
        out_s, pending = _strip(pre_s), pre_pending
    else:
        out_s, pending = _strip(post_s), pre_pending + post_s.pre
    out_s = replace(out_s, pre=pending)

    if old is not None:
        # fill-in onto an existing edge: flush deferred transforms on both, then add
        new_m, old_m = alg.materialise(out_s), alg.materialise(converted[2])
        out_s = alg.add(new_m, old_m)
        stats.num_adds += alg.num_adds(out_s, old_m)
    entries = _product_entries(dag, post.entries, pre.entries, None if old is None else old.entries)
    edge = Edge(out_s, entries)
    g.fwd[p][s] = edge
    g.bwd[s][p] = edge


def eliminate(g: Graph, vertex: int) -> StepStats:
    """Eliminate one vertex (core.py:155-272). Returns the reference's op counts."""
    stats = StepStats(vertex)
    dag = g.dag
    v = vertex
    for s in list(g.fwd[v].keys()):
        post = g.fwd[v][s]
        for p in list(g.bwd[v].keys()):
            pre = g.bwd[v][p]
            stats.pairs += 1
            old = g.fwd[p].get(s)
            m, a = _count_terms(post.entries, pre.entries, None if old is None else old.entries)
            stats.scalar_muls += m
            stats.scalar_adds += a
            if post.struct is None or pre.struct is None or (old is not None and old.struct is None):
                stats.structure_complete = False
                entries = _product_entries(dag, post.entries, pre.entries, None if old is None else old.entries)
                edge = Edge(None, entries)
                g.fwd[p][s] = edge
                g.bwd[s][p] = edge
                continue
            try:
                _structured_product(g, stats, post, pre, old, p, s)
            except (AssertionError, NotImplementedError, IndexError, AttributeError):
                # outside the reference's structure rules (it raises there): keep the exact entries only
                stats.structure_complete = False
                entries = _product_entries(dag, post.entries, pre.entries, None if old is None else old.entries)
                edge = Edge(None, entries)
                g.fwd[p][s] = edge
                g.bwd[s][p] = edge
            continue

    keep_in_edges = v in g.vo_vertices
    if not keep_in_edges:
        for p in g.bwd[v]:
            del g.fwd[p][v]
    for s in g.fwd[v]:
        del g.bwd[s][v]
    del g.fwd[v]
    if not keep_in_edges:
        del g.bwd[v]
    return stats


@dataclass
class Accumulated:
    """Result of eliminating a whole order."""
    graph: Graph
    order: List[int]
    steps: List[StepStats]
    # blocks[(out_idx, arg_idx)] = (struct or None, entries) in flush-ed (dense-coordinate) form
    blocks: Dict[Tuple[int, int], Tuple[Optional[st.Struct], Entries]]

    @property
    def structure_complete(self) -> bool:
        return all(s.structure_complete for s in self.steps) and all(s is not None for _, _, s in self.graph.elemental)

    @property
    def num_muls(self) -> int:
        return sum(s.num_muls for s in self.steps)

    @property
    def num_adds(self) -> int:
        return sum(s.num_adds for s in self.steps)

    @property
    def scalar_muls(self) -> int:
        return sum(s.scalar_muls for s in self.steps)

    @property
    def scalar_adds(self) -> int:
        return sum(s.scalar_adds for s in self.steps)


def accumulate(jaxpr: Jaxpr, order: Order, argnums: Sequence[int], division: str = "ieee",
               fuse: bool = True) -> Accumulated:
    """``vertex_elimination_jaxpr`` minus densification (core.py:522-562).  ``fuse=False``: no FMA nodes, every
    ``J[s,p] += J[s,v] J[v,p]`` is a rounded product followed by a rounded sum (the reference's own rounding)."""
    g = build_graph(jaxpr, division, fuse)
    for o in jaxpr.outvars:
        if not isinstance(o, Var) or o.eqn_index is None:
            raise NotImplementedError("outputs that are inputs or constants are not supported")
    vertices = checkify_order(order, g)
    steps = [eliminate(g, v) for v in vertices]
    blocks: Dict[Tuple[int, int], Tuple[Optional[st.Struct], Entries]] = {}
    for oi, o in enumerate(jaxpr.outvars):
        for ai, a in enumerate(argnums):
            src, dst = Graph.key(jaxpr.invars[a]), Graph.key(o)
            edge = g.fwd.get(src, {}).get(dst)
            if edge is None:
                blocks[(oi, ai)] = (None, {})
            else:
                alg = nd if isinstance(edge.struct, nd.NStruct) else st
                blocks[(oi, ai)] = (_try(alg.materialise, edge.struct) if edge.struct is not None else None, edge.entries)
    return Accumulated(g, vertices, steps, blocks)
