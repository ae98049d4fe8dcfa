"""Unrooted binary tree bookkeeping for the likelihood engine.

Host-side mirror of the part of PhyML's ``t_tree``/``t_edge``/``t_node`` that the likelihood hot
path reads (reference: src/utilities.h:635-830).  Conventions kept from the reference so that the
traversal code below reads like src/lk.c:

* tips are nodes ``0..n_otu-1``, internal nodes ``n_otu..2*n_otu-3`` (make.c / Read_Tree);
* edges are ``0..2*n_otu-4``; every edge has a ``left`` and a ``rght`` node and a tip, if any,
  is always on the right ("tip always on the right", src/lk.c:3232, src/beagle_utils.h:24-28);
* each edge owns two conditional-likelihood (CLV) buffers, ``p_lk_left`` and ``p_lk_rght``
  (src/utilities.h:759-762); here they are integer handles ``2*edge`` and ``2*edge+1`` into the
  device engine's buffer pool, and each edge's P-matrix handle is the edge number.

Only what the hot path needs is here; tree search (spr.c) and optimisation (optimiz.c) stay in the
reference's C and reach the engine through the C-ABI shim.
"""
from __future__ import annotations

import dataclasses
import re
from typing import List, Optional, Sequence, Tuple

import numpy as np

LEFT, RGHT = 0, 1


@dataclasses.dataclass
class Side:
    """One operand of an update: a tip (``tip`` >= 0) or an internal CLV buffer (``clv`` >= 0)."""

    tip: int = -1
    clv: int = -1

    @property
    def is_tip(self) -> bool:
        return self.tip >= 0


@dataclasses.dataclass
class PartialOp:
    """One ``Update_Partial_Lk(tree, b, d)`` resolved to buffer handles (cf. Set_All_Partial_Lk,
    src/lk.c:2922-3195): dst = CLV of edge ``b`` on the side of node ``d``; children are the two
    other edges of ``d`` with the far-side buffers/tips and their P-matrices."""

    dst: int
    c1: Side
    pmat1: int
    c2: Side
    pmat2: int
    edge: int = -1
    node: int = -1


class Tree:
    def __init__(self, n_otu: int, edges: Sequence[Tuple[int, int]], lengths: Sequence[float],
                 names: Optional[Sequence[str]] = None):
        self.n_otu = int(n_otu)
        self.n_nodes = 2 * self.n_otu - 2
        self.n_edges = 2 * self.n_otu - 3
        if len(edges) != self.n_edges:
            raise ValueError(f"expected {self.n_edges} edges, got {len(edges)}")
        self.left = np.zeros(self.n_edges, dtype=np.int64)
        self.rght = np.zeros(self.n_edges, dtype=np.int64)
        for e, (a, b) in enumerate(edges):
            # a tip is always on the right of its edge
            if a < self.n_otu and b >= self.n_otu:
                a, b = b, a
            self.left[e], self.rght[e] = a, b
        self.l = np.asarray(lengths, dtype=np.float64).copy()
        self.names = list(names) if names is not None else [f"t{i}" for i in range(self.n_otu)]
        # node -> list of (edge, neighbour) in insertion order (the reference's d->b[i], d->v[i])
        self.adj: List[List[Tuple[int, int]]] = [[] for _ in range(self.n_nodes)]
        for e in range(self.n_edges):
            a, b = int(self.left[e]), int(self.rght[e])
            self.adj[a].append((e, b))
            self.adj[b].append((e, a))
        for v in range(self.n_nodes):
            deg = len(self.adj[v])
            if (v < self.n_otu and deg != 1) or (v >= self.n_otu and deg != 3):
                raise ValueError(f"node {v} has degree {deg}: not an unrooted binary tree")
        self.tip_root = 0
        self.both_sides = False

    # ------------------------------------------------------------------ handles
    def is_tip(self, node: int) -> bool:
        return node < self.n_otu

    def clv_handle(self, edge: int, node: int) -> int:
        """CLV buffer of ``edge`` on the side where ``node`` lies (p_lk_left / p_lk_rght)."""
        if node == self.left[edge]:
            return 2 * edge + LEFT
        if node == self.rght[edge]:
            return 2 * edge + RGHT
        raise ValueError("node not on edge")

    @property
    def n_clv_handles(self) -> int:
        return 2 * self.n_edges

    def side_of(self, edge: int, node: int) -> Side:
        """Operand describing the subtree hanging off ``edge`` on ``node``'s side."""
        if self.is_tip(node):
            return Side(tip=node)
        return Side(clv=self.clv_handle(edge, node))

    # ------------------------------------------------------------------ Update_Partial_Lk
    def partial_op(self, edge: int, d: int) -> PartialOp:
        """Resolve Update_Partial_Lk(tree, edge, d) (src/lk.c:1282, Set_All_Partial_Lk :2937-2987)."""
        if self.is_tip(d):
            raise ValueError("Update_Partial_Lk on a tip is a no-op in the reference (lk.c:1297)")
        others = [(e, v) for (e, v) in self.adj[d] if e != edge]
        assert len(others) == 2
        (e1, v1), (e2, v2) = others
        return PartialOp(dst=self.clv_handle(edge, d), c1=self.side_of(e1, v1), pmat1=e1,
                         c2=self.side_of(e2, v2), pmat2=e2, edge=edge, node=d)

    def post_order_ops(self, a: Optional[int] = None, d: Optional[int] = None) -> List[PartialOp]:
        """Post_Order_Lk(a, d, tree) (src/lk.c:282-352) as a dependency-ordered list of updates.
        Defaults to the traversal of Lk(NULL): a = tip_root, d = its only neighbour (lk.c:562)."""
        if a is None:
            a = self.tip_root
            d = self.adj[a][0][1]
        ops: List[PartialOp] = []
        # iterative post-order (the reference recurses; the order of updates is identical)
        stack = [(a, d, False)]
        while stack:
            pa, pd, done = stack.pop()
            if self.is_tip(pd):
                continue
            if done:
                e = next(e for (e, v) in self.adj[pd] if v == pa)
                ops.append(self.partial_op(e, pd))
            else:
                stack.append((pa, pd, True))
                for (e, v) in reversed(self.adj[pd]):
                    if v != pa:
                        stack.append((pd, v, False))
        return ops

    def pre_order_ops(self, a: Optional[int] = None, d: Optional[int] = None) -> List[PartialOp]:
        """Pre_Order_Lk(a, d, tree) (src/lk.c:357-393): the 'down' partials."""
        if a is None:
            a = self.tip_root
            d = self.adj[a][0][1]
        ops: List[PartialOp] = []
        stack = [(a, d)]
        while stack:
            pa, pd = stack.pop()
            if self.is_tip(pd):
                continue
            nxt = []
            for (e, v) in self.adj[pd]:
                if v != pa:
                    ops.append(self.partial_op(e, pd))
                    nxt.append((pd, v))
            # depth-first in the reference's order: first child fully before the second.  Updates of
            # d's own out-going edges only depend on d's in-coming partial, so emitting both before
            # descending is dependency-equivalent.
            for item in reversed(nxt):
                stack.append(item)
        return ops

    def full_traversal_ops(self) -> List[PartialOp]:
        """Update_All_Partial_Lk for an unrooted tree (src/lk.c:432-437)."""
        ops = self.post_order_ops()
        if self.both_sides:
            ops += self.pre_order_ops()
        return ops

    # ------------------------------------------------------------------ Update_Partial_Pars
    def pars_ops(self, ops: Sequence[PartialOp]) -> List[Tuple[int, int, int]]:
        """The same traversal as parsimony updates (Update_Partial_Pars, src/pars.c:239-351): every edge side,
        tips included, owns a parsimony buffer; handle = 2*edge + side (a tip is always on the right)."""
        def h(c: Side, e: int) -> int:
            return c.clv if not c.is_tip else 2 * e + RGHT
        return [(o.dst, h(o.c1, o.pmat1), h(o.c2, o.pmat2)) for o in ops]

    def pars_tip_handle(self, tip: int) -> int:
        """Buffer of a tip: ui_r / p_pars_r of its only edge (Init_Ui_Tips, src/pars.c:164-233)."""
        return 2 * self.adj[tip][0][0] + RGHT

    @property
    def root_edge(self) -> int:
        """Edge at which Lk(NULL) sums site likelihoods: a_nodes[tip_root]->b[0] (src/lk.c:578-579)."""
        return self.adj[self.tip_root][0][0]

    def edge_sides(self, edge: int) -> Tuple[Side, Side]:
        """(left, right) operands of an edge-likelihood evaluation (src/lk.c:605-606)."""
        a, b = int(self.left[edge]), int(self.rght[edge])
        return self.side_of(edge, a), self.side_of(edge, b)

    # ------------------------------------------------------------------ I/O
    def to_newick(self, precision: int = 10) -> str:
        root = self.n_otu  # any internal node
        if self.n_otu == 2:
            return f"({self.names[0]}:{self.l[0]:.{precision}f},{self.names[1]}:0.0);"

        def rec(node: int, parent: int) -> str:
            if self.is_tip(node):
                return self.names[node]
            parts = []
            for (e, v) in self.adj[node]:
                if v != parent:
                    parts.append(f"{rec(v, node)}:{self.l[e]:.{precision}f}")
            return "(" + ",".join(parts) + ")"

        return rec(root, -1) + ";"

    @staticmethod
    def from_newick(text: str, names: Optional[Sequence[str]] = None) -> "Tree":
        """Parse an unrooted (trifurcating root) or rooted-binary Newick string.  Tip numbering
        follows ``names`` if given (the alignment's taxon order), else order of appearance."""
        toks = re.findall(r"[(),;]|:[^(),;:]+|[^(),;:\s]+", text.strip())
        pos = 0
        children: List[List[Tuple[int, float]]] = []  # per temp node: list of (child, length)
        labels: List[Optional[str]] = []

        def new_node(label=None):
            children.append([])
            labels.append(label)
            return len(children) - 1

        def parse() -> Tuple[int, float]:
            nonlocal pos
            if toks[pos] == "(":
                pos += 1
                me = new_node()
                while True:
                    ch, ln = parse()
                    children[me].append((ch, ln))
                    if toks[pos] == ",":
                        pos += 1
                        continue
                    if toks[pos] == ")":
                        pos += 1
                        break
                    raise ValueError("bad newick")
                if pos < len(toks) and toks[pos] not in "(),;:" and not toks[pos].startswith(":"):
                    pos += 1  # internal label / support value
            else:
                me = new_node(toks[pos])
                pos += 1
            ln = 0.0
            if pos < len(toks) and toks[pos].startswith(":"):
                ln = float(toks[pos][1:])
                pos += 1
            return me, ln

        root, _ = parse()
        # collapse a bifurcating root into a single edge
        if len(children[root]) == 2:
            (a, la), (b, lb) = children[root]
            if children[a]:
                children[a].append((b, la + lb))
                root = a
            elif children[b]:
                children[b].append((a, la + lb))
                root = b
            else:
                raise ValueError("two-taxon trees are not supported")
        tip_labels = [lab for lab, ch in zip(labels, children) if not ch]
        if names is None:
            names = tip_labels
        name_to_id = {nm: i for i, nm in enumerate(names)}
        if sorted(tip_labels) != sorted(names):
            raise ValueError("tree tips do not match taxon names")
        n_otu = len(names)
        ids = {}
        next_int = n_otu
        order = []
        stack = [root]
        while stack:
            v = stack.pop()
            order.append(v)
            for (c, _) in children[v]:
                stack.append(c)
        for v in order:
            if children[v]:
                ids[v] = next_int
                next_int += 1
            else:
                ids[v] = name_to_id[labels[v]]
        edges, lens = [], []
        for v in order:
            if len(children[v]) not in (0, 2, 3) or (v != root and len(children[v]) == 3):
                raise ValueError("tree is not binary")
            for (c, ln) in children[v]:
                edges.append((ids[v], ids[c]))
                lens.append(ln)
        return Tree(n_otu, edges, lens, names)

    @staticmethod
    def random(n_otu: int, seed: int = 1, mean_bl: float = 0.1, bl_min: float = 1e-4,
               bl_max: float = 1.0) -> "Tree":
        """Random unrooted binary topology by stepwise addition, branch lengths ~ Exp(mean_bl)
        clamped to [bl_min, bl_max]."""
        rng = np.random.default_rng(seed)
        if n_otu < 3:
            raise ValueError("need at least 3 taxa")
        edges: List[List[int]] = [[n_otu, 0], [n_otu, 1], [n_otu, 2]]
        next_int = n_otu + 1
        for tip in range(3, n_otu):
            e = int(rng.integers(len(edges)))
            a, b = edges[e]
            mid = next_int
            next_int += 1
            edges[e] = [a, mid]
            edges.append([mid, b])
            edges.append([mid, tip])
        lens = np.clip(rng.exponential(mean_bl, size=len(edges)), bl_min, bl_max)
        return Tree(n_otu, [tuple(e) for e in edges], lens)
