"""Host-side logic (no device): tree bookkeeping and traversal order, Newick round trip, pattern
compression, tip encoding, model maths, the update queue of LkTree."""
import numpy as np
import pytest

from phyml_b200 import alignment, model as pmodel
from phyml_b200.tree import Tree


def test_random_tree_is_binary_and_tips_are_on_the_right():
    t = Tree.random(37, seed=3)
    assert t.n_edges == 71 and t.n_nodes == 72
    for e in range(t.n_edges):
        if t.is_tip(int(t.left[e])) or t.is_tip(int(t.rght[e])):
            assert t.is_tip(int(t.rght[e])) and not t.is_tip(int(t.left[e]))   # lk.c:3232


def test_newick_round_trip_preserves_topology_and_lengths():
    t = Tree.random(15, seed=4)
    t2 = Tree.from_newick(t.to_newick(precision=12), t.names)

    def splits(tr):
        out = {}
        for e in range(tr.n_edges):
            seen, stack = set(), [(int(tr.rght[e]), int(tr.left[e]))]
            while stack:
                v, p = stack.pop()
                if tr.is_tip(v):
                    seen.add(v)
                for (_, w) in tr.adj[v]:
                    if w != p:
                        stack.append((w, v))
            key = frozenset(seen) if 0 not in seen else frozenset(range(tr.n_otu)) - frozenset(seen)
            out[key] = float(tr.l[e])
        return out

    a, b = splits(t), splits(t2)
    assert set(a) == set(b)
    for k in a:
        assert abs(a[k] - b[k]) < 1e-10


def test_post_order_is_dependency_ordered_and_complete():
    t = Tree.random(40, seed=5)
    t.both_sides = True
    ops = t.full_traversal_ops()
    assert len(ops) == 3 * (t.n_otu - 2)
    written = set()
    for op in ops:
        for s in (op.c1, op.c2):
            assert s.is_tip or s.clv in written, "operand used before it was computed"
        assert op.dst not in written
        written.add(op.dst)
    # every internal side of every edge is covered exactly once
    expect = {t.clv_handle(e, v) for e in range(t.n_edges) for v in (int(t.left[e]), int(t.rght[e])) if not t.is_tip(v)}
    assert written == expect


def test_compress_counts_patterns_and_invariant_states():
    seqs = ["ACGTAC-A", "ACGTACNA", "ACTTACGA"]
    codes = alignment.encode(seqs, 4)
    pat = alignment.compress(codes, 4)
    assert pat.wght.sum() == 8 and pat.n_pattern < 8
    # column 0 = AAA is constant in state A; column 2 = GGT is polymorphic
    cols = ["".join(s[j] for s in alignment.decode(pat.codes, 4)) for j in range(pat.n_pattern)]
    assert pat.invar[cols.index("AAA")] == 0 and pat.invar[cols.index("GGT")] == -1
    # a gap / N column can be constant
    assert pat.invar[cols.index("NNG")] == 2
    lo, hi = alignment.shard_bounds(pat.n_pattern, 1, 2)
    sh = pat.shard(1, 2)
    assert sh.n_pattern == hi - lo and np.array_equal(sh.codes, pat.codes[:, lo:hi])


def test_tip_tables_follow_the_reference_tables():
    t4 = alignment.tip_table(4)
    assert t4.shape == (16, 4) and (t4[15] == 1).all() and list(t4[5]) == [1, 0, 1, 0]   # R = A|G (lk.c:42)
    t20 = alignment.tip_table(20)
    assert t20.shape == (21, 20) and (t20[20] == 1).all() and t20[:20].sum() == 20
    aa = alignment.encode(["ARNDBZX-"], 20)[0]
    assert list(aa) == [0, 1, 2, 3, 2, 5, 20, 20]                                         # B->N, Z->Q (lk.c:151-152)


def test_model_eigen_system_and_gamma():
    m = pmodel.gtr(alpha=0.5)
    Q = (m.U * m.lam) @ m.V
    assert np.allclose(Q.sum(axis=1), 0, atol=1e-12)                 # rows of a rate matrix sum to 0
    assert abs(-(m.pi * np.diag(Q)).sum() - 1.0) < 1e-12             # one expected substitution per unit time
    assert np.allclose(m.pi[:, None] * Q, (m.pi[:, None] * Q).T, atol=1e-12)   # reversibility
    P = m.pmat(0.3)
    assert np.allclose(P.sum(axis=2), 1.0) and (P > 0).all()
    r, w = pmodel.discrete_gamma(0.5, 4)
    assert abs(r.mean() - 1.0) < 1e-12 and np.all(np.diff(r) > 0) and np.allclose(w, 0.25)
    # reference values of the 4-category discrete gamma, alpha = 0.5 (Yang 1994, mean of categories)
    assert np.allclose(r, [0.03338775, 0.25191592, 0.82026848, 2.89442785], atol=2e-6)


def test_lktree_queues_updates_until_a_scalar_is_needed():
    import sys
    import os

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_backend import OracleBackend

    from phyml_b200.lk import LkTree

    tree = Tree.random(9, seed=8)
    m = pmodel.hky85(kappa=3.0, alpha=0.8)
    pat = alignment.compress(alignment.simulate(tree, m, 300, seed=9), 4)
    t = LkTree(tree, pat, m, OracleBackend(tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges))
    t.Set_Both_Sides(1)
    a = t.Lk()
    assert t.n_flush == 1                       # 3(n-2) Update_Partial_Lk calls -> one batched flush
    t.Update_Partial_Lk(tree.root_edge, int(tree.left[tree.root_edge]))
    assert len(t._queue) == 1 and t.n_flush == 1
    b = t.Lk(tree.root_edge)
    assert t.n_flush == 2 and abs(a - b) < 1e-9 * abs(a)
    with pytest.raises(ValueError):
        tree.partial_op(0, 0)                    # Update_Partial_Lk on a tip
