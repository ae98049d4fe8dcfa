"""plk_spr_candidates: the scores of many SPR regraft positions from ONE call must equal
what the reference's per-candidate sequence gives (Test_One_Spr_Target, src/spr.c:589-650: two Update_PMat_At_Given_Edge,
one Update_Partial_Lk at the new node, one Lk(b_arrow)) -- checked against that sequence run call by call on the engine
and against the oracle.  `-m gpu`."""
import numpy as np
import pytest

from oracle_backend import OracleBackend
from phyml_b200.engine import Engine, EngineError
from phyml_b200.lk import LkTree
from phyml_b200.tree import PartialOp, Side
from test_gpu_parity import _synthetic

pytestmark = pytest.mark.gpu


def _setup(ns, n_taxa, n_sites, ncatg, pinv, mean_bl, devices=None):
    tree, m, pat = _synthetic(ns, n_taxa, n_sites, 9, 0.03, ncatg, pinv, mean_bl)
    args = (tree.n_otu, pat.n_pattern, ns, ncatg, tree.n_clv_handles + 1, tree.n_edges + 3)
    gpu = LkTree(tree, pat, m, Engine(*args, devices=devices))
    cpu = LkTree(tree, pat, m, OracleBackend(*args))
    for t in (gpu, cpu):
        t.Set_Both_Sides(1)
        t.Lk()
    return tree, gpu, cpu


def _candidates(tree, skip_edge):
    """both ends of every other edge, each half as long as the edge (Graft_Subtree splits the target edge)"""
    out = []
    for e in range(tree.n_edges):
        if e == skip_edge:
            continue
        a, b = tree.edge_sides(e)
        out.append((a, 0.5 * tree.l[e], b, 0.5 * tree.l[e] + 0.01))
    return out


def _sequential(tree, eng, prune, l_prune, link_on_left, cands):
    tmp, ha, hb, hp = tree.n_clv_handles, tree.n_edges, tree.n_edges + 1, tree.n_edges + 2
    out = []
    for a, la, b, lb in cands:
        eng.update_pmats([ha, hb, hp], [la, lb, l_prune])
        eng.update_partials([PartialOp(dst=tmp, c1=a, pmat1=ha, c2=b, pmat2=hb)])
        x = Side(clv=tmp)
        out.append(eng.edge_lnl(x, prune, hp) if link_on_left else eng.edge_lnl(prune, x, hp))
    return np.array(out)


@pytest.mark.parametrize("ns,n_taxa,n_sites,ncatg,pinv,mean_bl,devices", [
    (4, 20, 900, 4, 0.0, 0.2, None),
    (4, 20, 900, 4, 0.15, 0.2, None),     # +I
    (4, 160, 90, 4, 0.0, 0.45, None),     # deep tree: the new node's CLV is rescaled by 2^256 inside the kernel
    (4, 14, 500, 3, 0.0, 0.2, None),      # generic category count
    (20, 16, 300, 4, 0.0, 0.2, None),     # 20 states
    (4, 20, 900, 4, 0.0, 0.2, [0, 0, 0]),  # sharded instance: all-shard sums
])
def test_batched_candidates_match_the_sequence(ns, n_taxa, n_sites, ncatg, pinv, mean_bl, devices):
    tree, gpu, cpu = _setup(ns, n_taxa, n_sites, ncatg, pinv, mean_bl, devices)
    # pruned subtree = a tip (always the right-hand side) and an internal subtree on either side of its edge
    tip_edge = tree.adj[3][0][0]
    internal_edge = next(e for e in range(tree.n_edges) if tree.left[e] >= tree.n_otu and tree.rght[e] >= tree.n_otu)
    cases = [(tree.side_of(tip_edge, 3), tip_edge, True),
             (tree.edge_sides(internal_edge)[1], internal_edge, True),
             (tree.edge_sides(internal_edge)[0], internal_edge, False)]
    for prune, e, link_on_left in cases:
        cands = _candidates(tree, e)
        l_prune = float(tree.l[e])
        got, warn = gpu.eng.spr_candidates(prune, l_prune, link_on_left, cands)
        seq = _sequential(tree, gpu.eng, prune, l_prune, link_on_left, cands)
        ref, _ = cpu.eng.spr_candidates(prune, l_prune, link_on_left, cands)
        assert (np.abs(got - seq) <= 1e-13 * np.abs(seq)).all(), np.abs(got / seq - 1).max()
        assert (np.abs(got - ref) <= 1e-12 * np.abs(ref)).all(), np.abs(got / ref - 1).max()
        assert not warn.any()


def test_many_candidates_and_errors():
    tree, gpu, cpu = _setup(4, 12, 400, 4, 0.0, 0.2)
    e = tree.adj[0][0][0]
    prune = tree.side_of(e, 0)
    cands = _candidates(tree, e) * 130  # > one chunk of 2048 candidates
    got, _ = gpu.eng.spr_candidates(prune, 0.1, True, cands)
    one, _ = gpu.eng.spr_candidates(prune, 0.1, True, cands[:len(cands) // 130])
    assert (got.reshape(130, -1) == one[None, :]).all()  # deterministic, independent of the batch position
    with pytest.raises(EngineError):
        gpu.eng.spr_candidates(prune, 0.1, False, cands[:2])  # a tip cannot be the left-hand side
    with pytest.raises(EngineError):
        gpu.eng.spr_candidates(Side(clv=tree.n_clv_handles), 0.1, True, cands[:2])  # never-written handle
    assert len(gpu.eng.spr_candidates(prune, 0.1, True, [])[0]) == 0
