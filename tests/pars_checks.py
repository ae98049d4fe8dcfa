"""Parsimony parity checks shared by the CPU suite (oracle vs the reference's dumps) and the GPU suite (CUDA engine
through the C ABI vs the same dumps and vs the oracle).  Goldens: tests/golden/pars/*.npz, dumped from the
unmodified reference's Pars / Update_Partial_Pars / Pars_Core (src/pars.c) by tests/golden/make_golden_pars.py.
All comparisons are exact (integer work)."""
import os

import numpy as np

from golden_case import GoldenCase
from oracle_backend import GOLDEN_DIR

PARS_CASES = ["nucleic_hky", "proteic_lg", "synth_aa_small", "synth_dna_deep"]


def load(case):
    c = GoldenCase(case)
    g = dict(np.load(os.path.join(GOLDEN_DIR, "pars", case + ".npz")))
    return c, g


def setup(c, g, eng, general):
    """Weights + tips, as Make_Tree_For_Pars leaves them (Init_Ui_Tips / Init_Partial_Pars_Tips, src/pars.c:111-233)."""
    eng.set_weights(c.g["wght"], c.g["invar"])
    eng.pars_create(c.tree.n_clv_handles, g["step_mat"] if general else None)
    for i in range(c.n_otu):
        h = c.tree.pars_tip_handle(i)
        if general:
            eng.pars_set_buffer(h, p_pars=g["p_pars"][h])
        else:
            eng.pars_set_buffer(h, ui=g["ui"][h], pars=g["pars"][h])


def full_ops(c):
    """Pars(NULL) with both_sides == YES: Post_Order_Pars then Pre_Order_Pars from a_nodes[0] (src/pars.c:33-37)."""
    t = c.tree
    a = 0
    d = t.adj[a][0][1]
    return t.pars_ops(t.post_order_ops(a, d)), t.pars_ops(t.pre_order_ops(a, d)), t.adj[a][0][0]


def check_full(c, g, eng, general, fused):
    setup(c, g, eng, general)
    post, pre, e0 = full_ops(c)
    sfx = "_general" if general else ""
    if fused:  # the whole of Pars(NULL) as one call
        c_pars = eng.pars_traverse_edge(post + pre, 2 * e0, 2 * e0 + 1, general)
    else:
        eng.pars_update(post, general)
        eng.pars_update(pre, general)
        c_pars = eng.pars_edge(2 * e0, 2 * e0 + 1, general)
    assert c_pars == int(g["c_pars" + sfx])
    assert (eng.get_site_pars() == g["site_pars" + sfx]).all()
    for h in range(c.tree.n_clv_handles):
        if general:
            assert (eng.pars_get_buffer(h, True) == g["p_pars"][h]).all(), h
        else:
            ui, pars = eng.pars_get_buffer(h)
            assert (ui == g["ui"][h]).all() and (pars == g["pars"][h]).all(), h
    # Pars(b) at every edge (src/pars.c:39-48)
    for e in range(c.tree.n_edges):
        assert eng.pars_edge(2 * e, 2 * e + 1, general) == int(g["edge_pars" + sfx][e]), e


def check_single_updates(c, g, eng, general):
    """one Update_Partial_Pars at a time with the children taken from the reference"""
    setup(c, g, eng, general)
    post, pre, _ = full_ops(c)
    for (dst, c1, c2) in post + pre:
        for h in (c1, c2):
            if general:
                eng.pars_set_buffer(h, p_pars=g["p_pars"][h])
            else:
                eng.pars_set_buffer(h, ui=g["ui"][h], pars=g["pars"][h])
        eng.pars_update([(dst, c1, c2)], general)
        if general:
            assert (eng.pars_get_buffer(dst, True) == g["p_pars"][dst]).all()
        else:
            ui, pars = eng.pars_get_buffer(dst)
            assert (ui == g["ui"][dst]).all() and (pars == g["pars"][dst]).all()


def random_case(n_otu, P, ns, seed, frac_weights=False):
    """seeded synthetic inputs: random tree, random tip state sets (a few ambiguous), pattern weights"""
    from phyml_b200.tree import Tree

    rng = np.random.default_rng(seed)
    tree = Tree.random(n_otu, seed=seed)
    st = rng.integers(0, ns, size=(n_otu, P))
    ui = (1 << st).astype(np.int32)
    amb = rng.random((n_otu, P)) < 0.03
    ui[amb] |= (1 << rng.integers(0, ns, size=int(amb.sum()))).astype(np.int32)
    w = rng.integers(0, 6, size=P).astype(np.float64)
    if frac_weights:
        w = w + np.round(rng.random(P), 3)
    step = rng.integers(1, 4, size=(ns, ns)).astype(np.int32)
    step = np.minimum(step, step.T)
    np.fill_diagonal(step, 0)
    return tree, ui, w, step


def run_random(tree, ui, w, step, eng, general, split=False):
    ns = eng.ns
    P = ui.shape[1]
    eng.set_weights(w, np.full(P, -1, dtype=np.int16))
    eng.pars_create(tree.n_clv_handles, step if general else None)
    for i in range(tree.n_otu):
        h = tree.pars_tip_handle(i)
        if general:
            bits = (ui[i][:, None] >> np.arange(ns)[None, :]) & 1
            eng.pars_set_buffer(h, p_pars=np.where(bits > 0, 0, 1000000000).astype(np.int32))
        else:
            eng.pars_set_buffer(h, ui=ui[i], pars=np.zeros(P, dtype=np.int32))
    a, d = 0, tree.adj[0][0][1]
    e0 = tree.adj[0][0][0]
    ops = tree.pars_ops(tree.post_order_ops(a, d)) + tree.pars_ops(tree.pre_order_ops(a, d))
    if split:
        eng.pars_update(ops, general)
        c_pars = eng.pars_edge(2 * e0, 2 * e0 + 1, general)
    else:
        c_pars = eng.pars_traverse_edge(ops, 2 * e0, 2 * e0 + 1, general)
    return c_pars, eng.get_site_pars(), [eng.pars_edge(2 * e, 2 * e + 1, general) for e in range(0, tree.n_edges, 7)]
