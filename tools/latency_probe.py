#!/usr/bin/env python
"""Latency of the calls one SPR candidate / one Br_Len_Opt step costs (BASELINE config 5 shape:
100 taxa x 50 000 sites, GTR+G4), through the C ABI: the numbers that bound spr.c / optimiz.c when
they drive the engine (latency-, not bandwidth-bound)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from phyml_b200 import workloads as wl
from phyml_b200.engine import Engine, pack_ops
from phyml_b200.lk import LkTree

name = sys.argv[1] if len(sys.argv) > 1 else "dna_100x50k"
w = wl.WORKLOADS[name]
m, _pin = wl.evaluation_model(name)
tree = wl.make_tree(w)
pat = wl.make_patterns(name, wl.rank_blocks(w, 0, 1) if w.n_blocks > 1 else [0], procs=8)
desc = w.desc
eng = Engine(tree.n_otu, pat.n_pattern, m.ns, m.ncatg, tree.n_clv_handles, tree.n_edges)
t = LkTree(tree, pat, m, eng)
t.Set_Both_Sides(1)
print(desc, "patterns", pat.n_pattern, "lnL", t.Lk())


def timeit(fn, n=300, warm=20):
    for _ in range(warm):
        fn()
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    eng.sync()
    return (time.perf_counter() - t0) / n * 1e6


e = tree.n_edges // 2
d = int(tree.left[e])
op = pack_ops([tree.partial_op(e, d)])
left, rght = tree.edge_sides(e)
L = np.array([tree.l[e]])
E = np.array([e], dtype=np.int32)
res = {}
res["edge_lnl (Lk(b) without PMat)"] = timeit(lambda: eng.edge_lnl(left, rght, e))
res["update_pmat + edge_lnl (Lk(b))"] = timeit(lambda: (eng.update_pmats(E, L), eng.edge_lnl(left, rght, e)))
res["1 update_partial + pmat + edge_lnl (one SPR candidate)"] = timeit(
    lambda: (eng.update_pmats(E, L), eng.update_partials(op), eng.edge_lnl(left, rght, e)))
res["pmat + fused (1 update + edge_lnl) (one SPR candidate as the binding issues it: 2 launches)"] = timeit(
    lambda: (eng.update_pmats(E, L), eng.traverse_edge_lnl(op, left, rght, e)))
eng.eigen_lr(left, rght)
res["eigen_lr (Update_Eigen_Lr)"] = timeit(lambda: eng.eigen_lr(left, rght))
res["lnl_dlnl (one dLk)"] = timeit(lambda: eng.lnl_dlnl(0.05))
res["Br_Len_Opt (Lk(b)+eigen_lr+~20 dLk+pmat)"] = timeit(lambda: t.Br_Len_Opt(e), n=30, warm=3)
full = pack_ops(tree.post_order_ops())
res["Lk(NULL) post-order"] = timeit(lambda: (eng.update_pmats(np.arange(tree.n_edges, dtype=np.int32), tree.l),
                                             eng.update_partials(full), eng.edge_lnl(*tree.edge_sides(tree.root_edge), tree.root_edge)), n=50)
# all regraft positions of one pruned subtree in ONE call (plk_spr_candidates) vs one call sequence per candidate
tip = 3
te = tree.adj[tip][0][0]
cands = []
for ee in range(tree.n_edges):
    if ee != te:
        a, b = tree.edge_sides(ee)
        cands.append((a, 0.5 * tree.l[ee], b, 0.5 * tree.l[ee]))
packed = eng.pack_spr_cands(cands)
t_batch = timeit(lambda: eng.spr_candidates(tree.side_of(te, tip), float(tree.l[te]), True, packed), n=20, warm=3)
res["plk_spr_candidates: %d candidates in one call, per candidate" % len(cands)] = t_batch / len(cands)
for k, v in res.items():
    print(f"{v:10.1f} us  {k}")
