import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from phyml_b200 import workloads as wl
from phyml_b200.engine import Engine, pack_ops
name = "dna_100x100k"
w = wl.WORKLOADS[name]
m, _pin = wl.evaluation_model(name)
tree = wl.make_tree(w)
pat = wl.make_patterns(name, [0], procs=8)
eng = Engine(tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges)
eng.set_weights(pat.wght, pat.invar); eng.set_tip_table(pat.table()); eng.set_all_tip_codes(pat.codes); eng.set_model(m)
ops = pack_ops(tree.post_order_ops()); edges = np.arange(tree.n_edges, dtype=np.int32); L = tree.l.copy()
left, rght = tree.edge_sides(tree.root_edge)
for _ in range(5):
    eng.update_pmats(edges, L); eng.update_partials(ops); eng.edge_lnl(left, rght, tree.root_edge)
N=50; t=[0,0,0,0]
T0=time.perf_counter()
for _ in range(N):
    a=time.perf_counter(); eng.update_pmats(edges, L); b=time.perf_counter(); eng.update_partials(ops); c=time.perf_counter(); eng.edge_lnl(left, rght, tree.root_edge); d=time.perf_counter()
    t[0]+=b-a; t[1]+=c-b; t[2]+=d-c
tot=(time.perf_counter()-T0)/N
print("per eval us: total %.1f  update_pmats %.1f  update_partials %.1f  edge_lnl(wait) %.1f" % (tot*1e6, t[0]/N*1e6, t[1]/N*1e6, t[2]/N*1e6))
# back-to-back async enqueue of 20 evals without reading lnL (GPU-bound time)
eng.sync(); T0=time.perf_counter()
for _ in range(20):
    eng.update_pmats(edges, L); eng.update_partials(ops)
eng.sync(); print("K0+K1 pipelined per eval us: %.1f" % ((time.perf_counter()-T0)/20*1e6))
