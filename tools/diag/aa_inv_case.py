"""Diagnostic for tests/test_gpu_parity.py::test_engine_vs_oracle_synthetic[20-14-500-4-0.15-...]: where do the CUDA and
oracle CLVs differ by more than 1e-12 of the site maximum?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as t
from oracle_backend import OracleBackend
from phyml_b200.engine import Engine
from phyml_b200.lk import LkTree

ns, n_taxa, n_sites, ncatg, pinv, amb, bl = 20, 14, 500, 4, 0.15, 0.02, 0.1
tree, m, pat = t._synthetic(ns, n_taxa, n_sites, 3, amb, ncatg, pinv, bl)
args = (tree.n_otu, pat.n_pattern, ns, ncatg, tree.n_clv_handles, tree.n_edges)
gpu = LkTree(tree, pat, m, Engine(*args)); cpu = LkTree(tree, pat, m, OracleBackend(*args))
for x in (gpu, cpu): x.Set_Both_Sides(1)
print("lnL", gpu.Lk(), cpu.Lk(), "P", pat.n_pattern, "zero-weight patterns", int((pat.wght <= 0).sum()))
ops = {o.dst: o for o in tree.post_order_ops() + tree.pre_order_ops()}
for h in range(tree.n_clv_handles):
    if h not in cpu.eng.clv: continue
    a, sa = gpu.eng.get_clv(h); b, sb = cpu.eng.get_clv(h)
    lim = 1e-12 * np.abs(b).max(axis=(1, 2), keepdims=True)
    bad = np.argwhere(np.abs(a - b) > lim)
    if len(bad):
        o = ops[h]
        s, c, i = bad[0]
        print(f"handle {h} ({'post' if h in [q.dst for q in tree.post_order_ops()] else 'pre'}) c1={o.c1} c2={o.c2}: {len(bad)} bad entries, "
              f"first site {s} cat {c} state {i}: gpu {a[s,c,i]!r} cpu {b[s,c,i]!r} site max {np.abs(b[s]).max()!r} "
              f"wght {pat.wght[s]} scale {sa[s]} {sb[s]} codes {[int(pat.codes[k][s]) for k in range(tree.n_otu)]}")
        print("   gpu row", a[s, c]); print("   cpu row", b[s, c])
