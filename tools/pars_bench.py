"""Latency of the parsimony path (plk_pars_*) at the SPR configuration (BASELINE configs[4]: 100 taxa x 50 000 sites):
Pars(NULL) with both_sides (3n - 6 updates + the site loop, one launch) and one SPR candidate (1 update + Pars(b)),
against the oracle's scalar restatement of src/pars.c on one host core.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pars_checks as pk  # noqa: E402
from oracle_backend import OracleBackend  # noqa: E402
from phyml_b200.engine import Engine  # noqa: E402

n_otu, P, ns = (int(x) for x in (sys.argv[1:4] + [100, 50000, 4][len(sys.argv) - 1:]))
tree, ui, w, step = pk.random_case(n_otu, P, ns, seed=1)
args = (tree.n_otu, P, ns, 1, tree.n_clv_handles, tree.n_edges)
out = {"n_otu": n_otu, "n_pattern": P, "ns": ns}
for name, eng, reps in (("b200", Engine(*args), 200), ("oracle_1core", OracleBackend(*args), 3)):
    c_pars, site, _ = pk.run_random(tree, ui, w, step, eng, False)
    a, d = 0, tree.adj[0][0][1]
    e0 = tree.adj[0][0][0]
    ops = np.asarray(tree.pars_ops(tree.post_order_ops(a, d)) + tree.pars_ops(tree.pre_order_ops(a, d)), dtype=np.int32)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.pars_traverse_edge(ops, 2 * e0, 2 * e0 + 1)
    t_full = (time.perf_counter() - t0) / reps
    one = ops[-1:].copy()
    t0 = time.perf_counter()
    for _ in range(reps * 5):
        eng.pars_traverse_edge(one, int(one[0, 0]), int(one[0, 0]) ^ 1)
    t_one = (time.perf_counter() - t0) / (reps * 5)
    out[name] = {"c_pars": int(c_pars), "pars_null_both_sides_us": round(t_full * 1e6, 1), "n_updates": int(len(ops)),
                 "candidate_us": round(t_one * 1e6, 1),
                 "pattern_updates_per_s": float(f"{len(ops) * P / t_full:.4g}")}
assert out["b200"]["c_pars"] == out["oracle_1core"]["c_pars"]
print(json.dumps(out))
