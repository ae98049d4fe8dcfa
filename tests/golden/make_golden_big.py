#!/usr/bin/env python
"""Pin the BASELINE.json configurations to the UNMODIFIED reference on the SAME alignment and tree.

Run in the build container (needs /root/reference and ~30 GB of RAM):
    make -C oracle ref && python tests/golden/make_golden_big.py [workload ...]

For every pinned workload of phyml_b200/workloads.py and every column block of it, the block is
written as a PHYLIP file next to the workload's Newick tree and evaluated by oracle/_ref/ref_driver
(the reference's own Lk(NULL), AVX2+FMA build, fixed tree `-o n`).  Stored in
tests/golden/big/<workload>.npz:
  * the reference's eigen system / rates for that CLI (the evaluation model of bench.py and of the tests),
  * lnL of every block and their sum in block order (the value an N-GPU site-sharded run must reproduce),
  * the number of patterns and the total weight of every block (the pattern order of
    phyml_b200.alignment.compress is asserted identical to Compact_Data's: weights match element-wise),
  * per-pattern lnL and fact_sum_scale on a subset of patterns (evenly spread + the most rescaled ones),
    as indices into the concatenated pattern list.
The reference cannot travel to the GPU box; these small files do.
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import DRIVER, parse_dump  # noqa: E402
from phyml_b200 import alignment, workloads as wl  # noqa: E402

PINNED = ["dna_100x100k", "dna_100x50k", "dna_500x62k", "aa_200x50k", "dna_500x1M"]
N_SUB = 192


def pin(name):
    w = wl.WORKLOADS[name]
    tree = wl.make_tree(w)
    out = {}
    lnl_blocks, n_pat, w_sum, sub_idx, sub_lnl, sub_fact, secs = [], [], [], [], [], [], []
    offset = 0
    with tempfile.TemporaryDirectory() as wd:
        with open(os.path.join(wd, "tree.nwk"), "w") as f:
            f.write(tree.to_newick() + "\n")
        for b in range(w.n_blocks):
            codes = wl.block_codes(w, b)
            pat = alignment.compress(codes, w.ns)
            phy = os.path.join(wd, f"block{b}.phy")
            alignment.write_phylip(phy, codes, w.ns, tree.names)
            summ = os.path.join(wd, "summary.bin")
            cmd = [DRIVER, "--summary", summ, "--", "-i", phy, "-u", os.path.join(wd, "tree.nwk")] + wl.REF_ARGS[w.ns] + \
                  ["-o", "n", "-b", "0", "--r_seed", "1", "--no_memory_check"]
            t0 = time.time()
            res = subprocess.run(cmd, cwd=wd, capture_output=True, text=True)
            secs.append(time.time() - t0)
            if res.returncode != 0 or not os.path.exists(summ):
                sys.stderr.write(res.stdout[-2000:] + res.stderr[-2000:])
                raise RuntimeError("ref_driver failed")
            d = parse_dump(summ)
            os.remove(summ)
            os.remove(phy)
            P = int(d["n_pattern"][0])
            # same patterns in the same order as the engine's host-side compression
            assert P == pat.n_pattern, (P, pat.n_pattern)
            assert np.array_equal(d["wght"], pat.wght)
            assert np.array_equal(d["invar"], pat.invar)
            if b == 0:
                for k in ("U", "V", "lambda", "pi", "rates", "rate_probs"):
                    out[k] = d[k]
                for k in ("pinvar", "invar_flag", "l_min", "l_max", "br_len_mult", "alpha"):
                    out[k] = d[k][0]
            else:
                for k in ("U", "V", "lambda", "pi", "rates", "rate_probs"):
                    assert np.array_equal(out[k], d[k]), f"model differs between blocks ({k})"
            lnl_blocks.append(float(d["lnL"][0]))
            n_pat.append(P)
            w_sum.append(float(d["wght"].sum()))
            idx = set(np.linspace(0, P - 1, min(P, N_SUB)).astype(int).tolist())
            idx |= set(np.argsort(-d["fact_sum_scale"], kind="stable")[:8].tolist())
            idx = np.array(sorted(idx), dtype=np.int64)
            sub_idx.append(idx + offset)
            sub_lnl.append(d["site_lnl"][idx])
            sub_fact.append(d["fact_sum_scale"][idx])
            offset += P
            print(f"{name} block {b}: P={P} lnL={lnl_blocks[-1]:.17g} max_fact={int(d['fact_sum_scale'].max())} "
                  f"({secs[-1]:.0f} s)", flush=True)
    total = 0.0
    for v in lnl_blocks:          # block order, the order in which the ranks' partial sums are added
        total += v
    out.update(lnL=total, lnL_blocks=np.array(lnl_blocks), n_pattern_blocks=np.array(n_pat, dtype=np.int64),
               wght_sum_blocks=np.array(w_sum), sub_idx=np.concatenate(sub_idx), sub_site_lnl=np.concatenate(sub_lnl),
               sub_fact_sum_scale=np.concatenate(sub_fact), n_otu=w.n_taxa, ns=w.ns,
               ref_args=np.array(" ".join(wl.REF_ARGS[w.ns])))
    os.makedirs(wl.BIG_GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(wl.golden_path(name), **out)
    return {"lnL": total, "lnL_blocks": lnl_blocks, "n_pattern_blocks": n_pat, "ref_args": " ".join(wl.REF_ARGS[w.ns]),
            "ref_seconds": [round(s, 1) for s in secs]}


def main():
    if not os.path.exists(DRIVER):
        raise SystemExit("build the reference first: make -C oracle ref")
    names = sys.argv[1:] or PINNED
    jpath = os.path.join(wl.BIG_GOLDEN_DIR, "lnl_values.json")
    os.makedirs(wl.BIG_GOLDEN_DIR, exist_ok=True)
    vals = json.load(open(jpath)) if os.path.exists(jpath) else {}
    for nm in names:
        vals[nm] = pin(nm)
        with open(jpath, "w") as f:
            json.dump(vals, f, indent=1, sort_keys=True)
    print(json.dumps({k: v["lnL"] for k, v in vals.items()}, indent=1))


if __name__ == "__main__":
    main()
