#!/usr/bin/env python
"""Generate the parsimony goldens tests/golden/pars/<case>.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden_pars.py

For every data set of tests/golden/lnl_values.json (same inputs and phyml arguments as the likelihood
fixtures, so tree, weights and tip masks are those of <case>.npz) it runs
`oracle/_ref/ref_driver --dump_pars`, i.e. the reference's own Pars / Update_Partial_Pars / Pars_Core
(src/pars.c:20,239,397) with both_sides == YES, and stores

  ui, pars       [2*n_edges][P]      Fitch state sets and step counts, handle = 2*edge + (0: left, 1: rght)
  site_pars, c_pars, edge_pars       per-pattern steps at the root edge, weighted total, Pars(b) at every edge
  p_pars         [2*n_edges][P][ns]  the general (step-matrix, `general_pars`) variant, + its site_pars/c_pars/edge_pars
  step_mat       [ns][ns]            Get_Step_Mat (src/pars.c:498)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_golden import DRIVER, REF_EXAMPLES, parse_dump  # noqa: E402


def main():
    if not os.path.exists(DRIVER):
        raise SystemExit("build the reference first: make -C oracle ref")
    cases = json.load(open(os.path.join(HERE, "lnl_values.json")))
    out_dir = os.path.join(HERE, "pars")
    os.makedirs(out_dir, exist_ok=True)
    summary = {}
    with tempfile.TemporaryDirectory() as wd:
        for ex in ("nucleic", "proteic"):
            subprocess.run(["cp", os.path.join(REF_EXAMPLES, ex), wd], check=True)
        for f in os.listdir(HERE):
            if f.endswith((".phy", ".nwk")):
                subprocess.run(["cp", os.path.join(HERE, f), wd], check=True)
        for name, info in sorted(cases.items()):
            if "args" not in info or name in ("nucleic_gtr_inv", "nucleic_jc_c1"):  # same alignment as nucleic_hky
                continue
            dump = os.path.join(wd, "pars.bin")
            res = subprocess.run([DRIVER, "--dump_pars", dump, "--"] + info["args"].split(), cwd=wd,
                                 capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout[-2000:] + res.stderr[-2000:])
                raise RuntimeError("ref_driver failed")
            d = parse_dump(dump)
            n_otu, P, ns = (int(d[k][0]) for k in ("n_otu", "n_pattern", "ns"))
            n_edges = 2 * n_otu - 3
            main_fx = np.load(os.path.join(HERE, name + ".npz"))
            nodes = np.stack([d[f"edge{e}.nodes"] for e in range(n_edges)])
            assert (nodes == main_fx["edge_nodes"]).all() and (d["wght"] == main_fx["wght"]).all()
            fx = {"n_otu": n_otu, "n_pattern": P, "ns": ns, "step_mat": d["step_mat"].reshape(ns, ns)}
            fx["ui"] = np.stack([d[f"edge{e}.ui_{s}"] for e in range(n_edges) for s in "lr"])
            fx["pars"] = np.stack([d[f"edge{e}.pars_{s}"] for e in range(n_edges) for s in "lr"])
            fx["p_pars"] = np.stack([d[f"edge{e}.p_pars_{s}"].reshape(P, ns) for e in range(n_edges) for s in "lr"])
            for k in ("site_pars", "edge_pars", "site_pars_general", "edge_pars_general"):
                fx[k] = d[k]
            fx["c_pars"], fx["c_pars_general"] = int(d["c_pars"][0]), int(d["c_pars_general"][0])
            # the tips' Fitch sets are the bit masks of the likelihood fixture (Init_Ui_Tips, src/pars.c:164)
            for i in range(n_otu):
                e = int(np.where(nodes[:, 1] == i)[0][0])
                assert (fx["ui"][2 * e + 1] == main_fx["tip_mask"][i].astype(np.int32)).all()
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), **fx)
            summary[name] = {"c_pars": fx["c_pars"], "c_pars_general": fx["c_pars_general"]}
            print(name, summary[name], "P =", P)
    with open(os.path.join(out_dir, "pars_values.json"), "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
