#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py

For every data set it runs oracle/_ref/ref_driver (the reference's own Lk/dLk/Update_Partial_Lk
linked as a library, see oracle/ref_driver.c), parses the binary dump and stores
  * all inputs of the hot path (tip vectors as 0/1 bit masks, d_state, is_ambigu, weights, invar,
    topology, branch lengths, eigen system, rates),
  * all outputs that are small (lnL, per-site lnL, per-category site likelihoods, every scaler,
    every P-matrix, lnL at every edge, lnL/dlnL probes),
  * CLVs and dot_prod restricted to a subset of site patterns (sites are independent, so the
    subset is a coherent test case), plus the sum of every full CLV as a weak all-site check.
The synthetic inputs are produced by phyml_b200's seeded generators and written next to the
fixtures so the run is reproducible.  The reference cannot travel to the GPU box; these files do.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from phyml_b200 import alignment, model as pmodel  # noqa: E402
from phyml_b200.tree import Tree  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
REF_EXAMPLES = "/root/reference/examples"
_DT = {"d": np.float64, "i": np.int32, "h": np.int16, "B": np.uint8}


def parse_dump(path):
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0
    while pos < len(buf):
        name = buf[pos:pos + 48].split(b"\0", 1)[0].decode()
        dt = chr(buf[pos + 48])
        n = int(np.frombuffer(buf, dtype=np.int64, count=1, offset=pos + 56)[0])
        pos += 64
        arr = np.frombuffer(buf, dtype=_DT[dt], count=n, offset=pos).copy()
        pos += n * arr.itemsize
        out[name] = arr
    return out


def run_driver(workdir, phyml_args, n_dlk=4):
    dump = os.path.join(workdir, "dump.bin")
    cmd = [DRIVER, "--dump", dump, "--dlk", str(n_dlk), "--"] + phyml_args
    res = subprocess.run(cmd, cwd=workdir, capture_output=True, text=True)
    if res.returncode != 0 or not os.path.exists(dump):
        sys.stderr.write(res.stdout[-2000:] + res.stderr[-2000:])
        raise RuntimeError("ref_driver failed")
    return parse_dump(dump)


def to_fixture(d, n_sub, extra_sites=()):
    n_otu, P, ns, ncatg = (int(d[k][0]) for k in ("n_otu", "n_pattern", "ns", "ncatg"))
    n_edges = 2 * n_otu - 3
    fx = {}
    for k in ("n_otu", "n_pattern", "ns", "ncatg", "tip_root", "root_edge", "invar_flag", "pinvar",
              "l_min", "l_max", "br_len_mult", "scaling_method", "apply_lk_scaling", "alpha", "lnL",
              "n_dlk"):
        fx[k] = d[k][0]
    for k in ("wght", "invar", "U", "V", "lambda", "pi", "rates", "rate_probs", "site_lnl",
              "site_lk", "site_lk_cat", "fact_sum_scale", "edge_lnl"):
        fx[k] = d[k]
    if "qmat" in d:
        fx["qmat"] = d["qmat"]
    # tips: 0/1 vectors -> bit masks; keep the reference's d_state / is_ambigu verbatim
    masks = np.zeros((n_otu, P), dtype=np.uint32)
    for i in range(n_otu):
        v = d[f"tip{i}.vec"].reshape(P, ns)
        assert np.isin(v, (0.0, 1.0)).all()
        masks[i] = (v.astype(np.uint32) << np.arange(ns, dtype=np.uint32)[None, :]).sum(axis=1)
    fx["tip_mask"] = masks
    fx["tip_d_state"] = np.stack([d[f"tip{i}.d_state"] for i in range(n_otu)])
    fx["tip_is_ambigu"] = np.stack([d[f"tip{i}.is_ambigu"] for i in range(n_otu)])
    fx["tip_names"] = np.array([bytes(d[f"tip{i}.name"]).decode() for i in range(n_otu)])
    fx["edge_nodes"] = np.stack([d[f"edge{e}.nodes"] for e in range(n_edges)])
    fx["edge_l"] = np.array([d[f"edge{e}.l"][0] for e in range(n_edges)])
    fx["edge_P"] = np.stack([d[f"edge{e}.P"].reshape(ncatg, ns, ns) for e in range(n_edges)])

    # site subset: evenly spread + the sites with the largest scalers + requested extras
    scal_tot = np.zeros(P, dtype=np.int64)
    for e in range(n_edges):
        for side in ("left", "rght"):
            k = f"edge{e}.scale_{side}"
            if k in d:
                scal_tot = np.maximum(scal_tot, d[k])
    sub = set(np.linspace(0, P - 1, min(P, n_sub)).astype(int).tolist())
    sub |= set(np.argsort(-scal_tot, kind="stable")[:4].tolist())
    sub |= set(int(s) for s in extra_sites)
    sub = np.array(sorted(sub), dtype=np.int64)
    fx["sites_sub"] = sub

    scales = np.full((2 * n_edges, P), -1, dtype=np.int32)   # handle = 2*edge + side; -1: tip side
    clv_sub = np.zeros((2 * n_edges, len(sub), ncatg, ns))
    clv_sum = np.zeros(2 * n_edges)
    has_clv = np.zeros(2 * n_edges, dtype=np.uint8)
    for e in range(n_edges):
        for sidx, side in enumerate(("left", "rght")):
            k = f"edge{e}.clv_{side}"
            if k in d:
                clv = d[k].reshape(P, ncatg, ns)
                h = 2 * e + sidx
                clv_sub[h] = clv[sub]
                clv_sum[h] = clv.sum()
                scales[h] = d[f"edge{e}.scale_{side}"]
                has_clv[h] = 1
    fx["clv_sub"], fx["clv_sum"], fx["scales"], fx["has_clv"] = clv_sub, clv_sum, scales, has_clv

    n_dlk = int(d["n_dlk"][0])
    fx["dlk_edge"] = np.array([d[f"dlk{k}.edge"][0] for k in range(n_dlk)], dtype=np.int32)
    fx["dlk_dot_prod_sub"] = np.stack(
        [d[f"dlk{k}.dot_prod"].reshape(P, ncatg, ns)[sub] for k in range(n_dlk)])
    fx["dlk_probes"] = np.stack([d[f"dlk{k}.probes"].reshape(5, 4) for k in range(n_dlk)])
    return fx


def synth(workdir, name, n_taxa, n_sites, m, seed, ambiguity, mean_bl):
    tree = Tree.random(n_taxa, seed=seed, mean_bl=mean_bl)
    codes = alignment.simulate(tree, m, n_sites, seed=seed + 100, ambiguity=ambiguity)
    phy = os.path.join(HERE, f"{name}.phy")
    nwk = os.path.join(HERE, f"{name}.nwk")
    alignment.write_phylip(phy, codes, m.ns, tree.names)
    with open(nwk, "w") as f:
        f.write(tree.to_newick() + "\n")
    for p in (phy, nwk):
        subprocess.run(["cp", p, workdir], check=True)
    return os.path.basename(phy), os.path.basename(nwk)


def main():
    if not os.path.exists(DRIVER):
        raise SystemExit("build the reference first: make -C oracle ref")
    scalars = {}
    with tempfile.TemporaryDirectory() as wd:
        common = ["-b", "0", "--r_seed", "1", "--no_memory_check", "-o", "n"]
        for ex in ("nucleic", "proteic"):
            subprocess.run(["cp", os.path.join(REF_EXAMPLES, ex), wd], check=True)

        jobs = []
        # config 1 of BASELINE.json: examples/nucleic, HKY85 + Gamma4, BioNJ tree
        jobs.append(("nucleic_hky", ["-i", "nucleic", "-d", "nt", "-m", "HKY85", "-c", "4", "-a", "1.0",
                                     "-t", "4.0", "-f", "e"], 16))
        # GTR + Gamma4 + I (exercises Invariant_Lk)
        jobs.append(("nucleic_gtr_inv", ["-i", "nucleic", "-d", "nt", "-m", "GTR", "-c", "4", "-a", "0.5",
                                         "-v", "0.2", "-f", "e"], 12))
        jobs.append(("proteic_lg", ["-i", "proteic", "-d", "aa", "-m", "LG", "-c", "4", "-a", "1.0",
                                    "-f", "m"], 12))
        # deep tree, long branches: triggers the 2^256 rescaling many times; 4% ambiguity
        phy, nwk = synth(wd, "synth_dna_deep", 200, 150, pmodel.gtr(alpha=0.5), 11, 0.04, 0.35)
        jobs.append(("synth_dna_deep", ["-i", phy, "-u", nwk, "-d", "nt", "-m", "GTR", "-c", "4", "-a",
                                        "0.5", "-f", "e"], 12))
        phy, nwk = synth(wd, "synth_aa_small", 24, 200, pmodel.synthetic_aa(), 5, 0.05, 0.2)
        jobs.append(("synth_aa_small", ["-i", phy, "-u", nwk, "-d", "aa", "-m", "LG", "-c", "4", "-a",
                                        "0.7", "-f", "m"], 12))
        # single rate category, no gamma (ncatg = 1 code path)
        jobs.append(("nucleic_jc_c1", ["-i", "nucleic", "-d", "nt", "-m", "JC69", "-c", "1"], 8))

        for name, args, n_sub in jobs:
            d = run_driver(wd, args + common)
            fx = to_fixture(d, n_sub)
            np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **fx)
            scalars[name] = {"lnL": float(d["lnL"][0]), "args": " ".join(args + common),
                             "n_otu": int(d["n_otu"][0]), "n_pattern": int(d["n_pattern"][0])}
            print(f"{name}: lnL={d['lnL'][0]:.15g} P={int(d['n_pattern'][0])} "
                  f"max_scale={int(fx['scales'].max())}")
            if name == "proteic_lg":
                np.savez_compressed(os.path.join(HERE, "lg_model.npz"), U=d["U"], V=d["V"],
                                    **{"lambda": d["lambda"]}, pi=d["pi"], qmat=d.get("qmat", np.zeros(0)))

        # the plain GTR + Gamma4 value quoted in SURVEY.md section 8(c) / BASELINE.md
        d = run_driver(wd, ["-i", "nucleic", "-d", "nt", "-m", "GTR", "-c", "4", "-a", "0.5", "-f", "e"] + common, 0)
        scalars["nucleic_gtr"] = {"lnL": float(d["lnL"][0])}

    with open(os.path.join(HERE, "lnl_values.json"), "w") as f:
        json.dump(scalars, f, indent=1, sort_keys=True)
    print(json.dumps({k: v["lnL"] for k, v in scalars.items()}, indent=1))


if __name__ == "__main__":
    main()
