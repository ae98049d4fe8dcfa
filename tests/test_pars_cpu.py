"""Pins the oracle's parsimony restatement (oracle/plk_oracle.c: plk_oracle_pars_*) to the UNMODIFIED reference's
Pars / Update_Partial_Pars / Pars_Core (src/pars.c:20,239,397) through the dumps in tests/golden/pars.  CPU only."""
import json
import os

import numpy as np
import pytest

import pars_checks as pk
from oracle_backend import GOLDEN_DIR, OracleBackend


def make(case):
    c, g = pk.load(case)
    return c, g, OracleBackend(c.n_otu, c.P, c.ns, c.ncatg, c.tree.n_clv_handles, c.tree.n_edges)


@pytest.mark.parametrize("general", [False, True])
@pytest.mark.parametrize("case", pk.PARS_CASES)
def test_full_traversal(case, general):
    pk.check_full(*make(case), general, fused=False)


@pytest.mark.parametrize("general", [False, True])
@pytest.mark.parametrize("case", ["nucleic_hky", "synth_aa_small"])
def test_single_updates(case, general):
    pk.check_single_updates(*make(case), general)


def test_reference_totals():
    vals = json.load(open(os.path.join(GOLDEN_DIR, "pars", "pars_values.json")))
    for case in pk.PARS_CASES:
        c, g = pk.load(case)
        assert int(g["c_pars"]) == vals[case]["c_pars"] and int(g["c_pars_general"]) == vals[case]["c_pars_general"]
        # Fitch never exceeds the step-matrix score when every step costs at least 1
        assert (g["site_pars"] <= g["site_pars_general"]).all()


def test_truncating_accumulation():
    """c_pars is an int accumulated with `+= site_pars * wght` (src/pars.c:46): fractional weights truncate per site"""
    tree, ui, w, step = pk.random_case(9, 257, 4, seed=3, frac_weights=True)
    eng = OracleBackend(tree.n_otu, 257, 4, 1, tree.n_clv_handles, tree.n_edges)
    c_pars, site, _ = pk.run_random(tree, ui, w, step, eng, False)
    c = 0
    for s in range(257):
        c = int(float(c) + float(site[s]) * w[s])
    assert c_pars == c and c_pars != int(np.dot(site, w))
