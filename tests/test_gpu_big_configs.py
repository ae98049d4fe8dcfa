"""BASELINE.json configurations at their stated shapes against the reference's OWN numbers on the SAME
alignment, tree and model (tests/golden/big/*.npz, produced by tests/golden/make_golden_big.py from the
unmodified reference build): lnL to <= 1e-9 relative (north_star), per-pattern lnL and scaler sums on a
subset of patterns.  The alignments are re-simulated here from the seeded generators of
phyml_b200/workloads.py (the pattern weights are checked against the reference's block by block).
Through the C ABI; needs a B200: marked gpu."""
import numpy as np
import pytest

from phyml_b200 import workloads as wl
from phyml_b200.engine import Engine
from phyml_b200.lk import LkTree

pytestmark = pytest.mark.gpu

RTOL = 1e-9   # BASELINE.json north_star: "lnL matches the reference's own AVX build ... to <= 1e-9 relative"


def evaluate(name, blocks=None, devices=None):
    w = wl.WORKLOADS[name]
    m, pin = wl.evaluation_model(name)
    assert pin is not None, f"no reference pin for {name}: run tests/golden/make_golden_big.py"
    tree = wl.make_tree(w)
    blocks = list(range(w.n_blocks)) if blocks is None else blocks
    pat = wl.make_patterns(name, blocks, procs=min(8, len(blocks)))
    eng = Engine(tree.n_otu, pat.n_pattern, w.ns, m.ncatg, tree.n_clv_handles, tree.n_edges, devices=devices)
    t = LkTree(tree, pat, m, eng)
    lnl = t.Lk()
    return w, pin, pat, t, lnl


def check_sites(pin, pat, eng, blocks):
    """per-pattern lnL / fact_sum_scale of the pinned subset that falls into `blocks` (contiguous from block 0)."""
    n_pat = pin["n_pattern_blocks"]
    assert pat.n_pattern == int(n_pat[blocks].sum())
    assert abs(pat.wght.sum() - pin["wght_sum_blocks"][blocks].sum()) < 1e-6
    s = eng.get_site_lnl()
    idx = pin["sub_idx"]
    keep = idx < pat.n_pattern
    np.testing.assert_allclose(s["site_lnl"][idx[keep]], pin["sub_site_lnl"][keep], rtol=1e-10, atol=0)
    assert (s["fact_sum_scale"][idx[keep]] == pin["sub_fact_sum_scale"][keep]).all()


@pytest.mark.parametrize("name", ["dna_100x100k", "dna_100x50k", "dna_500x62k", "aa_200x50k"])
def test_config_matches_reference_on_the_same_alignment(name):
    """configs[1] (DNA 100 x 100k), the configs[4] alignment (100 x 50k), a 1/16 column block of configs[3]
    (500 taxa, rescaling fires: fact_sum_scale up to 768) and configs[2] (AA 200 x 50k, where the 2^256
    rescaling of k_traverse_aa fires as well)."""
    w, pin, pat, t, lnl = evaluate(name)
    ref = float(pin["lnL"])
    assert abs(lnl - ref) <= RTOL * abs(ref), (name, lnl, ref)
    check_sites(pin, pat, t.eng, list(range(w.n_blocks)))
    if name in ("dna_500x62k", "aa_200x50k"):
        assert int(pin["sub_fact_sum_scale"].max()) > 0   # the rescaling branch is exercised


def test_config4_blocks_add_up():
    """configs[3] is evaluated by the reference in 16 column blocks (it cannot allocate the whole alignment,
    phyml_b200/workloads.py): the first four blocks on one GPU must give the sum of the reference's four
    block values -- the property the site-sharded 8-GPU run relies on."""
    blocks = [0, 1, 2, 3]
    w, pin, pat, t, lnl = evaluate("dna_500x1M", blocks)
    ref = float(pin["lnL_blocks"][blocks].sum())
    assert abs(lnl - ref) <= RTOL * abs(ref), (lnl, ref)
    check_sites(pin, pat, t.eng, blocks)


def test_sharded_instance_on_one_device_matches_reference():
    """plk_create_sharded with the same device listed three times: every entry point fans out over three
    shards (uploads, K0-K4, read-backs) and the all-shard sum must be the reference's value."""
    w, pin, pat, t, lnl = evaluate("dna_100x50k", devices=[0, 0, 0])
    assert t.eng.n_shards == 3
    ref = float(pin["lnL"])
    assert abs(lnl - ref) <= RTOL * abs(ref), (lnl, ref)
    check_sites(pin, pat, t.eng, [0])
    # the eigen-basis path (K3 + K4) through the shards against a single-device instance
    w2, _, pat2, t1, lnl1 = evaluate("dna_100x50k")
    assert abs(lnl1 - lnl) <= 1e-12 * abs(lnl)
    e = 11
    for x in (t, t1):
        x.Set_Both_Sides(1)
        x.Lk()
        x.Set_Update_Eigen_Lr(1)
        x.Lk(e)
        x.Set_Update_Eigen_Lr(0)
    for l in (0.01, 0.2):
        a, b = t.dLk(l, e), t1.dLk(l, e)
        assert a[0] == b[0] and abs(a[1] - b[1]) <= 1e-12 * abs(b[1])
        assert abs(t.c_dlnL - t1.c_dlnL) <= 1e-9 * max(1.0, abs(t1.c_dlnL))
    h = t.tree.post_order_ops()[10].dst
    clv_a, sc_a = t.eng.get_clv(h)
    clv_b, sc_b = t1.eng.get_clv(h)
    assert np.array_equal(clv_a, clv_b) and np.array_equal(sc_a, sc_b)
