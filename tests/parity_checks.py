"""Parity checks shared by the CPU (oracle vs reference goldens) and GPU (CUDA engine vs reference
goldens / vs oracle) test files.  `eng` is anything with the Engine interface."""
import numpy as np


def assert_rows_close(a, b, tol):
    """dot_prod entries are signed and cancel: compare on the scale of each (site, category) row."""
    scale = np.abs(b).max(axis=-1, keepdims=True)
    assert (np.abs(a - b) <= tol * scale + 1e-300).all()


def check_pmat(c, eng, atol=5e-15):
    """K0 vs b->Pij_rr of every edge."""
    eng.update_pmats(range(c.tree.n_edges), c.tree.l)
    got = np.stack([eng.get_pmat(e) for e in range(c.tree.n_edges)])
    np.testing.assert_allclose(got, c.g["edge_P"], rtol=0, atol=atol)
    np.testing.assert_allclose(got.sum(axis=3), 1.0, rtol=0, atol=1e-14)


def check_full_traversal(c, eng, clv_rtol, exact=False):
    """K1 over post- and pre-order from the tips alone, with the reference's P-matrices."""
    for e in range(c.tree.n_edges):
        eng.set_pmat(e, c.g["edge_P"][e])
    c.tree.both_sides = True
    ops = c.tree.full_traversal_ops()
    assert len(ops) == 3 * (c.n_otu - 2)
    eng.update_partials(ops)
    n_checked = 0
    for h in range(c.tree.n_clv_handles):
        if not c.g["has_clv"][h]:
            continue
        clv, scale = eng.get_clv(h)
        assert (scale == c.g["scales"][h]).all(), f"scaler mismatch on handle {h}"
        if exact:
            assert np.array_equal(clv[c.sub], c.g["clv_sub"][h]), f"CLV {h} not bit-identical"
        else:
            np.testing.assert_allclose(clv[c.sub], c.g["clv_sub"][h], rtol=clv_rtol, atol=0)
        np.testing.assert_allclose(clv.sum(), c.g["clv_sum"][h], rtol=1e-12)
        n_checked += 1
    assert n_checked == 3 * (c.n_otu - 2)


def check_lnl_end_to_end(c, eng, rtol=1e-12):
    """K0+K1+K2 from tips and branch lengths only."""
    eng.update_pmats(range(c.tree.n_edges), c.tree.l)
    c.tree.both_sides = False
    eng.update_partials(c.tree.post_order_ops())
    left, rght = c.tree.edge_sides(c.tree.root_edge)
    lnl = eng.edge_lnl(left, rght, c.tree.root_edge)
    g = c.g
    assert abs(lnl - float(g["lnL"])) <= rtol * abs(float(g["lnL"])), (lnl, float(g["lnL"]))
    s = eng.get_site_lnl()
    live = g["wght"] > 0
    np.testing.assert_allclose(s["site_lnl"][live], g["site_lnl"][live], rtol=max(rtol, 1e-12))
    np.testing.assert_allclose(s["site_lk"][live], g["site_lk"][live], rtol=1e-11)
    # per-category terms: tiny categories inherit the absolute (not relative) accuracy of the
    # near-zero P entries (cancellation in U diag(exp) V), so compare on the scale of the site
    ref_cat = g["site_lk_cat"].reshape(s["site_lk_cat"].shape)
    np.testing.assert_allclose(s["site_lk_cat"], ref_cat, rtol=1e-8)
    assert (np.abs(s["site_lk_cat"] - ref_cat) <= 1e-11 * ref_cat.max(axis=1, keepdims=True)).all()
    assert (s["fact_sum_scale"] == g["fact_sum_scale"]).all()
    return lnl


def check_lnl_every_edge(c, eng, rtol=1e-12):
    eng.update_pmats(range(c.tree.n_edges), c.tree.l)
    c.tree.both_sides = True
    eng.update_partials(c.tree.full_traversal_ops())
    vals = []
    for e in range(c.tree.n_edges):
        left, rght = c.tree.edge_sides(e)
        vals.append(eng.edge_lnl(left, rght, e))
    np.testing.assert_allclose(vals, c.g["edge_lnl"], rtol=rtol)


def check_eigen_lr_and_dlk(c, eng, row_tol=1e-12, golden_pmat=False):
    g = c.g
    if golden_pmat:
        # the reference's own P-matrices: CLVs are then bit-identical and dot_prod can be compared tightly
        for e in range(c.tree.n_edges):
            eng.set_pmat(e, g["edge_P"][e])
    else:
        eng.update_pmats(range(c.tree.n_edges), c.tree.l)
    c.tree.both_sides = True
    eng.update_partials(c.tree.full_traversal_ops())
    for k, e in enumerate(g["dlk_edge"]):
        left, rght = c.tree.edge_sides(int(e))
        eng.eigen_lr(left, rght)
        assert_rows_close(eng.get_dot_prod()[c.sub], g["dlk_dot_prod_sub"][k], row_tol)
        l0 = c.tree.l[int(e)]
        for j, mult in enumerate((0.1, 0.5, 1.0, 2.0, 10.0)):
            lc, lnl, dlnl = eng.lnl_dlnl(l0 * mult)
            ref_l, ref_lnl, ref_dlnl, ref_lnl_eig = g["dlk_probes"][k][j]
            assert lc == ref_l
            assert abs(lnl - ref_lnl) <= 1e-12 * abs(ref_lnl)
            assert abs(dlnl - ref_dlnl) <= 1e-9 * max(1.0, abs(ref_dlnl))
            assert abs(eng.lnl_eigen(lc) - ref_lnl_eig) <= 1e-12 * abs(ref_lnl_eig)
