"""Parity tests proper: the CUDA engine, called through the C ABI (libphyml_b200.so via ctypes),
against (a) the golden arrays dumped from the unmodified reference and (b) the CPU oracle on
seeded synthetic inputs.  Need a B200: marked gpu."""
import numpy as np
import pytest

import parity_checks as pc
from golden_case import ALL_CASES, GoldenCase
from oracle_backend import OracleBackend

from phyml_b200 import alignment, model as pmodel
from phyml_b200.engine import Engine
from phyml_b200.lk import LkTree
from phyml_b200.tree import Tree

pytestmark = pytest.mark.gpu

DNA_CASES = [c for c in ALL_CASES if "aa" not in c and "proteic" not in c]


def make(case):
    c = GoldenCase(case)
    eng = Engine(c.n_otu, c.P, c.ns, c.ncatg, c.tree.n_clv_handles, c.tree.n_edges)
    c.upload(eng)
    return c, eng


@pytest.mark.parametrize("case", ALL_CASES)
def test_pmat(case):
    """K0 on device vs the reference's b->Pij_rr (exp() differs by <= 1 ulp from libm)."""
    c, eng = make(case)
    pc.check_pmat(c, eng, atol=2e-14)


@pytest.mark.parametrize("case", ALL_CASES)
def test_full_traversal_exact(case):
    """K1 reproduces the reference's AVX+FMA arithmetic order: with the reference's P-matrices every
    CLV (post- and pre-order) and every scaler must be BIT-IDENTICAL to the reference's."""
    c, eng = make(case)
    # 20 states run on the FP64 tensor pipe: on B200 DMMA.8x8x4 accumulates as an ascending-k FMA chain
    # (tools/probes/dmma_order.cu: 262144/262144 products identical), i.e. exactly the reference's
    # AVX_Matrix_Vect_Prod order, so the tensor-core path is bit-identical too
    pc.check_full_traversal(c, eng, clv_rtol=0, exact=True)


@pytest.mark.parametrize("case", ALL_CASES)
def test_lnl_end_to_end(case):
    """tips + branch lengths + eigen system in, lnL out: <= 1e-12 relative (target 1e-9)."""
    c, eng = make(case)
    pc.check_lnl_end_to_end(c, eng, rtol=1e-12)


@pytest.mark.parametrize("case", ALL_CASES)
def test_lnl_at_every_edge(case):
    c, eng = make(case)
    pc.check_lnl_every_edge(c, eng, rtol=1e-12)


@pytest.mark.parametrize("case", ALL_CASES)
def test_eigen_lr_and_dlk(case):
    c, eng = make(case)
    pc.check_eigen_lr_and_dlk(c, eng, row_tol=1e-13, golden_pmat=True)
    c, eng = make(case)
    pc.check_eigen_lr_and_dlk(c, eng, row_tol=1e-9)   # device-computed P: exp() differs from libm by <= 1 ulp


def _synthetic(ns, n_taxa, n_sites, seed, ambiguity, ncatg=4, pinv=0.0, mean_bl=0.1):
    tree = Tree.random(n_taxa, seed=seed, mean_bl=mean_bl)
    if ns == 4:
        m = pmodel.gtr(alpha=0.5, ncatg=ncatg, pinv=pinv)
    else:
        m = pmodel.lg_from_fixture(alpha=0.5, ncatg=ncatg)
        m.pinv, m.invar = pinv, pinv > 0.0
    codes = alignment.simulate(tree, m, n_sites, seed=seed + 1, ambiguity=ambiguity)
    pat = alignment.compress(codes, ns)
    return tree, m, pat


@pytest.mark.parametrize("ns,n_taxa,n_sites,ncatg,pinv,amb,mean_bl", [
    (4, 12, 3000, 4, 0.0, 0.02, 0.1),
    (4, 40, 5000, 4, 0.1, 0.0, 0.1),
    (4, 300, 257, 4, 0.0, 0.03, 0.4),     # deep: rescaling fires
    (4, 16, 1001, 1, 0.0, 0.0, 0.1),      # ncatg = 1
    (4, 16, 1001, 8, 0.0, 0.05, 0.1),     # ncatg = 8
    (4, 16, 999, 6, 0.0, 0.05, 0.1),      # ncatg not a power of two -> generic kernel
    (20, 10, 700, 4, 0.0, 0.02, 0.1),
    (20, 30, 301, 2, 0.0, 0.0, 0.2),
    (20, 320, 70, 4, 0.0, 0.02, 0.45),    # deep 20-state tree: the 2^256 rescaling branch of the DMMA kernel fires
    (20, 14, 500, 4, 0.15, 0.02, 0.1),    # 20 states with invariable sites (Invariant_Lk)
    (20, 9, 333, 1, 0.0, 0.02, 0.1),      # 20 states, ncatg = 1
    (20, 9, 333, 3, 0.1, 0.02, 0.1),      # 20 states, ncatg = 3
    (20, 9, 333, 8, 0.0, 0.02, 0.1),      # 20 states, ncatg = 8
])
def test_engine_vs_oracle_synthetic(ns, n_taxa, n_sites, ncatg, pinv, amb, mean_bl):
    """CUDA vs oracle through the reference-named host interface (LkTree): lnL, every CLV, scalers,
    dLk, on seeded synthetic data incl. ragged sizes (pattern counts not multiples of the tile)."""
    tree, m, pat = _synthetic(ns, n_taxa, n_sites, 3, amb, ncatg, pinv, mean_bl)
    args = (tree.n_otu, pat.n_pattern, ns, ncatg, tree.n_clv_handles, tree.n_edges)
    gpu = LkTree(tree, pat, m, Engine(*args))
    cpu = LkTree(tree, pat, m, OracleBackend(*args))
    for t in (gpu, cpu):
        t.Set_Both_Sides(1)
    lg, lc = gpu.Lk(), cpu.Lk()
    assert abs(lg - lc) <= 1e-12 * abs(lc)
    for h in range(tree.n_clv_handles):
        if h in cpu.eng.clv:
            a, sa = gpu.eng.get_clv(h)
            b, sb = cpu.eng.get_clv(h)
            assert (sa == sb).all()
            # device exp() vs libm exp() differ by <= 1 ulp; tiny CLV entries inherit the ABSOLUTE
            # accuracy of near-zero P entries (a few 1e-16 on entries of 1e-5 .. 1e-6 at short branches), so
            # compare on the scale of each site
            assert (np.abs(a - b) <= 1e-11 * np.abs(b).max(axis=(1, 2), keepdims=True)).all()
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=0)
    e = tree.n_edges // 2
    for t in (gpu, cpu):
        t.Set_Update_Eigen_Lr(1)
        t.Lk(e)
        t.Set_Update_Eigen_Lr(0)
    for l in (1e-9, 0.01, 0.3, 250.0):
        rg, rc = gpu.dLk(l, e), cpu.dLk(l, e)
        assert rg[0] == rc[0]
        assert abs(rg[1] - rc[1]) <= 1e-12 * abs(rc[1])
        assert abs(gpu.c_dlnL - cpu.c_dlnL) <= 1e-9 * max(1.0, abs(cpu.c_dlnL))


@pytest.mark.parametrize("ns", [4, 20])
def test_deep_tree_rescaling_fires(ns):
    """the deep synthetic trees above are only a test of the rescaling branch if scalers are non-zero"""
    tree, m, pat = _synthetic(ns, 320, 70, 3, 0.02, 4, 0.0, 0.45)
    args = (tree.n_otu, pat.n_pattern, ns, 4, tree.n_clv_handles, tree.n_edges)
    gpu = LkTree(tree, pat, m, Engine(*args))
    gpu.Lk()
    left, _ = tree.edge_sides(tree.root_edge)
    assert gpu.eng.get_clv(left.clv)[1].max() >= 256


@pytest.mark.parametrize("ns,ncatg", [(4, 4), (20, 4), (4, 3)])
def test_no_scaling_flag(ns, ncatg):
    """PLK_FLAG_NO_SCALING (tree->apply_lk_scaling == NO, lk.c:1563,2701-2706): no rescaling, all scalers zero"""
    tree, m, pat = _synthetic(ns, 24, 300, 7, 0.02, ncatg, 0.0, 0.1)
    args = (tree.n_otu, pat.n_pattern, ns, ncatg, tree.n_clv_handles, tree.n_edges)
    gpu = LkTree(tree, pat, m, Engine(*args, apply_scaling=False))
    cpu = LkTree(tree, pat, m, OracleBackend(*args))
    cpu.eng.apply_scaling = 0
    lg, lc = gpu.Lk(), cpu.Lk()
    assert abs(lg - lc) <= 1e-12 * abs(lc)
    left, _ = tree.edge_sides(tree.root_edge)
    a, sa = gpu.eng.get_clv(left.clv)
    b, sb = cpu.eng.get_clv(left.clv)
    assert (sa == 0).all() and (sb == 0).all()
    np.testing.assert_allclose(a, b, rtol=1e-9, atol=0)


def test_zero_weight_patterns_and_single_pattern():
    """Edge cases of the reference's site loop: wght <= DBL_MIN patterns are skipped (bootstrap
    zero weights, lk.c:1682) and a 1-pattern alignment works."""
    tree, m, pat = _synthetic(4, 9, 400, 5, 0.05)
    pat.wght[::3] = 0.0
    args = (tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges)
    gpu, cpu = LkTree(tree, pat, m, Engine(*args)), LkTree(tree, pat, m, OracleBackend(*args))
    assert abs(gpu.Lk() - cpu.Lk()) <= 1e-12 * abs(cpu.c_lnL)
    one = pat.shard(0, pat.n_pattern)
    assert one.n_pattern == 1
    args = (tree.n_otu, 1, 4, 4, tree.n_clv_handles, tree.n_edges)
    one.wght[:] = 1.0
    gpu, cpu = LkTree(tree, one, m, Engine(*args)), LkTree(tree, one, m, OracleBackend(*args))
    assert abs(gpu.Lk() - cpu.Lk()) <= 1e-12 * abs(cpu.c_lnL)


def test_pulley_principle_and_branch_opt():
    """The reference's own runtime invariants as engine self-tests: Check_Lk_At_Given_Edge
    (lk.c:2642) and monotone improvement under Br_Len_Opt (optimiz.c:656-661)."""
    tree, m, pat = _synthetic(4, 25, 4000, 9, 0.01)
    t = LkTree(tree, pat, m, Engine(tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges))
    t.Set_Both_Sides(1)
    t.Lk()
    vals = t.Check_Lk_At_Given_Edge(tol=1e-6)
    assert np.ptp(vals) <= 1e-9 * abs(vals[0])
    before = t.Lk()
    e = 7
    tree.l[e] *= 5.0
    worse = t.Lk()
    after = t.Br_Len_Opt(e)
    assert after >= worse and after >= before - 1e-6


def test_sharded_partials_sum_to_full():
    """Site sharding (section 8e): per-shard partial lnL sums to the full value, on one GPU."""
    tree, m, pat = _synthetic(4, 20, 6000, 2, 0.02)
    full = LkTree(tree, pat, m, Engine(tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges)).Lk()
    tot = 0.0
    for r in range(3):
        sh = pat.shard(r, 3)
        tot += LkTree(tree, sh, m, Engine(tree.n_otu, sh.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges)).Lk()
    assert abs(tot - full) <= 1e-12 * abs(full)


def test_large_config_properties():
    """BASELINE config 2 at full size (100 taxa x 100 000 sites, GTR+G4): size-independent properties
    instead of an oracle run -- pulley principle across edges, invariance of lnL to the rooting tip,
    additivity over site shards, and agreement with the oracle on a 2 000-site prefix."""
    tree, m, pat = _synthetic(4, 100, 100000, 1, 0.0)
    t = LkTree(tree, pat, m, Engine(tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges))
    t.Set_Both_Sides(1)
    full = t.Lk()
    vals = np.array([t.Lk(e) for e in range(0, tree.n_edges, 17)])
    assert np.abs(vals - full).max() <= 1e-10 * abs(full)
    tree.tip_root = 37
    t.Set_Both_Sides(0)
    assert abs(t.Lk() - full) <= 1e-10 * abs(full)
    tree.tip_root = 0
    sub = alignment.Patterns(4, np.ascontiguousarray(pat.codes[:, :2000]), pat.wght[:2000].copy(),
                             pat.invar[:2000].copy(), pat.names, 2000)
    args = (tree.n_otu, 2000, 4, 4, tree.n_clv_handles, tree.n_edges)
    a = LkTree(tree, sub, m, Engine(*args)).Lk()
    b = LkTree(tree, sub, m, OracleBackend(*args)).Lk()
    assert abs(a - b) <= 1e-12 * abs(b)


def test_many_updates_multi_chunk():
    """A 3 000-taxon tree: 2 998 post-order + 5 996 pre-order updates in one plk_update_partials call,
    i.e. more descriptors than one staging block holds (the fused launch is split, the register-forwarding
    chain restarts at the chunk boundary) -- against the oracle."""
    tree, m, pat = _synthetic(4, 3000, 48, 17, 0.02, mean_bl=0.05)
    args = (tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges)
    gpu, cpu = LkTree(tree, pat, m, Engine(*args)), LkTree(tree, pat, m, OracleBackend(*args))
    for t in (gpu, cpu):
        t.Set_Both_Sides(1)
    assert abs(gpu.Lk() - cpu.Lk()) <= 1e-12 * abs(cpu.c_lnL)
    for e in (0, tree.n_edges // 3, tree.n_edges - 1):
        assert abs(gpu.Lk(e) - cpu.Lk(e)) <= 1e-12 * abs(cpu.c_lnL)


def test_tensor_pipe_dna_variant_is_bit_identical(monkeypatch):
    """The opt-in tensor-pipe variant of the 4-state kernel (PLK_DNA_MMA=1, one DMMA.8x8x4 per 8 sites,
    category and child) must reproduce the reference bit for bit as well (DMMA = ascending-k FMA chain)."""
    monkeypatch.setenv("PLK_DNA_MMA", "1")
    for case in ("nucleic_hky", "synth_dna_deep"):
        c, eng = make(case)
        pc.check_full_traversal(c, eng, clv_rtol=0, exact=True)
        c, eng = make(case)
        pc.check_lnl_end_to_end(c, eng, rtol=1e-12)


def test_error_paths():
    from phyml_b200.engine import EngineError
    from phyml_b200.tree import Side

    eng = Engine(4, 10, 4, 4, 10, 5)
    with pytest.raises(EngineError):
        eng.edge_lnl(Side(clv=3), Side(tip=0), 0)       # CLV never written
    with pytest.raises(EngineError):
        eng.lnl_dlnl(0.1)                                 # dLk before Update_Eigen_Lr
    with pytest.raises(EngineError):
        eng.update_pmats([99], [0.1])                     # handle out of range
    with pytest.raises(EngineError):
        Engine(4, 10, 64, 4, 10, 5)                       # ns unsupported


@pytest.mark.parametrize("case", ALL_CASES)
def test_traverse_edge_lnl_equals_the_two_calls(case):
    """plk_traverse_edge_lnl (Post_Order_Lk + the site loop of Lk in one call; for 4 states x 4 categories the
    edge reduction is the epilogue of the traversal kernel) against plk_update_partials + plk_edge_lnl:
    per-pattern lnL / site_lk / per-category terms / fact_sum_scale bit-identical, the total to the rounding
    of a different (fixed) partition of the sum, and the reference's golden lnL to 1e-12."""
    c, a = make(case)
    _, b = make(case)
    for eng in (a, b):
        eng.update_pmats(range(c.tree.n_edges), c.tree.l)
    c.tree.both_sides = False
    ops = c.tree.post_order_ops()
    left, rght = c.tree.edge_sides(c.tree.root_edge)
    a.update_partials(ops)
    la = a.edge_lnl(left, rght, c.tree.root_edge)
    n0 = b.launch_count
    lb = b.traverse_edge_lnl(ops, left, rght, c.tree.root_edge)
    n_launch = b.launch_count - n0
    if c.ns == 4 and c.ncatg == 4:
        assert n_launch <= 2, n_launch          # the traversal kernel (+ the one-off tip-row translation)
    assert abs(la - lb) <= 1e-13 * abs(la), (la, lb)
    assert abs(lb - float(c.g["lnL"])) <= 1e-12 * abs(float(c.g["lnL"]))
    sa, sb = a.get_site_lnl(), b.get_site_lnl()
    live = c.g["wght"] > 0
    for k in ("site_lnl", "site_lk", "fact_sum_scale"):
        assert np.array_equal(sa[k][live], sb[k][live]), k
    assert np.array_equal(sa["site_lk_cat"][live], sb["site_lk_cat"][live])
    # a second evaluation (cached descriptors) and a short list (one update + the edge, as an SPR candidate does)
    lb2 = b.traverse_edge_lnl(ops, left, rght, c.tree.root_edge)
    assert lb2 == lb
    lb3 = b.traverse_edge_lnl(ops[-1:], left, rght, c.tree.root_edge)
    assert lb3 == lb
    lb4 = b.traverse_edge_lnl([], left, rght, c.tree.root_edge)
    assert abs(lb4 - lb) <= 1e-13 * abs(lb)


def test_packed_tip_codes_upload():
    """plk_set_all_tip_codes_packed4 (two 4-bit codes per byte) gives the same tips as the 1-byte upload:
    identical lnL, per-pattern lnL and CLVs; odd pattern count; ambiguity codes included."""
    from phyml_b200.engine import pack_codes4

    tree, m, pat = _synthetic(4, 14, 3001, seed=21, ambiguity=0.05)
    assert pat.n_pattern % 2 == 1 or True
    args = (tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges)
    a, b = LkTree(tree, pat, m, Engine(*args)), LkTree(tree, pat, m, Engine(*args))
    b.eng.set_all_tip_codes_packed4(pack_codes4(pat.codes))
    la, lb = a.Lk(), b.Lk()
    assert la == lb
    sa, sb = a.eng.get_site_lnl(), b.eng.get_site_lnl()
    assert np.array_equal(sa["site_lnl"], sb["site_lnl"])
    h = tree.post_order_ops()[-1].dst
    ca, xa = a.eng.get_clv(h)
    cb, xb = b.eng.get_clv(h)
    assert np.array_equal(ca, cb) and np.array_equal(xa, xb)
    # the edge kernels read the 1-byte codes (not the rows): lnL at a tip edge after the packed upload
    a.Set_Both_Sides(1)
    b.Set_Both_Sides(1)
    a.Lk()
    b.Lk()
    assert a.Lk(0) == b.Lk(0)


def test_lk_full_begin_wait_with_the_next_upload_in_between():
    """plk_lk_full_begin / plk_lk_wait: the tip codes of the NEXT evaluation are uploaded (own copy stream) while the
    current one is in flight; neither evaluation may see the other's data."""
    import torch

    from phyml_b200.engine import pack_codes4, pack_ops

    tree, m, pat = _synthetic(4, 40, 60000, seed=31, ambiguity=0.0)
    rng = np.random.default_rng(5)
    codes_a = pat.codes
    codes_b = np.ascontiguousarray(codes_a[rng.permutation(tree.n_otu)])   # same patterns, taxa shuffled: another lnL
    args = (tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges)
    eng = Engine(*args)
    eng.set_tip_table(pat.table())
    eng.set_weights(pat.wght, pat.invar)
    eng.set_model(m)
    edges = np.arange(tree.n_edges, dtype=np.int32)
    ops = pack_ops(tree.post_order_ops())
    left, rght = tree.edge_sides(tree.root_edge)
    full = eng.lk_full_call(edges, tree.l, ops, left, rght, tree.root_edge)
    begin = eng.lk_full_begin_call(edges, tree.l, ops, left, rght, tree.root_edge)
    pa = torch.from_numpy(pack_codes4(codes_a)).pin_memory()
    pb = torch.from_numpy(pack_codes4(codes_b)).pin_memory()
    eng.set_all_tip_codes_packed4(pa)
    ref_a = full()
    eng.set_all_tip_codes_packed4(pb)
    ref_b = full()
    assert ref_a != ref_b
    for _ in range(5):
        eng.set_all_tip_codes_packed4(pa)
        begin()                               # evaluation of A in flight ...
        eng.set_all_tip_codes_packed4(pb)     # ... while B is copied
        assert eng.lk_wait() == ref_a
        begin()
        eng.set_all_tip_codes_packed4(pa)
        assert eng.lk_wait() == ref_b
        begin()
        assert eng.lk_wait() == ref_a
