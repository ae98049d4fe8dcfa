"""Drop-in test (config 5 of BASELINE.json in miniature): the UNMODIFIED reference (its main(), spr.c,
optimiz.c, models.c ... built from /root/reference into oracle/_ref/libphyml_ref.so) driving the B200
engine through integration/lk_b200_shim.c, compared with the same reference running its own AVX
likelihood.  Both binaries are build products that travel to the GPU box; skipped if absent."""
import os
import re
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200 = os.path.join(ROOT, "integration", "_build", "phyml_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "phyml_ref")
GOLD = os.path.join(ROOT, "tests", "golden")

needs_bins = pytest.mark.skipif(not (os.path.exists(B200) and os.path.exists(REF)),
                                reason="drop-in binaries not built (need /root/reference at build time)")


def run(binary, tmp, args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    res = subprocess.run([binary] + args, cwd=tmp, capture_output=True, text=True, timeout=900, env=e)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    m = re.findall(r"Log likelihood of the current tree: (-?[0-9]+\.[0-9]+)", res.stdout)
    assert m, res.stdout[-2000:]
    return float(m[-1]), res.stdout


def stage(tmp_path, name):
    for ext in (".phy", ".nwk"):
        shutil.copy(os.path.join(GOLD, name + ext), tmp_path)
    return name + ".phy", name + ".nwk"


@needs_bins
def test_fixed_tree_lnl_matches_reference(tmp_path):
    """-o n: main() -> Lk(NULL) twice; value printed with 21 decimals must agree to 1e-11 relative."""
    phy, nwk = stage(tmp_path, "synth_dna_deep")
    args = ["-i", phy, "-u", nwk, "-d", "nt", "-m", "GTR", "-c", "4", "-a", "0.5", "-f", "e", "-o", "n", "-b", "0",
            "--r_seed", "1", "--no_memory_check"]
    a, _ = run(B200, str(tmp_path), args)
    b, _ = run(REF, str(tmp_path), args)
    assert abs(a - b) <= 1e-11 * abs(b), (a, b)
    assert abs(b - (-21640.146617685834)) <= 1e-9 * abs(b)


@needs_bins
def test_branch_length_and_rate_optimisation(tmp_path):
    """-o lr: Round_Optimize / Br_Len_Opt / dLk of optimiz.c, unchanged, on the GPU engine."""
    phy, nwk = stage(tmp_path, "synth_dna_deep")
    args = ["-i", phy, "-u", nwk, "-d", "nt", "-m", "HKY85", "-c", "4", "-f", "e", "-o", "lr", "-b", "0",
            "--r_seed", "1", "--no_memory_check"]
    a, out = run(B200, str(tmp_path), args, {"PLK_SHIM_VERBOSE": "1"})
    b, _ = run(REF, str(tmp_path), args)
    assert "phyml_b200: Lk" in out
    assert abs(a - b) <= 1e-6 * abs(b), (a, b)


@needs_bins
def test_spr_search(tmp_path):
    """-o tlr -s SPR: Global_Spr_Search of spr.c (Prune/Graft pointer swaps, Update_Partial_Lk per
    candidate, Lk(b), Triple_Dist) on the GPU engine; the final lnL must match the CPU run's."""
    phy, nwk = stage(tmp_path, "synth_dna_deep")
    # 200 taxa is a long CPU search: use the first 24 taxa of the alignment
    lines = open(os.path.join(str(tmp_path), phy)).read().splitlines()
    n_sites = lines[0].split()[1]
    with open(os.path.join(str(tmp_path), "small.phy"), "w") as f:
        f.write(f"24 {n_sites}\n" + "\n".join(lines[1:25]) + "\n")
    args = ["-i", "small.phy", "-d", "nt", "-m", "HKY85", "-c", "4", "-a", "0.5", "-f", "e", "-o", "tlr", "-s", "SPR",
            "-b", "0", "--r_seed", "1", "--no_memory_check"]
    a, out = run(B200, str(tmp_path), args, {"PLK_SHIM_VERBOSE": "1"})
    topo_a = _topology(str(tmp_path), "small.phy")
    b, _ = run(REF, str(tmp_path), args)
    topo_b = _topology(str(tmp_path), "small.phy")
    assert abs(a - b) <= 1e-5 * abs(b), (a, b)
    assert topo_a == topo_b
    # the parsimony pre-filter of the search (spr_pars, init.c:786; pars.c) ran on the device as well
    m = re.search(r"phyml_b200: Pars (\d+)  Update_Partial_Pars (\d+)", out)
    assert m and int(m.group(1)) > 0 and int(m.group(2)) > 0, out[-1500:]


@needs_bins
def test_spr_search_step_matrix_parsimony(tmp_path):
    """--g_pars: the step-matrix (Sankoff) branch of pars.c:355-372,409-431 as the pre-filter of the same search."""
    phy, nwk = stage(tmp_path, "synth_dna_deep")
    lines = open(os.path.join(str(tmp_path), phy)).read().splitlines()
    n_sites = lines[0].split()[1]
    with open(os.path.join(str(tmp_path), "small.phy"), "w") as f:
        f.write(f"14 {n_sites}\n" + "\n".join(lines[1:15]) + "\n")
    args = ["-i", "small.phy", "-d", "nt", "-m", "HKY85", "-c", "4", "-a", "0.5", "-f", "e", "-o", "tlr", "-s", "SPR",
            "-b", "0", "--r_seed", "1", "--no_memory_check", "--g_pars"]
    a, out = run(B200, str(tmp_path), args, {"PLK_SHIM_VERBOSE": "1"})
    topo_a = _topology(str(tmp_path), "small.phy")
    b, _ = run(REF, str(tmp_path), args)
    assert abs(a - b) <= 1e-5 * abs(b), (a, b)
    assert topo_a == _topology(str(tmp_path), "small.phy")
    assert re.search(r"phyml_b200: Pars (\d+)", out)


@needs_bins
def test_rooted_tree(tmp_path):
    """A rooted t_tree (Add_Root, utilities.c:8426; the n_root branches of lk.c:529-576 and the rooted cases of
    Set_All_Partial_Lk, lk.c:2988-3194, with ignore_root == YES): oracle/ref_driver.c --rooted linked against the
    binding vs the same driver on the reference's CPU code.  main() never evaluates a rooted tree, hence the driver."""
    import json
    drv_b200 = os.path.join(ROOT, "integration", "_build", "ref_driver_b200")
    drv_ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not (os.path.exists(drv_b200) and os.path.exists(drv_ref)):
        pytest.skip("drivers not built")
    phy, nwk = stage(tmp_path, "synth_dna_deep")
    args = ["--rooted", "7", "--", "-i", phy, "-u", nwk, "-d", "nt", "-m", "GTR", "-c", "4", "-a", "0.5", "-f", "e", "-o", "n",
            "-b", "0", "--r_seed", "1", "--no_memory_check"]
    out = []
    for drv in (drv_b200, drv_ref):
        res = subprocess.run([drv] + args, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        out.append(json.loads(re.search(r"REF_ROOTED (\{.*\})", res.stdout).group(1)))
    a, b = out
    assert a["root_edge"] == b["root_edge"] == 7 and b["ignore_root"] == 1
    assert abs(a["lnL"] - b["lnL"]) <= 1e-11 * abs(b["lnL"]), (a["lnL"], b["lnL"])
    assert abs(a["lnL"] - (-21640.146617685834)) <= 1e-9 * abs(b["lnL"])
    assert len(a["edge_lnl"]) == len(b["edge_lnl"]) > 10
    for x, y in zip(a["edge_lnl"], b["edge_lnl"]):
        assert abs(x - y) <= 1e-11 * abs(y), (x, y)


@needs_bins
def test_protein_search_through_the_binding(tmp_path):
    """20 states end to end: LG + G4 on the amino-acid fixture (ambiguity codes B / Z / X included), SPR search with the
    parsimony pre-filter on 20-bit state sets; final lnL and topology must equal the CPU run's."""
    phy, nwk = stage(tmp_path, "synth_aa_small")
    args = ["-i", phy, "-d", "aa", "-m", "LG", "-c", "4", "-a", "0.7", "-f", "m", "-o", "tlr", "-s", "SPR", "-b", "0",
            "--r_seed", "1", "--no_memory_check"]
    a, out = run(B200, str(tmp_path), args, {"PLK_SHIM_VERBOSE": "1"})
    topo_a = _topology(str(tmp_path), phy)
    b, _ = run(REF, str(tmp_path), args)
    assert abs(a - b) <= 1e-5 * abs(b), (a, b)
    assert topo_a == _topology(str(tmp_path), phy)
    assert "ns=20" in out and re.search(r"phyml_b200: Pars (\d+)", out)


def _numeric_rows(path):
    """every line of a PhyML text output as a list of tokens, numbers converted to float"""
    rows = []
    for ln in open(path).read().splitlines():
        toks = []
        for t in ln.split():
            try:
                toks.append(float(t))
            except ValueError:
                toks.append(t)
        if toks:
            rows.append(toks)
    return rows


@needs_bins
def test_host_readers_ancestral_sequences_and_site_likelihoods(tmp_path):
    """--ancestral: Ancestral_Sequences (ancestral.c:527) reads every edge's CLVs, scalers and P-matrices on the HOST;
    --print_site_lnl: Print_Site_Lk (io.c:1870) reads cur_site_lk / unscaled_site_lk_cat / fact_sum_scale.  The binding
    backs the address-only arena with pages filled from the device while the reader runs, and mirrors the per-site arrays
    after Lk(NULL).  Both output files must equal the CPU run's."""
    phy, nwk = stage(tmp_path, "synth_dna_deep")
    lines = open(os.path.join(str(tmp_path), phy)).read().splitlines()
    n_sites = lines[0].split()[1]
    for name in ("a.phy", "b.phy"):
        with open(os.path.join(str(tmp_path), name), "w") as f:
            f.write(f"16 {n_sites}\n" + "\n".join(lines[1:17]) + "\n")
    args = ["-d", "nt", "-m", "HKY85", "-c", "4", "-a", "0.5", "-f", "e", "-o", "lr", "-b", "0", "--r_seed", "1",
            "--no_memory_check", "--ancestral", "--print_site_lnl"]
    a, _ = run(B200, str(tmp_path), ["-i", "a.phy"] + args)
    b, _ = run(REF, str(tmp_path), ["-i", "b.phy"] + args)
    assert abs(a - b) <= 1e-6 * abs(b), (a, b)
    for suffix, min_rows in (("_phyml_ancestral_seq.txt", 1000), ("_phyml_lk.txt", 100)):
        ra = _numeric_rows(os.path.join(str(tmp_path), "a.phy" + suffix))
        rb = _numeric_rows(os.path.join(str(tmp_path), "b.phy" + suffix))
        assert len(ra) == len(rb) >= min_rows, (suffix, len(ra), len(rb))
        for x, y in zip(ra, rb):
            assert len(x) == len(y), (x, y)
            for u, v in zip(x, y):
                if isinstance(v, float) and isinstance(u, float):
                    assert abs(u - v) <= 2e-4 * max(abs(v), 1e-6) + 1e-12, (suffix, x, y)
                elif "phy" not in str(v):   # file names differ (a.phy / b.phy)
                    assert u == v, (suffix, x, y)


def _topology(tmp, phy):
    """Newick of the tree PhyML wrote next to the alignment with branch lengths and supports removed."""
    txt = open(os.path.join(tmp, phy + "_phyml_tree.txt")).read().strip()
    return re.sub(r"\)[0-9.eE+-]+", ")", re.sub(r":[0-9.eE+-]+", "", txt))


def _supports(tmp, phy):
    """Branch supports (internal-node labels) of the tree PhyML wrote next to the alignment."""
    txt = open(os.path.join(tmp, phy + "_phyml_tree.txt")).read()
    return [float(x) for x in re.findall(r"\)([0-9.eE+-]+):", txt)]


@needs_bins
@pytest.mark.parametrize("bflag", [None, "-4", "-2"])
def test_branch_supports_match_reference(tmp_path, bflag):
    """aLRT-type branch supports (alrt.c:172): default aBayes (no -b: init.c:604), SH-like (-b -4) and
    Chi2-based aLRT (-b -2).  NNI_Neigh_BL reads tree->c_lnL_sorted right after edge-level Lk() calls
    (alrt.c:453,555,682): the binding mirrors it after every Lk(b) while aLRT() runs.  Supports must match the
    CPU run's."""
    phy, nwk = stage(tmp_path, "synth_dna_deep")
    lines = open(os.path.join(str(tmp_path), phy)).read().splitlines()
    n_sites = lines[0].split()[1]
    for name in ("a.phy", "b.phy"):
        with open(os.path.join(str(tmp_path), name), "w") as f:
            f.write(f"16 {n_sites}\n" + "\n".join(lines[1:17]) + "\n")
    args = ["-d", "nt", "-m", "HKY85", "-c", "4", "-a", "0.5", "-f", "e", "-o", "lr", "--r_seed", "1", "--no_memory_check"]
    if bflag is not None:
        args += ["-b", bflag]
    a, _ = run(B200, str(tmp_path), ["-i", "a.phy"] + args)
    b, _ = run(REF, str(tmp_path), ["-i", "b.phy"] + args)
    assert abs(a - b) <= 1e-6 * abs(b), (a, b)
    sa, sb = _supports(str(tmp_path), "a.phy"), _supports(str(tmp_path), "b.phy")
    assert len(sa) == len(sb) and len(sa) >= 10
    for x, y in zip(sa, sb):
        assert abs(x - y) <= 2e-3 * max(1.0, abs(y)), (sa, sb)


@needs_bins
def test_bootstrap_replicates_are_refused(tmp_path):
    """-b N (N > 0): replicate trees share the main tree's likelihood structures (utilities.c:4042) and have no
    device instance; the binding must refuse loudly instead of silently running the CPU code."""
    phy, nwk = stage(tmp_path, "synth_dna_deep")
    lines = open(os.path.join(str(tmp_path), phy)).read().splitlines()
    n_sites = lines[0].split()[1]
    with open(os.path.join(str(tmp_path), "a.phy"), "w") as f:
        f.write(f"8 {n_sites}\n" + "\n".join(lines[1:9]) + "\n")
    res = subprocess.run([B200, "-i", "a.phy", "-d", "nt", "-m", "HKY85", "-c", "4", "-o", "n", "-b", "2", "--r_seed", "1",
                          "--no_memory_check"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert res.returncode != 0
    assert "without a device instance" in (res.stdout + res.stderr)
