"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) times the reference's
own CPU implementation (oracle/_ref/ref_driver, built from /root/reference) on all host cores and prints one
JSON line with the contract's keys; helper arithmetic (algorithmic bytes of SURVEY.md section 8d)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="reference build (oracle/_ref) not present")
def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                          "--warmup", "3", "--workload", "dna_16x4k"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "site_edge_updates_per_s" and d["higher_is_better"] is True
    assert d["unit"] == "site*edge-updates/s" and d["value"] > 0 and d["steps"] == 3 and d["warmup"] == 3
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"]
    assert "sample" in cb and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f64"
    # a single-block workload that fits the budget is evaluated whole: the slices' lnL add up to the alignment's
    if 4096 // (os.cpu_count() or 1) >= 50:
        assert d.get("same_alignment") is True and d["sites_evaluated"] == 4096 // cb["cores"] * cb["cores"]
        assert d["lnL"] < 0 and abs(d["lnL"] - sum(cb["lnL_sample"])) < 1e-9


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "3", "--warmup", "3"], capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_algorithmic_bytes_match_survey_8d():
    sys.path.insert(0, ROOT)
    import bench
    from phyml_b200.tree import Tree

    tree = Tree.random(100, seed=1)
    ops = tree.post_order_ops()
    assert len(ops) == 98
    per_site = bench.k1_algorithmic_bytes(tree, ops, 1, 4, 4)
    # (n-2)(8 ns ncatg + 4) written + (n-3)(8 ns ncatg + 4) read from internal children + (n-1) tip bytes
    assert per_site == 98 * 132 + 97 * 132 + 99
    n_tt = sum(1 for o in ops if o.c1.is_tip and o.c2.is_tip)
    n_ii = sum(1 for o in ops if not o.c1.is_tip and not o.c2.is_tip)
    n_it = 98 - n_tt - n_ii
    assert per_site == n_tt * 134 + n_it * 265 + n_ii * 396
