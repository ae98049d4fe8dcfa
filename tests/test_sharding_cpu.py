"""World-size-2 (gloo, CPU) test of the site-sharding host logic: each rank evaluates its block of
patterns (oracle backend), partial lnL / dlnL are summed with an all-reduce, and the result must
equal the single-process value.  Covers LkTree.reduce_fn, Patterns.shard and shard_bounds."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case():
    from phyml_b200 import alignment, model as pmodel
    from phyml_b200.tree import Tree

    tree = Tree.random(14, seed=21)
    m = pmodel.gtr(alpha=0.6, pinv=0.1)
    pat = alignment.compress(alignment.simulate(tree, m, 1501, seed=22, ambiguity=0.03), 4)
    return tree, m, pat


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_backend import OracleBackend

    from phyml_b200.lk import LkTree
    from phyml_b200.sharding import dist_reduce_fn

    tree, m, pat = _case()
    sh = pat.shard(rank, world)
    t = LkTree(tree, sh, m, OracleBackend(tree.n_otu, sh.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges),
               reduce_fn=dist_reduce_fn())
    t.Set_Both_Sides(1)
    lnl = t.Lk()
    e = 5
    t.Set_Update_Eigen_Lr(1)
    t.Lk(e)
    t.Set_Update_Eigen_Lr(0)
    _, lnl2 = t.dLk(0.07, e)
    q.put((rank, sh.n_pattern, lnl, lnl2, t.c_dlnL))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_lnl_matches_single_process():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_backend import OracleBackend

    from phyml_b200.alignment import shard_bounds
    from phyml_b200.lk import LkTree

    tree, m, pat = _case()
    ref = LkTree(tree, pat, m, OracleBackend(tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges))
    ref.Set_Both_Sides(1)
    ref_lnl = ref.Lk()
    ref.Set_Update_Eigen_Lr(1)
    ref.Lk(5)
    ref.Set_Update_Eigen_Lr(0)
    _, ref_lnl2 = ref.dLk(0.07, 5)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sum(r[1] for r in res) == pat.n_pattern
    for r in res:
        assert abs(r[2] - ref_lnl) <= 1e-12 * abs(ref_lnl)
        assert abs(r[3] - ref_lnl2) <= 1e-12 * abs(ref_lnl2)
        assert abs(r[4] - ref.c_dlnL) <= 1e-9 * max(1.0, abs(ref.c_dlnL))


@pytest.mark.parametrize("n,world", [(10, 3), (7, 8), (100000, 8), (1, 1)])
def test_shard_bounds_partition(n, world):
    from phyml_b200.alignment import shard_bounds

    covered = []
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        assert 0 <= lo <= hi <= n
        covered += list(range(lo, hi))
    assert covered == list(range(n))
    sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
