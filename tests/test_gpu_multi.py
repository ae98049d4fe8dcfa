"""Two-GPU test of the site-sharded engine (skipped with fewer than 2 devices): each rank evaluates its
block of patterns; the cross-GPU sum happens inside the reduction kernel over NVLink peer memory
(plk_comm_p2p_*) or through the in-engine ncclAllReduce (plk_comm_init).  Both must return, on every rank,
the single-GPU value; the P2P result must be bitwise identical across ranks (rank-order addition)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case():
    from phyml_b200 import alignment, model as pmodel
    from phyml_b200.tree import Tree

    tree = Tree.random(18, seed=31)
    m = pmodel.gtr(alpha=0.5)
    pat = alignment.compress(alignment.simulate(tree, m, 6001, seed=32, ambiguity=0.02), 4)
    return tree, m, pat


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from phyml_b200.engine import Engine
    from phyml_b200.lk import LkTree
    from phyml_b200.sharding import init_engine_comm

    tree, m, pat = _case()
    sh = pat.shard(rank, world)
    eng = Engine(tree.n_otu, sh.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges, device=rank)
    init_engine_comm(eng, rank, world, mode=mode)
    t = LkTree(tree, sh, m, eng)
    t.Set_Both_Sides(1)
    lnl = t.Lk()
    t.Set_Update_Eigen_Lr(1)
    t.Lk(3)
    t.Set_Update_Eigen_Lr(0)
    t.dLk(0.05, 3)
    q.put((rank, lnl, t.c_lnL, t.c_dlnL))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_two_gpu_sharded_lnl(mode):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from phyml_b200.engine import Engine
    from phyml_b200.lk import LkTree

    tree, m, pat = _case()
    ref = LkTree(tree, pat, m, Engine(tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges))
    ref.Set_Both_Sides(1)
    ref_lnl = ref.Lk()
    ref.Set_Update_Eigen_Lr(1)
    ref.Lk(3)
    ref.Set_Update_Eigen_Lr(0)
    ref.dLk(0.05, 3)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in res:
        assert abs(r[1] - ref_lnl) <= 1e-12 * abs(ref_lnl)
        assert abs(r[2] - ref.c_lnL) <= 1e-12 * abs(ref.c_lnL)
        assert abs(r[3] - ref.c_dlnL) <= 1e-9 * max(1.0, abs(ref.c_dlnL))
    if mode == "p2p":
        assert res[0][1:] == res[1][1:], "rank-order addition must give bitwise identical results on all ranks"


def test_single_process_sharded_instance_two_devices():
    """plk_create_sharded over devices 0 and 1 in THIS process (the way lk.c's single t_tree drives a box,
    SURVEY.md section 8e): the exchange runs inside the reduction kernels over peer memory enabled in-process;
    lnL / dLk / CLV read-backs must equal the single-device instance's."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from phyml_b200.engine import Engine
    from phyml_b200.lk import LkTree

    tree, m, pat = _case()
    args = (tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges)
    one = LkTree(tree, pat, m, Engine(*args))
    two = LkTree(tree, pat, m, Engine(*args, devices=[0, 1]))
    assert two.eng.n_shards == 2
    for t in (one, two):
        t.Set_Both_Sides(1)
    a, b = one.Lk(), two.Lk()
    assert abs(a - b) <= 1e-12 * abs(a)
    for e in (0, 3, tree.n_edges - 1):
        assert abs(one.Lk(e) - two.Lk(e)) <= 1e-12 * abs(a)
    for t in (one, two):
        t.Set_Update_Eigen_Lr(1)
        t.Lk(3)
        t.Set_Update_Eigen_Lr(0)
        t.dLk(0.05, 3)
    assert abs(one.c_lnL - two.c_lnL) <= 1e-12 * abs(one.c_lnL)
    assert abs(one.c_dlnL - two.c_dlnL) <= 1e-9 * max(1.0, abs(one.c_dlnL))
    h = tree.post_order_ops()[5].dst
    ca, sa = one.eng.get_clv(h)
    cb, sb = two.eng.get_clv(h)
    assert np.array_equal(ca, cb) and np.array_equal(sa, sb)
    s1, s2 = one.eng.get_site_lnl(), two.eng.get_site_lnl()
    np.testing.assert_allclose(s1["site_lnl"], s2["site_lnl"], rtol=1e-13)


def test_single_process_sharded_parsimony_and_spr_candidates_two_devices():
    """the round-2 entry points on a one-process, two-device instance: parsimony totals travel through the same
    in-kernel exchange as lnL (exact integers), fractional weights through the serial chain in shard order, batched SPR
    candidate scores are all-shard sums."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import pars_checks as pk
    from oracle_backend import OracleBackend
    from phyml_b200.engine import Engine
    from phyml_b200.lk import LkTree

    c, g = pk.load("proteic_lg")
    for general in (False, True):
        eng = Engine(c.n_otu, c.P, c.ns, c.ncatg, c.tree.n_clv_handles, c.tree.n_edges, devices=[0, 1])
        pk.check_full(c, g, eng, general, True)
    tr, ui, w, step = pk.random_case(14, 4099, 4, seed=8, frac_weights=True)
    args = (tr.n_otu, 4099, 4, 1, tr.n_clv_handles, tr.n_edges)
    a = pk.run_random(tr, ui, w, step, Engine(*args, devices=[0, 1]), False)
    b = pk.run_random(tr, ui, w, step, OracleBackend(*args), False)
    assert a[0] == b[0] and (a[1] == b[1]).all() and a[2] == b[2]

    tree, m, pat = _case()
    args = (tree.n_otu, pat.n_pattern, 4, 4, tree.n_clv_handles, tree.n_edges)
    one = LkTree(tree, pat, m, Engine(*args))
    two = LkTree(tree, pat, m, Engine(*args, devices=[0, 1]))
    for t in (one, two):
        t.Set_Both_Sides(1)
        t.Lk()
    te = tree.adj[3][0][0]
    cands = []
    for e in range(tree.n_edges):
        if e != te:
            x, y = tree.edge_sides(e)
            cands.append((x, 0.5 * tree.l[e], y, 0.5 * tree.l[e]))
    r1, _ = one.eng.spr_candidates(tree.side_of(te, 3), float(tree.l[te]), True, cands)
    r2, _ = two.eng.spr_candidates(tree.side_of(te, 3), float(tree.l[te]), True, cands)
    assert (np.abs(r1 - r2) <= 1e-12 * np.abs(r1)).all()
