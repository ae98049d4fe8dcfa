#!/usr/bin/env python
"""bench.py -- full-tree log-likelihood throughput of the B200 engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank/GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step = one full-tree lnL evaluation Lk(NULL) (all P-matrices, n-2 CLV updates in post-order,
edge reduction) of a fixed random tree on a synthetic alignment:
  N=1 workload: BASELINE.json configs[1], DNA 100 taxa x 100 000 sites, GTR+G4.
  N>1: each rank owns a contiguous block of 100 000 sites (weak scaling, site sharding); the only
       exchange is one ncclAllReduce of the partial lnL per evaluation.
metric = site*edge updates/s = patterns * (n_taxa - 2) * evaluations / s, whole job.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from phyml_b200 import alignment, model as pmodel  # noqa: E402
from phyml_b200.tree import Tree  # noqa: E402

METRIC = "site_edge_updates_per_s"
UNIT = "site*edge-updates/s"

WORKLOADS = {
    # name: (ns, n_taxa, sites_per_gpu, description)
    "dna_100x100k": (4, 100, 100_000, "synthetic DNA 100 taxa x 100000 sites, GTR+G4, fixed random tree (BASELINE configs[1])"),
    "aa_200x50k": (20, 200, 50_000, "synthetic AA 200 taxa x 50000 sites, LG+G4 (BASELINE configs[2])"),
    "dna_500x125k": (4, 500, 125_000, "synthetic DNA 500 taxa x 125000 sites per GPU, GTR+G4 (BASELINE configs[3] sharded 8 ways)"),
    "dna_100x50k": (4, 100, 50_000, "synthetic DNA 100 taxa x 50000 sites, GTR+G4 (BASELINE configs[4] alignment)"),
    "dna_16x4k": (4, 16, 4_096, "tiny DNA smoke workload"),
}


def make_workload(name, rank, world, seed=1):
    ns, n_taxa, sites, desc = WORKLOADS[name]
    tree = Tree.random(n_taxa, seed=seed)
    m = pmodel.gtr(alpha=0.5) if ns == 4 else pmodel.lg_from_fixture(alpha=0.5)
    # every rank simulates its own block of columns (same tree, different seed): site sharding
    codes = alignment.simulate(tree, m, sites, seed=1000 + rank)
    pat = alignment.compress(codes, ns)
    return tree, m, pat, codes, desc


def k1_algorithmic_bytes(tree, ops, P, ns, ncatg):
    """B1 = 8*ns*ncatg*(1+n_int) + 4*(1+n_int) + 1*n_tip per site and update."""
    tot = 0
    for o in ops:
        n_int = (0 if o.c1.is_tip else 1) + (0 if o.c2.is_tip else 1)
        tot += 8 * ns * ncatg * (1 + n_int) + 4 * (1 + n_int) + (2 - n_int)
    return tot * P


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None
        self.first = 0

    def mark(self):
        """Samples before this point (GPU idle) are not part of the statistics."""
        self.first = len(self.rows)

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.rows = self.rows[self.first:] or self.rows
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = float(self.rows[0][1])
            out["power_w_max"] = max(float(r[2]) for r in self.rows if len(r) >= 7)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for j, nm in enumerate(names):
                if any(len(r) >= 7 and r[3 + j].lower().startswith("active") for r in self.rows):
                    out["reasons"].append(nm)
            out["samples"] = len(sm)
        return out


# ======================================================================================================
def run_b200(args):
    import torch
    import torch.distributed as dist

    from phyml_b200.engine import Engine, pack_ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    tree, m, pat, codes, desc = make_workload(args.workload, rank, world)
    ns, ncatg, P, n = m.ns, m.ncatg, pat.n_pattern, tree.n_otu
    eng = Engine(n, P, ns, ncatg, tree.n_clv_handles, tree.n_edges, device=local)
    if world > 1:
        # cross-GPU sum of the partial lnL: fused into the reduction kernel over NVLink peer memory
        # (PLK_COMM=nccl selects the ncclAllReduce fallback)
        from phyml_b200.sharding import init_engine_comm

        init_engine_comm(eng, rank, world, mode=os.environ.get("PLK_COMM", "p2p"))

    ops = tree.post_order_ops()
    ops_packed = pack_ops(ops)
    edges = np.arange(tree.n_edges, dtype=np.int32)
    lengths = tree.l.copy()
    left, rght = tree.edge_sides(tree.root_edge)
    stream = torch.cuda.ExternalStream(eng.stream, device=local)

    # host-resident inputs in pinned memory (the e2e leg re-uploads them every step)
    h_codes = torch.from_numpy(pat.codes).pin_memory()
    h_wght = torch.from_numpy(pat.wght).pin_memory()
    h_invar = torch.from_numpy(pat.invar).pin_memory()
    h_site = torch.empty(P, dtype=torch.float64).pin_memory()

    def upload_inputs():
        eng.set_all_tip_codes(h_codes)
        eng.set_weights_ptr(h_wght.data_ptr(), h_invar.data_ptr())
        eng.set_model(m)

    def evaluate(ev=None):
        eng.update_pmats(edges, lengths)                      # K0, all edges (lk.c:500-505)
        if ev is not None:
            ev[0].record(stream)
        eng.update_partials(ops_packed)                       # K1, post-order (lk.c:562)
        if ev is not None:
            ev[1].record(stream)
        return eng.edge_lnl(left, rght, tree.root_edge)       # K2 (+ all-reduce when sharded)

    eng.set_tip_table(pat.table())
    upload_inputs()
    lnl0 = evaluate()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg: `value`
    for _ in range(args.warmup):
        evaluate()
    sampler = ClockSampler(local)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 15.0:   # nvidia-smi needs a moment to start
        time.sleep(0.05)
    k1_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launch_count
    barrier()
    sampler.mark()
    e0.record(stream)
    for s in range(args.steps):
        lnl = evaluate(k1_ev[s])
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    k1_ms = [a.elapsed_time(b) for a, b in k1_ev]
    # keep the GPU under the same load until the sampler has seen it (its period is 100 ms, the timed leg ~10 ms);
    # it is stopped before the e2e leg, whose many small driver calls an NVML poller would perturb
    n_rows = len(sampler.rows)
    t_wait = time.time()
    while len(sampler.rows) < n_rows + 3 and time.time() - t_wait < 3.0:
        # LOCAL work only (K0 + K1, no reduction): the number of iterations differs between ranks, so
        # nothing in this loop may involve the cross-GPU exchange
        eng.update_pmats(edges, lengths)
        eng.update_partials(ops_packed)
        eng.sync()
    clocks = sampler.finish()

    # ---------------- end-to-end leg: host buffers in, lnL (+ per-site lnL) out, every step
    for _ in range(max(1, args.warmup // 2)):
        upload_inputs()
        evaluate()
        eng.get_site_lnl_ptr(h_site.data_ptr())
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    for s in range(args.steps):
        upload_inputs()
        lnl_e2e = evaluate()
        eng.get_site_lnl_ptr(h_site.data_ptr())
    g1.record(stream)
    barrier()
    e2e_ms = max(g0.elapsed_time(g1), (time.perf_counter() - t0) * 1e3)
    assert abs(lnl_e2e - lnl) <= 1e-12 * abs(lnl)
    assert abs(float(np.dot(h_site.numpy(), pat.wght)) - (lnl if world == 1 else float("nan"))) <= 1e-9 * abs(lnl) or world > 1

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e_ms, float(P)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e2e_ms, P_total = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        P_total = float(P)

    updates = P_total * (n - 2) * args.steps
    value = updates / (ms * 1e-3)
    e2e_value = updates / (e2e_ms * 1e-3)
    h2d = int(h_codes.numel() + h_wght.numel() * 8 + h_invar.numel() * 2 + (2 * ns * ns + 2 * ns + 2 * ncatg + 4) * 8
              + tree.n_edges * 16 + len(ops) * 28)
    d2h = int(P * 8 + 32)

    out = None
    if rank == 0:
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        k1_bytes = k1_algorithmic_bytes(tree, ops, P, ns, ncatg)
        k1_avg_ms = statistics.mean(k1_ms)
        achieved = k1_bytes / (k1_avg_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "name": args.workload, "n_taxa": n, "sites_per_gpu": WORKLOADS[args.workload][2],
                       "patterns_per_gpu": P, "ns": ns, "ncatg": ncatg, "updates_per_eval": n - 2,
                       "parallelism": f"site-shard x{world}", "both_sides": False,
                       "exchange": ("none" if world == 1 else os.environ.get("PLK_COMM", "p2p")),
                       "l2": "CLV working set %.2f GB per evaluation > 126 MB L2 (no flush needed)"
                             % ((n - 2) * P * ns * ncatg * 8 / 1e9)},
            "evals_per_s": args.steps / (ms * 1e-3), "lnL": lnl,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_traverse_dna<4,2>" if ns == 4 else ("k_traverse_aa (DMMA.8x8x4)" if ns == 20 else "k_partial_generic"),
                         "achieved": achieved, "peak": peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "algorithmic_bytes_per_eval": k1_bytes, "k1_ms_per_eval": k1_avg_ms,
                         "k1_launches_per_eval": None},
        }
        if ns == 20:
            # the 20-state path is the dense contraction on the FP64 tensor pipe: also report its flop rate
            n_int = sum((0 if o.c1.is_tip else 1) + (0 if o.c2.is_tip else 1) for o in ops)
            flops = 2.0 * 20 * 20 * ncatg * P * n_int            # useful MACs*2 of the P.x products
            out["roofline"]["fp64_tensor"] = {"achieved_tflops": flops / (k1_avg_ms * 1e-3) / 1e12,
                                              "nominal_peak_tflops": 40.0,
                                              "note": "useful flops of the (20x20).(20xsites) contractions; MMA tiles are padded 20->24"}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                out["roofline"]["traffic"] = json.load(open(tr)).get(args.workload)
            except Exception:
                pass
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(tree, codes, m, cores=1, sites=args.cpu_sites, evals=args.cpu_evals)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    if rank == 0:
        print(json.dumps(out))


# ======================================================================================================
def cpu_baseline(tree, codes, m, cores, sites, evals, warm=1):
    """Times the reference's own Lk(NULL) (oracle/_ref/ref_driver = the unmodified reference built
    from /root/reference sources, AVX2+FMA path) on `cores` processes, each on its own slice of
    columns (the reference is single-threaded).  Falls back to the C oracle port if _ref is absent."""
    n = tree.n_otu
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    sites = min(sites, codes.shape[1])
    if os.path.exists(driver):
        with tempfile.TemporaryDirectory() as wd:
            with open(os.path.join(wd, "tree.nwk"), "w") as f:
                f.write(tree.to_newick() + "\n")
            procs = []
            per = sites // cores
            for c in range(cores):
                phy = os.path.join(wd, f"aln{c}.phy")
                alignment.write_phylip(phy, codes[:, c * per:(c + 1) * per], m.ns, tree.names)
                dt = ["-d", "nt", "-m", "GTR"] if m.ns == 4 else ["-d", "aa", "-m", "LG"]
                cmd = [driver, "--time", str(evals), "--warmup", str(warm), "--", "-i", phy, "-u",
                       os.path.join(wd, "tree.nwk")] + dt + ["-c", "4", "-a", "0.5", "-f", "e" if m.ns == 4 else "m",
                                                              "-o", "n", "-b", "0", "--r_seed", "1", "--no_memory_check"]
                procs.append(subprocess.Popen(cmd, cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
            res = []
            for p in procs:
                outp = p.communicate()[0]
                line = [ln for ln in outp.splitlines() if ln.startswith("REF_TIMING")]
                if not line:
                    raise RuntimeError("ref_driver produced no timing")
                res.append(json.loads(line[0][len("REF_TIMING "):]))
        slowest = max(r["mean_s"] for r in res)
        pats = sum(r["n_pattern"] for r in res)
        return {"value": pats * (n - 2) / slowest, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"{n} taxa x {per * cores} sites ({pats} patterns), {evals} timed Lk(NULL) per process after "
                          f"{warm} warm-up, {cores} process(es) x 1 thread, gcc -O3 -march=haswell (AVX2+FMA kernels)",
                "s_per_eval": slowest, "lnL_sample": [r["lnL"] for r in res]}
    # port: C oracle through the test backend
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_backend import OracleBackend
    from phyml_b200.lk import LkTree

    sites = min(sites, 5000)
    pat = alignment.compress(codes[:, :sites], m.ns)
    t = LkTree(tree, pat, m, OracleBackend(n, pat.n_pattern, m.ns, m.ncatg, tree.n_clv_handles, tree.n_edges))
    t.Lk()
    t0 = time.perf_counter()
    for _ in range(evals):
        t.Lk()
    dt = (time.perf_counter() - t0) / evals
    return {"value": pat.n_pattern * (n - 2) / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} taxa x {sites} sites, {evals} evals, scalar C oracle", "s_per_eval": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ns, n_taxa, sites, desc = WORKLOADS[args.workload]
    tree = Tree.random(n_taxa, seed=1)
    m = pmodel.gtr(alpha=0.5) if ns == 4 else pmodel.lg_from_fixture(alpha=0.5)
    cores = os.cpu_count() or 1
    # bounded sample: every host core gets its own block of columns of the workload's shape (sites are
    # independent, the reference is single-threaded: one process per core).  1 000 columns per core keeps a
    # process's likelihood arena near 38 MB: measured on the 128-core GPU-box host this is the reference's
    # best case (2.2e8 updates/s; 3 000 columns per core drop to 1.6e8, memory-bandwidth bound)
    per_core = 1000 if ns == 4 else 200
    budget = int(1.5e7 * 90 / max(1, (args.steps + args.warmup)) / (n_taxa - 2))   # ~90 s at 1.5e7 updates/s/core
    per_core = max(200, min(per_core, budget))
    codes = alignment.simulate(tree, m, per_core * cores, seed=1000)
    cb = cpu_baseline(tree, codes, m, cores=cores, sites=per_core * cores, evals=args.steps, warm=args.warmup)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["s_per_eval"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": desc, "name": args.workload, "n_taxa": n_taxa, "ns": ns, "ncatg": 4,
                      "note": "reference CPU implementation (host cores); each step = one Lk(NULL) on a bounded column sample"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dna_100x100k", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sites", type=int, default=25_000)
    ap.add_argument("--cpu-evals", type=int, default=40)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
