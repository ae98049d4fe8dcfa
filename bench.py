#!/usr/bin/env python
"""bench.py -- full-tree log-likelihood throughput of the B200 engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank/GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step = one full-tree lnL evaluation Lk(NULL) (all P-matrices, n-2 CLV updates in post-order,
edge reduction) of a fixed random tree on a synthetic alignment (phyml_b200/workloads.py):
  N=1 workload: BASELINE.json configs[1], DNA 100 taxa x 100 000 sites, GTR+G4.
  N>1 workload: BASELINE.json configs[3], DNA 500 taxa x 1 000 000 sites, GTR+G4, STRONG scaling:
       the alignment's 16 column blocks are split evenly over the ranks (site sharding); the only
       exchange is the sum of the per-rank partial lnL, fused into the reduction kernel over NVLink
       peer memory (PLK_COMM=nccl: one ncclAllReduce per evaluation).
  Both are evaluated under the reference's own eigen system for that CLI and compared with the
  reference's own lnL on the SAME alignment and tree (tests/golden/big/*.npz -> `lnL_reference`,
  `parity_rel_err`).
  The N=1 line also carries, under `secondary`, BASELINE configs[2] (AA 200 x 50 000, the 20-state
  tensor-pipe path) and configs[3] on ONE GPU (the strong-scaling base of the N>1 lines).
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

from phyml_b200 import alignment, workloads as wl  # noqa: E402

METRIC = "site_edge_updates_per_s"
UNIT = "site*edge-updates/s"
PARITY_RTOL = 1e-9   # BASELINE.json north_star


def default_workload(gpus):
    return "dna_100x100k" if gpus == 1 else "dna_500x1M"


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


def load_peaks():
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        try:
            return json.load(open(pk_path))
        except Exception:
            pass
    return {}


def k1_kernel_name(ns, ncatg):
    if ns == 4 and ncatg in (1, 2, 4, 8):
        v = int(os.environ.get("PLK_T2_VARIANT", "20"))
        if os.environ.get("PLK_TRAV_V1", "0") not in ("", "0"):
            return "k_traverse_dna<%d,2>" % ncatg
        if v < 10:
            return "k_traverse_dna2<%d>" % ncatg
        if v < 20:
            return "k_traverse_dna3<%d> (DMMA.8x8x4)" % ncatg
        return "k_traverse_dna4<%d,15> (DMMA.8x8x4)" % ncatg
    if ns == 20:
        if ncatg in (1, 2, 4, 8) and os.environ.get("PLK_AA_V1", "0") in ("", "0"):
            return "k_traverse_aa3<%d> (DMMA.8x8x4)" % ncatg
        return "k_traverse_aa (DMMA.8x8x4)"
    return "k_partial_generic"


# ======================================================================================================
def measure(name, rank, world, local, steps, warmup, e2e=True, clocks=True, procs=8):
    """One workload on this rank's GPU: device-resident leg, end-to-end leg, parity against the reference pin.
    Returns the per-rank record (rank 0's record carries the whole-job numbers after the max-over-ranks)."""
    import torch
    import torch.distributed as dist

    from phyml_b200.engine import Engine, pack_codes4, pack_ops

    w = wl.WORKLOADS[name]
    strong = w.n_blocks > 1
    m, pin = wl.evaluation_model(name)
    tree = wl.make_tree(w)
    blocks = wl.rank_blocks(w, rank, world) if (strong or world > 1) else [0]
    pat = wl.make_patterns(name, blocks, procs=min(procs, len(blocks)))
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
    # (4-state data: 4-bit tip codes, two patterns per byte -- plk_set_all_tip_codes_packed4)
    packed = ns == 4 and os.environ.get("PLK_BENCH_UNPACKED_TIPS") is None
    h_codes = torch.from_numpy(pack_codes4(pat.codes) if packed else pat.codes).pin_memory()
    h_wght = torch.from_numpy(pat.wght).pin_memory()
    h_invar = torch.from_numpy(pat.invar).pin_memory()
    h_site = torch.empty(P, dtype=torch.float64).pin_memory()

    def upload_inputs():
        if packed:
            eng.set_all_tip_codes_packed4(h_codes)
        else:
            eng.set_all_tip_codes(h_codes)
        eng.set_weights_ptr(h_wght.data_ptr(), h_invar.data_ptr())
        eng.set_model(m)

    # one step = Lk(NULL): K0 for all edges (lk.c:500-505), post-order traversal (lk.c:562) and the site loop at the
    # root edge (lk.c:578-645) through ONE C-ABI call (plk_lk_full); for 4-state / 4-category data the edge reduction
    # (+ the cross-GPU sum when sharded) is the epilogue of the traversal kernel: two launches per step
    evaluate = eng.lk_full_call(edges, lengths, ops_packed, left, rght, tree.root_edge)

    eng.set_tip_table(pat.table())
    upload_inputs()
    evaluate()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg: `value`
    for _ in range(warmup):
        evaluate()
    sampler = None
    if clocks:
        sampler = ClockSampler(local)
        sampler.start()
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 15.0:   # nvidia-smi needs a moment to start
            time.sleep(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launch_count
    barrier()
    if sampler:
        sampler.mark()
    e0.record(stream)
    for s in range(steps):
        lnl = evaluate()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    # the dominant kernel alone (roofline): the same traversal launched stand-alone (plk_update_partials, i.e. without
    # the fused edge epilogue), bracketed by CUDA events on the engine's stream, same process, right after the timed leg
    k1_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    l0 = eng.launch_count
    for a, b in k1_ev:
        a.record(stream)
        eng.update_partials(ops_packed)
        b.record(stream)
    eng.sync()
    k1_launches = (eng.launch_count - l0) // steps
    k1_ms = [a.elapsed_time(b) for a, b in k1_ev]
    clk = None
    if sampler:
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
        clk = sampler.finish()

    # ---------------- end-to-end leg: host buffers in, lnL (+ per-site lnL) out, every step
    e2e_ms = None
    lnl_e2e = lnl
    e2e_pipelined = os.environ.get("PLK_BENCH_E2E_SERIAL") is None
    if e2e:
        begin = eng.lk_full_begin_call(edges, lengths, ops_packed, left, rght, tree.root_edge)
        for _ in range(max(2, warmup // 2)):
            upload_inputs()
            begin()
            eng.lk_wait()
            eng.get_site_lnl_ptr(h_site.data_ptr())
        barrier()
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        if e2e_pipelined:
            # every step still copies its own inputs host -> device and reads lnL + per-site lnL back, but step s+1's
            # upload is issued while step s runs (plk_lk_full_begin / plk_lk_wait): the tip-code copy (its own copy
            # stream) overlaps the traversal, the small uploads queue behind it on the instance's stream
            upload_inputs()
            for s in range(steps):
                begin()
                if s + 1 < steps:
                    upload_inputs()
                lnl_e2e = eng.lk_wait()
                eng.get_site_lnl_ptr(h_site.data_ptr())
        else:
            for s in range(steps):
                upload_inputs()
                lnl_e2e = evaluate()
                eng.get_site_lnl_ptr(h_site.data_ptr())
        g1.record(stream)
        barrier()
        e2e_ms = max(g0.elapsed_time(g1), (time.perf_counter() - t0) * 1e3)
    else:
        eng.get_site_lnl_ptr(h_site.data_ptr())
        eng.sync()
    if abs(lnl_e2e - lnl) > 1e-12 * abs(lnl):
        raise RuntimeError(f"e2e lnL {lnl_e2e!r} differs from the device-resident lnL {lnl!r}")

    # the exchange checked against an independent sum: this rank's partial (weights . per-pattern lnL on the host)
    local_partial = float(np.dot(h_site.numpy(), pat.wght))
    if world > 1:
        t = torch.tensor([local_partial, float(P)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        host_sum, P_total = float(t[0]), float(t[1])
        tm = torch.tensor([ms, e2e_ms or 0.0, statistics.mean(k1_ms)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms, e2e_ms_max = float(tm[0]), float(tm[1])
        e2e_ms = e2e_ms_max if e2e else None
    else:
        host_sum, P_total = local_partial, float(P)
    exchange_rel_err = abs(host_sum - lnl) / abs(lnl)
    if exchange_rel_err > 1e-11:
        raise RuntimeError(f"all-rank lnL {lnl!r} differs from the sum of the per-rank partials {host_sum!r}")

    lnl_ref = None
    parity = None
    if pin is not None and (strong or world == 1):
        lnl_ref = float(pin["lnL"]) if (strong or w.n_blocks == 1) else None
        if lnl_ref is not None:
            parity = abs(lnl - lnl_ref) / abs(lnl_ref)
            if parity > PARITY_RTOL:
                raise RuntimeError(f"{name}: lnL {lnl!r} vs the reference's {lnl_ref!r}: relative error {parity:.3e} > {PARITY_RTOL}")

    updates = P_total * (n - 2) * steps
    k1_bytes = k1_algorithmic_bytes(tree, ops, P, ns, ncatg)
    k1_avg_ms = statistics.mean(k1_ms)
    h2d = int(h_codes.numel() + h_wght.numel() * 8 + h_invar.numel() * 2 + (2 * ns * ns + 2 * ns + 2 * ncatg + 4) * 8
              + tree.n_edges * 16 + len(ops) * 28)
    d2h = int(P * 8 + 32)
    rec = {
        "name": name, "desc": w.desc, "strong": strong, "n_taxa": n, "ns": ns, "ncatg": ncatg, "P": P, "P_total": P_total,
        "n_sites_total": w.block_sites * (w.n_blocks if strong else world), "blocks": blocks,
        "ms": ms, "steps": steps, "value": updates / (ms * 1e-3), "launches": int(launches), "lnL": lnl,
        "e2e_pipelined": bool(e2e and e2e_pipelined),
        "lnL_reference": lnl_ref, "parity_rel_err": parity, "exchange_rel_err": exchange_rel_err,
        "e2e_ms": e2e_ms, "e2e_value": (updates / (e2e_ms * 1e-3)) if e2e_ms else None, "h2d": h2d, "d2h": d2h,
        "k1_bytes": k1_bytes, "k1_ms": k1_avg_ms, "k1_launches": int(k1_launches), "clocks": clk,
        "updates_per_eval": n - 2, "device_bytes": eng.device_bytes, "tree": tree, "model": m, "ops": ops,
    }
    eng.close()
    return rec


def roofline_of(rec, peaks):
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = rec["k1_bytes"] / (rec["k1_ms"] * 1e-3) / 1e9
    out = {"bound": "hbm", "kernel": k1_kernel_name(rec["ns"], rec["ncatg"]), "achieved": achieved, "peak": peak,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
           "unit": "GB/s", "frac": achieved / peak, "traffic": None, "traffic_source": None,
           "algorithmic_bytes_per_eval": rec["k1_bytes"], "k1_ms_per_eval": rec["k1_ms"],
           "k1_launches_per_eval": rec["k1_launches"]}
    if rec["ns"] == 20:
        # the 20-state path is the dense contraction on the FP64 tensor pipe: also report its flop rate
        n_int = sum((0 if o.c1.is_tip else 1) + (0 if o.c2.is_tip else 1) for o in rec["ops"])
        flops = 2.0 * 20 * 20 * rec["ncatg"] * rec["P"] * n_int            # useful MACs*2 of the P.x products
        out["fp64_tensor"] = {"achieved_tflops": flops / (rec["k1_ms"] * 1e-3) / 1e12, "nominal_peak_tflops": 40.0,
                              "note": "useful flops of the (20x20).(20xsites) contractions; MMA tiles are padded 20->24"}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            t = json.load(open(tr))
            out["traffic"] = t.get(rec["name"])
            if out["traffic"] is not None:
                out["traffic_source"] = t.get("_source", "profiles/ (ncu --set full capture of this command, dram read + write per K1 launch; static, not re-measured in this run)")
        except Exception:
            pass
    return out


def measured_traffic(name, timeout=120):
    """DRAM bytes (read + write) of ONE launch of the traversal kernel, measured now: this same script re-run for one
    step under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` (warm launch: 3 skipped).  Only the byte counters
    are taken from the profiled run -- never a time.  Returns (bytes, source) or (None, why)."""
    import shutil

    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    if os.environ.get("CUDA_INJECTION64_PATH") or any(k.startswith("NV_NSIGHT") for k in os.environ):
        return None, "already running under a profiler"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base",
           "-k", "regex:k_traverse", "-s", "3", "-c", "1", "--csv", sys.executable, os.path.abspath(__file__),
           "--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-secondary", "--no-traffic", "--workload", name]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    except Exception as ex:  # noqa: BLE001
        return None, f"ncu pass failed: {type(ex).__name__}"
    tot, kern = 0.0, None
    for ln in res.stdout.splitlines():
        if "dram__bytes_" in ln and ln.startswith('"'):
            f = [x.strip('"') for x in ln.split('","')]
            try:
                tot += float(f[-1].replace(",", ""))
                kern = f[4].split("(")[0]
            except (ValueError, IndexError):
                pass
    if tot <= 0.0:
        return None, "ncu pass produced no dram counters"
    return tot, f"measured in this run: ncu dram__bytes_read.sum + dram__bytes_write.sum of one warm launch of {kern}"


def secondary_of(rec, peaks):
    rf = roofline_of(rec, peaks)
    return {"workload": rec["desc"], "name": rec["name"], "patterns": rec["P"], "n_taxa": rec["n_taxa"],
            "value": rec["value"], "unit": UNIT, "ms_per_step": rec["ms"] / rec["steps"], "steps": rec["steps"],
            "lnL": rec["lnL"], "lnL_reference": rec["lnL_reference"], "parity_rel_err": rec["parity_rel_err"],
            "gpu_launches": rec["launches"], "device_gb": rec["device_bytes"] / 1e9,
            "e2e": None if rec["e2e_value"] is None else {"value": rec["e2e_value"], "unit": UNIT,
                                                          "ms_per_step": rec["e2e_ms"] / rec["steps"],
                                                          "h2d_bytes_per_step": rec["h2d"], "d2h_bytes_per_step": rec["d2h"]},
            "roofline": rf}


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = args.workload or default_workload(world)
    rec = measure(name, rank, world, local, args.steps, args.warmup)
    out = None
    if rank == 0:
        peaks = load_peaks()
        w = wl.WORKLOADS[name]
        out = {
            "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": rec["ms"] / args.steps, "higher_is_better": True,
            "scaling": "strong" if rec["strong"] else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": rec["desc"], "name": name, "n_taxa": rec["n_taxa"], "sites_total": rec["n_sites_total"],
                       "patterns_total": int(rec["P_total"]), "patterns_rank0": rec["P"], "ns": rec["ns"], "ncatg": rec["ncatg"],
                       "updates_per_eval": rec["updates_per_eval"], "parallelism": f"site-shard x{world}",
                       "column_blocks_per_rank": len(rec["blocks"]), "both_sides": False,
                       "exchange": ("none" if world == 1 else os.environ.get("PLK_COMM", "p2p")),
                       "model": "the reference's eigen system for `%s` (tests/golden/big/%s.npz)" % (" ".join(wl.REF_ARGS[w.ns]), name),
                       "l2": "CLV working set %.2f GB per evaluation and GPU > 126 MB L2 (no flush needed)"
                             % ((rec["n_taxa"] - 2) * rec["P"] * rec["ns"] * rec["ncatg"] * 8 / 1e9)},
            "evals_per_s": args.steps / (rec["ms"] * 1e-3), "lnL": rec["lnL"], "lnL_reference": rec["lnL_reference"],
            "parity_rel_err": rec["parity_rel_err"], "parity_rtol": PARITY_RTOL,
            "exchange_rel_err": rec["exchange_rel_err"],
            "e2e": {"value": rec["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": rec["h2d"], "d2h_bytes_per_step": rec["d2h"],
                    "ms_per_step": rec["e2e_ms"] / args.steps,
                    "pipelined": rec["e2e_pipelined"],
                    "note": "every step copies its inputs H2D and reads lnL + per-site lnL D2H; pipelined = step s+1's "
                            "upload is issued while step s runs (plk_lk_full_begin / plk_lk_wait), PLK_BENCH_E2E_SERIAL=1 "
                            "disables it"},
            "gpu_launches": rec["launches"],
            "clocks": rec["clocks"],
            "roofline": roofline_of(rec, peaks),
        }
        if world == 1 and not args.no_traffic:
            pass  # filled in below, after the process group / timing legs are done
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and world == 1:
        if not args.no_traffic:
            tr, src = measured_traffic(name)
            if tr is not None:
                out["roofline"]["traffic"], out["roofline"]["traffic_source"] = tr, src
            else:
                out["roofline"]["traffic_note"] = src + "; value from the committed capture instead"
        if not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(name, cores=1, sites=args.cpu_sites, evals=args.cpu_evals, gpu_check=local)
        if not args.no_secondary and name == "dna_100x100k":
            sec = {}
            for nm, st in (("aa_200x50k", min(args.steps, 10)), ("dna_500x1M", min(args.steps, 5))):
                try:
                    r2 = measure(nm, 0, 1, local, st, 3, e2e=(nm == "aa_200x50k"), clocks=False, procs=16)
                    sec[nm] = secondary_of(r2, peaks)
                except Exception as ex:  # a secondary workload must never take the headline line down
                    sec[nm] = {"error": f"{type(ex).__name__}: {ex}"}
            try:
                sec["spr_path_dna_100x50k"] = spr_path_latency(local)
            except Exception as ex:
                sec["spr_path_dna_100x50k"] = {"error": f"{type(ex).__name__}: {ex}"}
            out["secondary"] = sec
    if rank == 0:
        print(json.dumps(out))


def spr_path_latency(local, name="dna_100x50k"):
    """The calls that bound an SPR search (BASELINE configs[4] shape, 100 taxa x 50 000 sites), through the C ABI:
    one regraft candidate as the reference issues it (P-matrices + Update_Partial_Lk at the new node + Lk(b_arrow),
    spr.c:589-650) against the same candidates scored in one call (plk_spr_candidates), one dLk, and Pars(NULL) with both
    sides (pars.c:20) -- host wall-clock per call, results cross-checked between the two candidate paths."""
    from phyml_b200.engine import Engine, pack_ops
    from phyml_b200.lk import LkTree
    from phyml_b200.tree import PartialOp, Side

    w = wl.WORKLOADS[name]
    m, _pin = wl.evaluation_model(name)
    tree = wl.make_tree(w)
    pat = wl.make_patterns(name, [0], procs=8)
    eng = Engine(tree.n_otu, pat.n_pattern, m.ns, m.ncatg, tree.n_clv_handles + 1, tree.n_edges + 3, device=local)
    t = LkTree(tree, pat, m, eng)
    t.Set_Both_Sides(1)
    t.Lk()

    def timeit(fn, n, warm=5):
        for _ in range(warm):
            fn()
        eng.sync()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        eng.sync()
        return (time.perf_counter() - t0) / n * 1e6

    tip = 3
    te = tree.adj[tip][0][0]
    prune, l_prune = tree.side_of(te, tip), float(tree.l[te])
    cands = []
    for e in range(tree.n_edges):
        if e != te:
            a, b = tree.edge_sides(e)
            cands.append((a, 0.5 * tree.l[e], b, 0.5 * tree.l[e]))
    packed = eng.pack_spr_cands(cands)
    tmp, ha, hb, hp = tree.n_clv_handles, tree.n_edges, tree.n_edges + 1, tree.n_edges + 2
    hs = np.array([ha, hb, hp], dtype=np.int32)

    def one(c):
        a, la, b, lb = c
        eng.update_pmats(hs, np.array([la, lb, l_prune]))
        eng.update_partials(pack_ops([PartialOp(dst=tmp, c1=a, pmat1=ha, c2=b, pmat2=hb)]))
        return eng.edge_lnl(Side(clv=tmp), prune, hp)

    seq = np.array([one(c) for c in cands])
    got, _ = eng.spr_candidates(prune, l_prune, True, packed)
    rel = float(np.max(np.abs(got - seq) / np.abs(seq)))
    if rel > 1e-12:
        raise RuntimeError(f"batched candidate scores differ from the call sequence: {rel:.3e}")
    l0 = eng.launch_count
    t_batch = timeit(lambda: eng.spr_candidates(prune, l_prune, True, packed), n=20)
    launches_batch = (eng.launch_count - l0) // 25
    t_seq = timeit(lambda: one(cands[7]), n=200)
    e = tree.n_edges // 2
    left, rght = tree.edge_sides(e)
    eng.eigen_lr(left, rght)
    t_dlk = timeit(lambda: eng.lnl_dlnl(0.05), n=200)
    # parsimony: tips = the bit masks of the tip table, Pars(NULL) with both sides in one call
    table = pat.table()
    masks = (table.astype(np.int64) << np.arange(m.ns)[None, :]).sum(axis=1).astype(np.int32)
    eng.pars_create(tree.n_clv_handles + 1)
    zero = np.zeros(pat.n_pattern, dtype=np.int32)
    for i in range(tree.n_otu):
        eng.pars_set_buffer(tree.pars_tip_handle(i), ui=masks[pat.codes[i]], pars=zero)
    a0, d0 = 0, tree.adj[0][0][1]
    e0 = tree.adj[0][0][0]
    pops = np.asarray(tree.pars_ops(tree.post_order_ops(a0, d0)) + tree.pars_ops(tree.pre_order_ops(a0, d0)), dtype=np.int32)
    c_pars = eng.pars_traverse_edge(pops, 2 * e0, 2 * e0 + 1)
    if c_pars != eng.pars_edge(2 * (tree.n_edges // 2), 2 * (tree.n_edges // 2) + 1):
        raise RuntimeError("parsimony score depends on the edge it is summed at")
    t_pars = timeit(lambda: eng.pars_traverse_edge(pops, 2 * e0, 2 * e0 + 1), n=100)
    res = {"workload": w.desc, "patterns": pat.n_pattern, "unit": "us per call (host wall-clock through the C ABI)",
           "candidate_as_the_reference_issues_it": round(t_seq, 2),
           "candidate_batched": round(t_batch / len(cands), 2), "candidates_per_batch": len(cands),
           "batch_launches": int(launches_batch), "batched_vs_sequence_rel_err": rel,
           "dLk": round(t_dlk, 2), "pars_null_both_sides": round(t_pars, 2), "pars_updates": int(len(pops)),
           "c_pars": int(c_pars)}
    eng.close()
    return res


# ======================================================================================================
def cpu_baseline(name, cores, sites, evals, warm=1, gpu_check=None, seed_block=0):
    """Times the reference's own Lk(NULL) (oracle/_ref/ref_driver = the unmodified reference built
    from /root/reference sources, AVX2+FMA path) on `cores` processes, each on its own slice of
    columns of the workload's first block(s) (the reference is single-threaded).  Falls back to the C
    oracle port if _ref is absent.  With `gpu_check` (a device index) the engine evaluates the SAME
    column sample and the relative difference of the two lnL is reported."""
    w = wl.WORKLOADS[name]
    tree = wl.make_tree(w)
    n = tree.n_otu
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    codes = wl.block_codes(w, seed_block)
    while codes.shape[1] < sites and w.n_blocks > 1 and seed_block + 1 < w.n_blocks:
        seed_block += 1
        codes = np.concatenate([codes, wl.block_codes(w, seed_block)], axis=1)
    sites = min(sites, codes.shape[1])
    per = sites // cores
    if os.path.exists(driver):
        with tempfile.TemporaryDirectory() as wd:
            with open(os.path.join(wd, "tree.nwk"), "w") as f:
                f.write(tree.to_newick() + "\n")
            procs = []
            for c in range(cores):
                phy = os.path.join(wd, f"aln{c}.phy")
                alignment.write_phylip(phy, codes[:, c * per:(c + 1) * per], w.ns, tree.names)
                cmd = [driver, "--time", str(evals), "--warmup", str(warm), "--", "-i", phy, "-u",
                       os.path.join(wd, "tree.nwk")] + wl.REF_ARGS[w.ns] + ["-o", "n", "-b", "0", "--r_seed", "1", "--no_memory_check"]
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
        out = {"value": pats * (n - 2) / slowest, "unit": UNIT, "cores": cores, "kind": "reference",
               "sample": f"{n} taxa x {per * cores} sites ({pats} patterns; the first columns of the benchmarked alignment), "
                         f"{evals} timed Lk(NULL) per process after {warm} warm-up, {cores} process(es) x 1 thread, "
                         f"gcc -O3 -march=haswell (AVX2+FMA kernels)",
               "s_per_eval": slowest, "lnL_sample": [r["lnL"] for r in res]}
        if gpu_check is not None and cores == 1:
            from phyml_b200.engine import Engine
            from phyml_b200.lk import LkTree

            m, _ = wl.evaluation_model(name)
            pat = alignment.compress(codes[:, :per], w.ns)
            t = LkTree(tree, pat, m, Engine(n, pat.n_pattern, w.ns, m.ncatg, tree.n_clv_handles, tree.n_edges, device=gpu_check))
            g = t.Lk()
            t.eng.close()
            out["lnL_sample_gpu"] = g
            out["sample_parity_rel_err"] = abs(g - res[0]["lnL"]) / abs(res[0]["lnL"])
        return out
    # port: C oracle through the test backend
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_backend import OracleBackend
    from phyml_b200.lk import LkTree

    m, _ = wl.evaluation_model(name)
    sites = min(sites, 5000)
    pat = alignment.compress(codes[:, :sites], w.ns)
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
    name = args.workload or default_workload(args.gpus)
    w = wl.WORKLOADS[name]
    cores = os.cpu_count() or 1
    # bounded sample: every host core gets its own block of columns of the workload's shape (sites are
    # independent, the reference is single-threaded: one process per core).  1 000 columns per core keeps a
    # process's likelihood arena near 38 MB: measured on the 128-core GPU-box host this is the reference's
    # best case (2.2e8 updates/s; 3 000 columns per core drop to 1.6e8, memory-bandwidth bound)
    per_core = 1000 if w.ns == 4 else 200
    budget = int(1.5e7 * 90 / max(1, (args.steps + args.warmup)) / (w.n_taxa - 2))   # ~90 s at 1.5e7 updates/s/core
    total = w.block_sites * w.n_blocks
    # when the WHOLE alignment of the workload fits the time budget and the host's memory (the reference's likelihood
    # arena is (3n-2) x columns x ncatg x ns doubles per process), every core takes its share of ALL columns: the same
    # alignment as the B200 arm, and the slices' lnL add up to the alignment's lnL
    whole = total // cores if cores <= total else 0
    arena = (3 * w.n_taxa - 2) * whole * 4 * w.ns * 8 * cores
    try:
        ram = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
    except (ValueError, OSError):
        ram = 64 << 30
    same_alignment = bool(w.n_blocks == 1 and whole >= 50 and whole <= budget and arena < 0.4 * ram
                          and os.environ.get("PLK_REF_SAMPLE") is None)
    if same_alignment:
        per_core = whole
    else:
        per_core = max(100, min(per_core, budget))
        if per_core * cores > total:
            per_core = max(50, total // cores)
            cores = max(1, min(cores, total // per_core))
    cb = cpu_baseline(name, cores=cores, sites=per_core * cores, evals=args.steps, warm=args.warmup)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["s_per_eval"] * 1e3,
           "higher_is_better": True, "scaling": "strong" if w.n_blocks > 1 else "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": w.desc, "name": name, "n_taxa": w.n_taxa, "ns": w.ns, "ncatg": 4,
                      "note": "reference CPU implementation (host cores); each step = one Lk(NULL) on a bounded column sample "
                              "of the workload's alignment (its first columns, one slice per core)"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if same_alignment:
        out["config"]["note"] = ("reference CPU implementation (host cores); each step = one Lk(NULL) over the WHOLE alignment of "
                                 "the workload, its columns split over one single-threaded process per core")
        out["same_alignment"] = True
        out["sites_evaluated"] = per_core * cores
        out["evals_per_s"] = 1.0 / cb["s_per_eval"]
        out["lnL"] = float(sum(cb.get("lnL_sample", [])))
        pin = wl.load_pin(name)
        if pin is not None and w.n_blocks == 1 and per_core * cores == total:
            out["lnL_reference_single_process"] = float(pin["lnL"])
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(wl.WORKLOADS),
                    help="default: dna_100x100k on 1 GPU (BASELINE configs[1]), dna_500x1M on N>1 (configs[3], strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-traffic", action="store_true", help="skip the in-run ncu pass that measures roofline.traffic")
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
