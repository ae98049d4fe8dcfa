"""Seeded synthetic workloads of BASELINE.json (one definition shared by bench.py, the GPU parity tests
and tests/golden/make_golden_big.py, so that all three see the SAME alignment, tree and model).

An alignment is a sequence of column BLOCKS; block b is simulated down the workload's fixed random
tree with seed 1000 + b and pattern-compressed on its own (``Compact_Data``, src/utilities.c:215).
Site sharding assigns whole blocks to ranks, so the all-rank lnL of a
multi-block workload is the sum of its per-block lnL whatever the number of GPUs: that is what lets
a 1-, 2-, 4- or 8-GPU run of the 1M-site configuration be compared with the reference's own value
on the same alignment (the reference evaluates one block at a time in tests/golden/make_golden_big.py).
Block sizes are bounded by the reference itself: Make_Tree_For_Lk computes the size of its likelihood
arena, (3n-2) * n_pattern * ncatg * ns doubles, in 32-bit int arithmetic (src/make.c:96-104), so the
unmodified reference aborts ("Err. in file make.c (line 104)") on 200 taxa x >44 889 amino-acid patterns or
500 taxa x >89 600 DNA patterns -- BASELINE configs[2] and [3] can only be evaluated by it in column blocks.

Branch lengths are rounded to the 10 decimals the Newick file carries, so the reference (which reads
the tree from that file) and the engine evaluate exactly the same lengths.

The evaluation model of a pinned workload is the reference's own eigen system for the CLI arguments
in ``REF_ARGS`` (dumped by oracle/ref_driver --summary into tests/golden/big/<name>.npz): fixed
user frequencies for DNA (``-f 0.30,0.20,0.25,0.25``) and the LG frequencies for amino acids, so
the model does not depend on the block.  Simulation always uses the generator model below.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional

import numpy as np

from . import alignment, model as pmodel
from .tree import Tree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIG_GOLDEN_DIR = os.path.join(ROOT, "tests", "golden", "big")


@dataclasses.dataclass(frozen=True)
class Workload:
    name: str
    ns: int
    n_taxa: int
    block_sites: int
    n_blocks: int          # blocks of the whole alignment (weak-scaling workloads: one per rank)
    desc: str
    alias_of: Optional[str] = None   # block 0 of another workload


WORKLOADS = {w.name: w for w in [
    Workload("dna_100x100k", 4, 100, 100_000, 1,
             "synthetic DNA 100 taxa x 100000 sites, GTR+G4, fixed random tree (BASELINE configs[1])"),
    Workload("aa_200x50k", 20, 200, 25_000, 2,
             "synthetic AA 200 taxa x 50000 sites (2 column blocks of 25000), LG+G4 (BASELINE configs[2])"),
    Workload("dna_500x1M", 4, 500, 62_500, 16,
             "synthetic DNA 500 taxa x 1000000 sites (16 column blocks of 62500), GTR+G4, site-sharded (BASELINE configs[3])"),
    Workload("dna_500x62k", 4, 500, 62_500, 1,
             "block 0 (62500 columns) of the 500-taxon 1M-site alignment (1/16 of BASELINE configs[3])",
             alias_of="dna_500x1M"),
    Workload("dna_100x50k", 4, 100, 50_000, 1,
             "synthetic DNA 100 taxa x 50000 sites, GTR+G4 (alignment of BASELINE configs[4])"),
    Workload("dna_16x4k", 4, 16, 4_096, 1, "tiny DNA smoke workload"),
]}

# reference CLI of a pinned workload (model part); tests/golden/make_golden_big.py adds -i/-u/-o n ...
REF_ARGS = {
    4: ["-d", "nt", "-m", "GTR", "-c", "4", "-a", "0.5", "-f", "0.30,0.20,0.25,0.25"],
    20: ["-d", "aa", "-m", "LG", "-c", "4", "-a", "0.5", "-f", "m"],
}


def generator_model(ns: int):
    return pmodel.gtr(alpha=0.5) if ns == 4 else pmodel.lg_from_fixture(alpha=0.5)


def make_tree(w: Workload) -> Tree:
    tree = Tree.random(w.n_taxa, seed=1)
    tree.l = np.array([float(f"{x:.10f}") for x in tree.l])
    return tree


def block_codes(w: Workload, block: int) -> np.ndarray:
    """Raw columns [n_taxa, block_sites] of block `block` (seed 1000 + block)."""
    tree = make_tree(w)
    if w.alias_of:
        src = WORKLOADS[w.alias_of]
        assert src.block_sites == w.block_sites and block == 0
        return block_codes(src, 0)
    return alignment.simulate(tree, generator_model(w.ns), w.block_sites, seed=1000 + block)


def _block_patterns(args):
    name, block = args
    w = WORKLOADS[name]
    return alignment.compress(block_codes(w, block), w.ns)


def concat_patterns(parts: List[alignment.Patterns]) -> alignment.Patterns:
    if len(parts) == 1:
        return parts[0]
    return alignment.Patterns(parts[0].ns, np.ascontiguousarray(np.concatenate([p.codes for p in parts], axis=1)),
                              np.concatenate([p.wght for p in parts]), np.concatenate([p.invar for p in parts]),
                              parts[0].names, sum(p.n_sites for p in parts))


def rank_blocks(w: Workload, rank: int, world: int, weak: bool = False) -> List[int]:
    """Blocks owned by `rank`.  Strong scaling: the workload's n_blocks split evenly over the ranks.
    Weak scaling (single-block workloads run on N GPUs): rank r evaluates its own block r."""
    if weak or w.n_blocks == 1:
        return [rank]
    if w.n_blocks % world:
        raise ValueError(f"{w.name}: {w.n_blocks} blocks do not split over {world} ranks")
    per = w.n_blocks // world
    return list(range(rank * per, (rank + 1) * per))


def make_patterns(name: str, blocks: List[int], procs: int = 1) -> alignment.Patterns:
    """Per-block compressed patterns of `blocks`, concatenated in block order."""
    jobs = [(name, b) for b in blocks]
    if procs > 1 and len(jobs) > 1:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(min(procs, len(jobs))) as pool:
            parts = pool.map(_block_patterns, jobs)
    else:
        parts = [_block_patterns(j) for j in jobs]
    return concat_patterns(parts)


def golden_path(name: str) -> str:
    return os.path.join(BIG_GOLDEN_DIR, name + ".npz")


def load_pin(name: str):
    """The reference's numbers for this workload (tests/golden/big/<name>.npz) or None."""
    p = golden_path(name)
    if not os.path.exists(p):
        return None
    return dict(np.load(p))


def evaluation_model(name: str):
    """(model, pin): the reference's own eigen system when the workload is pinned, else the generator model."""
    w = WORKLOADS[name]
    pin = load_pin(name)
    if pin is None:
        return generator_model(w.ns), None
    m = pmodel.Model(ns=w.ns, U=pin["U"].reshape(w.ns, w.ns), V=pin["V"].reshape(w.ns, w.ns), lam=pin["lambda"],
                     pi=pin["pi"], rates=pin["rates"], rate_probs=pin["rate_probs"], pinv=float(pin["pinvar"]),
                     invar=bool(int(pin["invar_flag"])), l_min=float(pin["l_min"]), l_max=float(pin["l_max"]),
                     br_len_mult=float(pin["br_len_mult"]), name=name + " (reference eigen system)")
    return m, pin
