"""Alignments as the likelihood engine sees them: site patterns, weights, 1-byte tip codes.

Reference counterparts: ``calign`` (src/utilities.h:1148-1175: ``wght[]``, ``invar[]``,
``n_pattern``), ``Compact_Data`` (src/utilities.c:215, site-pattern compression) and the tip
initialisation tables (src/lk.c:26-69 nucleotides, :122-161 amino acids).  The engine keeps one
byte per (taxon, pattern) and a per-instance table ``code -> 0/1 state vector`` instead of the
reference's fp64 tip vectors (``p_lk_tip_r``, 8*ns bytes per taxon and pattern).
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Sequence

import numpy as np

from .model import AA, DNA, Model
from .tree import Tree

# IUPAC nucleotide codes -> 4-bit mask over (A, C, G, T); src/lk.c:26-69
_NT_MASK = {
    "A": 1, "C": 2, "G": 4, "T": 8, "U": 8, "M": 3, "R": 5, "W": 9, "S": 6, "Y": 10, "K": 12,
    "B": 14, "D": 13, "H": 11, "V": 7, "N": 15, "X": 15, "?": 15, "O": 15, "-": 15,
}
_NT_CHARS = "?ACMGRSVTWYHKDBN"  # index = mask (0 is never produced)

# amino-acid codes: 0..19 = states (order of src/lk.c:129-149), 20 = all ones (X ? -);
# 'B' -> N and 'Z' -> Q as in src/lk.c:151-152
_AA_CODE = {c: i for i, c in enumerate(AA)}
_AA_CODE.update({"B": 2, "Z": 5, "X": 20, "?": 20, "-": 20})


def tip_table(ns: int) -> np.ndarray:
    """Default code -> 0/1 vector table uploaded with ``plk_set_tip_table``."""
    if ns == 4:
        t = np.zeros((16, 4))
        for m in range(16):
            for k in range(4):
                t[m, k] = 1.0 if (m >> k) & 1 else 0.0
        return t
    if ns == 20:
        t = np.zeros((21, 20))
        t[np.arange(20), np.arange(20)] = 1.0
        t[20, :] = 1.0
        return t
    raise ValueError("ns must be 4 or 20")


def encode(seqs: Sequence[str], ns: int) -> np.ndarray:
    """Characters -> 1-byte codes, shape [n_taxa, n_sites]."""
    lut = np.full(256, 255, dtype=np.uint8)
    table = _NT_MASK if ns == 4 else _AA_CODE
    for ch, code in table.items():
        lut[ord(ch)] = code
        lut[ord(ch.lower())] = code
    out = np.stack([lut[np.frombuffer(s.encode("ascii"), dtype=np.uint8)] for s in seqs])
    if (out == 255).any():
        raise ValueError("unknown character state in alignment")
    return out


def decode(codes: np.ndarray, ns: int) -> list:
    chars = np.frombuffer((_NT_CHARS if ns == 4 else AA + "X").encode("ascii"), dtype=np.uint8)
    return [chars[row].tobytes().decode("ascii") for row in codes]


@dataclasses.dataclass
class Patterns:
    """Compressed alignment: what ``plk_create``/``plk_set_tip_codes`` receive."""

    ns: int
    codes: np.ndarray       # uint8 [n_taxa, n_pattern]
    wght: np.ndarray        # float64 [n_pattern]   calign->wght
    invar: np.ndarray       # int16 [n_pattern]     calign->invar (state if constant site, else -1)
    names: list
    n_sites: int

    @property
    def n_taxa(self) -> int:
        return int(self.codes.shape[0])

    @property
    def n_pattern(self) -> int:
        return int(self.codes.shape[1])

    def table(self) -> np.ndarray:
        return tip_table(self.ns)

    def shard(self, rank: int, world: int) -> "Patterns":
        """Contiguous block of patterns owned by ``rank``."""
        lo, hi = shard_bounds(self.n_pattern, rank, world)
        return Patterns(self.ns, np.ascontiguousarray(self.codes[:, lo:hi]), self.wght[lo:hi].copy(),
                        self.invar[lo:hi].copy(), self.names, self.n_sites)


def shard_bounds(n_pattern: int, rank: int, world: int):
    base, rem = divmod(n_pattern, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def invariant_states(codes: np.ndarray, ns: int) -> np.ndarray:
    """calign->invar: the state shared by all taxa at a pattern if the column can be constant,
    else -1 (cf. Set_D_States / Check_Ambiguities / invar bookkeeping in Compact_Data)."""
    table = tip_table(ns) > 0
    compat = np.ones((codes.shape[1], ns), dtype=bool)
    for row in codes:
        compat &= table[row]
    inv = np.where(compat.any(axis=1), compat.argmax(axis=1), -1).astype(np.int16)
    return inv


def compress(codes: np.ndarray, ns: int, names: Optional[list] = None, collapse: bool = True) -> Patterns:
    """Site-pattern compression (Compact_Data, src/utilities.c:215): identical columns are merged
    and counted in ``wght``.  Patterns are ordered by first occurrence."""
    n_taxa, n_sites = codes.shape
    names = names or [f"t{i}" for i in range(n_taxa)]
    if collapse:
        cols = np.ascontiguousarray(codes.T)
        view = cols.view(np.dtype((np.void, cols.dtype.itemsize * n_taxa))).ravel()
        _, first, counts = np.unique(view, return_index=True, return_counts=True)
        order = np.argsort(first, kind="stable")
        first, counts = first[order], counts[order]
        pc = np.ascontiguousarray(codes[:, first])
        w = counts.astype(np.float64)
    else:
        pc, w = np.ascontiguousarray(codes), np.ones(n_sites)
    return Patterns(ns, pc, w, invariant_states(pc, ns), names, n_sites)


def simulate(tree: Tree, model: Model, n_sites: int, seed: int = 1, ambiguity: float = 0.0) -> np.ndarray:
    """Simulate ``n_sites`` columns down ``tree`` under ``model`` (one Gamma category per site).
    Returns codes [n_taxa, n_sites]; a fraction ``ambiguity`` of cells becomes fully ambiguous
    (and, for DNA, a further equal fraction two-fold ambiguous) to exercise the tip-mask path."""
    rng = np.random.default_rng(seed)
    ns = model.ns
    cat = rng.choice(model.ncatg, size=n_sites, p=model.rate_probs / model.rate_probs.sum())
    states = np.empty((tree.n_nodes, n_sites), dtype=np.int8)
    root = tree.n_otu
    states[root] = rng.choice(ns, size=n_sites, p=model.pi / model.pi.sum())
    u = None
    stack = [(root, -1)]
    while stack:
        node, parent = stack.pop()
        for (e, v) in tree.adj[node]:
            if v == parent:
                continue
            P = model.pmat(float(tree.l[e]))                    # [ncatg, ns, ns]
            cdf = np.cumsum(P, axis=2)
            cdf[:, :, -1] = 1.0
            u = rng.random(n_sites)
            rows = cdf[cat, states[node].astype(np.int64)]       # [n_sites, ns]
            states[v] = (u[:, None] > rows).sum(axis=1).astype(np.int8)
            stack.append((v, node))
    tips = states[: tree.n_otu].astype(np.int64)
    if ns == 4:
        codes = (1 << tips).astype(np.uint8)
    else:
        codes = tips.astype(np.uint8)
    if ambiguity > 0.0:
        r = rng.random(codes.shape)
        full = 15 if ns == 4 else 20
        codes = np.where(r < ambiguity, full, codes).astype(np.uint8)
        if ns == 4:
            other = (1 << rng.integers(0, 4, size=codes.shape)).astype(np.uint8)
            two = (r >= ambiguity) & (r < 2 * ambiguity)
            codes = np.where(two, codes | other, codes).astype(np.uint8)
    return codes


def write_phylip(path: str, codes: np.ndarray, ns: int, names: Sequence[str]) -> None:
    seqs = decode(codes, ns)
    with open(path, "w") as f:
        f.write(f"{codes.shape[0]} {codes.shape[1]}\n")
        for nm, s in zip(names, seqs):
            f.write(f"{nm}  {s}\n")


def read_phylip(path: str):
    """Minimal sequential/interleaved PHYLIP reader (names separated from data by blanks)."""
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    n, length = (int(x) for x in lines[0].split()[:2])
    names, seqs = [], []
    for ln in lines[1: n + 1]:
        parts = ln.split()
        names.append(parts[0])
        seqs.append("".join(parts[1:]))
    k = 0
    for ln in lines[n + 1:]:
        seqs[k % n] += "".join(ln.split())
        k += 1
    seqs = [s.upper() for s in seqs]
    if any(len(s) != length for s in seqs):
        raise ValueError("PHYLIP length mismatch")
    return names, seqs
