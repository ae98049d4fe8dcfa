"""Substitution-model host maths feeding the device engine.

In PhyML these are ns x ns host computations (src/models.c: Update_Eigen :881, Update_RAS :669,
Update_Efrq :766; src/eigen.c; src/stats.c:DiscreteGamma :1974) that stay on the host
(SURVEY.md section 1): only their results -- U = right eigenvectors, V = U^-1, lambda, pi, the
category rates r_c and weights w_c -- are uploaded (``plk_set_model``).  This module provides the
same quantities for the stand-alone harness (tests, bench) and carries models dumped from the
reference (tests/golden/*.npz) unchanged.
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Sequence

import numpy as np

DNA = "ACGT"
AA = "ARNDCQEGHILKMFPSTWYV"


@dataclasses.dataclass
class Model:
    """Everything ``plk_set_model`` uploads ("Model upload set")."""

    ns: int
    U: np.ndarray          # [ns, ns] right eigenvectors, eigen->r_e_vect
    V: np.ndarray          # [ns, ns] inverse,            eigen->l_e_vect
    lam: np.ndarray        # [ns]     eigenvalues,        eigen->e_val
    pi: np.ndarray         # [ns]     e_frq->pi->v
    rates: np.ndarray      # [ncatg]  ras->gamma_rr->v
    rate_probs: np.ndarray  # [ncatg]  ras->gamma_r_proba->v
    pinv: float = 0.0      # ras->pinvar->v
    invar: bool = False    # ras->invar
    l_min: float = 1e-8    # mod->l_min  (src/init.c defaults)
    l_max: float = 100.0   # mod->l_max
    br_len_mult: float = 1.0
    name: str = ""

    @property
    def ncatg(self) -> int:
        return int(self.rates.shape[0])

    def pmat(self, l: float) -> np.ndarray:
        """Host restatement of Update_PMat_At_Given_Edge + PMat_Empirical (numpy; for simulation
        only -- the product computes P on the device, K0)."""
        out = np.empty((self.ncatg, self.ns, self.ns))
        for c in range(self.ncatg):
            ln = min(max(max(0.0, l) * self.rates[c] * self.br_len_mult, self.l_min), self.l_max)
            P = (self.U * np.exp(self.lam * ln)) @ self.V
            P = np.maximum(P, 1e-100)
            out[c] = P / P.sum(axis=1, keepdims=True)
        return out


def discrete_gamma(alpha: float, ncatg: int):
    """Mean-of-category discrete Gamma(alpha, alpha) rates, equal weights (Yang 1994), as
    DiscreteGamma(..., median=0) in src/stats.c:1974."""
    from scipy.special import gammainc, gammaincinv

    if ncatg == 1:
        return np.ones(1), np.ones(1)
    bounds = gammaincinv(alpha, np.arange(1, ncatg) / ncatg) / alpha
    upper = np.concatenate([gammainc(alpha + 1.0, bounds * alpha), [1.0]])
    lower = np.concatenate([[0.0], upper[:-1]])
    rates = (upper - lower) * ncatg
    rates = rates / rates.mean()
    return rates, np.full(ncatg, 1.0 / ncatg)


def reversible_eigen(S: np.ndarray, pi: np.ndarray):
    """Eigen system of the reversible rate matrix Q_ij = S_ij pi_j (i != j), normalised to one
    expected substitution per unit time (-sum_i pi_i Q_ii = 1), via the symmetric similarity
    transform (what src/models.c:Update_Eigen obtains with the general solver of src/eigen.c)."""
    ns = len(pi)
    S = np.array(S, dtype=np.float64)
    S = 0.5 * (S + S.T)
    np.fill_diagonal(S, 0.0)
    Q = S * pi[None, :]
    np.fill_diagonal(Q, -Q.sum(axis=1))
    mu = -float(np.dot(pi, np.diag(Q)))
    Q /= mu
    sq = np.sqrt(pi)
    B = (sq[:, None] * Q) / sq[None, :]
    B = 0.5 * (B + B.T)
    lam, W = np.linalg.eigh(B)
    U = W / sq[:, None]
    V = W.T * sq[None, :]
    # the stationary eigenvalue is exactly 0 in exact arithmetic
    lam[np.argmax(lam)] = 0.0
    assert np.allclose(U @ V, np.eye(ns), atol=1e-10)
    return U, V, lam, Q


def gtr(rr: Sequence[float] = (1.0, 2.5, 0.8, 1.2, 3.0, 1.0),
        pi: Sequence[float] = (0.30, 0.20, 0.25, 0.25), alpha: float = 0.5, ncatg: int = 4,
        pinv: float = 0.0) -> Model:
    """GTR + Gamma (+I). rr order = (AC, AG, AT, CG, CT, GT) as in BASELINE.md section 3."""
    pi = np.asarray(pi, dtype=np.float64)
    pi = pi / pi.sum()
    S = np.zeros((4, 4))
    idx = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    for (i, j), r in zip(idx, rr):
        S[i, j] = S[j, i] = r
    U, V, lam, _ = reversible_eigen(S, pi)
    rates, probs = discrete_gamma(alpha, ncatg)
    return Model(4, U, V, lam, pi, rates, probs, pinv=pinv, invar=pinv > 0.0, name="GTR+G%d" % ncatg)


def hky85(kappa: float = 4.0, pi: Sequence[float] = (0.25, 0.25, 0.25, 0.25), alpha: float = 1.0,
          ncatg: int = 4, pinv: float = 0.0) -> Model:
    m = gtr((1.0, kappa, 1.0, 1.0, kappa, 1.0), pi, alpha, ncatg, pinv)
    m.name = "HKY85+G%d" % ncatg
    return m


def from_exchangeabilities(S: np.ndarray, pi: np.ndarray, alpha: float = 0.5, ncatg: int = 4,
                           pinv: float = 0.0, name: str = "") -> Model:
    pi = np.asarray(pi, dtype=np.float64)
    pi = pi / pi.sum()
    U, V, lam, _ = reversible_eigen(np.asarray(S, dtype=np.float64), pi)
    rates, probs = discrete_gamma(alpha, ncatg)
    return Model(len(pi), U, V, lam, pi, rates, probs, pinv=pinv, invar=pinv > 0.0, name=name)


def from_golden(g, name: str = "") -> Model:
    """Model exactly as the reference computed it (arrays dumped by oracle/ref_driver.c)."""
    ns = int(g["ns"])
    return Model(ns, np.array(g["U"]).reshape(ns, ns), np.array(g["V"]).reshape(ns, ns),
                 np.array(g["lambda"]), np.array(g["pi"]), np.array(g["rates"]),
                 np.array(g["rate_probs"]), pinv=float(g["pinvar"]), invar=bool(int(g["invar_flag"])),
                 l_min=float(g["l_min"]), l_max=float(g["l_max"]), br_len_mult=float(g["br_len_mult"]),
                 name=name)


def with_gamma(m: Model, alpha: float, ncatg: int) -> Model:
    rates, probs = discrete_gamma(alpha, ncatg)
    return dataclasses.replace(m, rates=rates, rate_probs=probs)


def synthetic_aa(seed: int = 7, alpha: float = 0.5, ncatg: int = 4) -> Model:
    """A seeded random reversible 20-state model (fallback when the LG fixture is unavailable)."""
    rng = np.random.default_rng(seed)
    S = rng.gamma(0.5, 2.0, size=(20, 20)) + 0.01
    pi = rng.dirichlet(np.full(20, 8.0))
    return from_exchangeabilities(S, pi, alpha, ncatg, name="RAND20+G%d" % ncatg)


def lg_from_fixture(path: Optional[str] = None, alpha: float = 0.5, ncatg: int = 4) -> Model:
    """LG (Le & Gascuel 2008) eigen system as dumped from the reference (init.c:4026 ->
    Update_Eigen) into tests/golden/lg_model.npz by tests/golden/make_golden.py."""
    import os

    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests",
                            "golden", "lg_model.npz")
    g = np.load(path)
    rates, probs = discrete_gamma(alpha, ncatg)
    return Model(20, g["U"].reshape(20, 20), g["V"].reshape(20, 20), g["lambda"], g["pi"], rates,
                 probs, name="LG+G%d" % ncatg)
