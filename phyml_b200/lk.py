"""Host-side mirror of the reference's likelihood interface (src/lk.h) over the device engine.

Function names, argument meaning and flag semantics follow the reference so parity tests read
like calls into src/lk.c:

    Lk(b=None)                       src/lk.c:443     full traversal (b is None) or edge-only
    dLk(l, b)                        src/lk.c:655     lnL and d lnL/dl at trial length l
    Update_Partial_Lk(b, d)          src/lk.c:1282
    Update_PMat_At_Given_Edge(b)     src/lk.c:2238
    Update_Eigen_Lr(b)               src/lk.c:1038
    Post_Order_Lk / Pre_Order_Lk     src/lk.c:282 / :357
    Set_Both_Sides, Set_Use_Eigen_Lr, Set_Update_Eigen_Lr   src/utilities.c:11614-11632
    Check_Lk_At_Given_Edge           src/lk.c:2642 (pulley-principle self test)
    Br_Len_Opt                       src/optimiz.c:607 (driver of the dLk kernel)

The arithmetic happens in the engine passed in (phyml_b200.engine.Engine = the CUDA library;
tests substitute the CPU oracle through the same interface).  CLV updates are queued and flushed
in dependency order as one batched call when a scalar is needed, which is how the per-node
``Update_Partial_Lk`` calls of the reference's host recursion become a few large launches.

Site sharding: each rank owns a contiguous block of patterns; the only
exchange is the sum of the per-rank partial lnL (and d lnL) -- ``reduce_fn``.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from .alignment import Patterns
from .model import Model
from .tree import PartialOp, Tree

YES, NO = 1, 0


class LkTree:
    def __init__(self, tree: Tree, patterns: Patterns, model: Model, engine,
                 reduce_fn: Optional[Callable[[Sequence[float]], Sequence[float]]] = None):
        if patterns.n_taxa != tree.n_otu:
            raise ValueError("alignment / tree taxon count mismatch")
        self.tree, self.data, self.mod, self.eng = tree, patterns, model, engine
        self.reduce_fn = reduce_fn
        self.both_sides = NO
        self.use_eigen_lr = NO
        self.update_eigen_lr = NO
        self.c_lnL = 0.0
        self.old_lnL = 0.0
        self.c_dlnL = 0.0
        self.numerical_warning = NO
        self._queue: List[PartialOp] = []
        self._eigen_edge = -1
        self.n_flush = 0
        engine.set_weights(patterns.wght, patterns.invar)
        engine.set_tip_table(patterns.table())
        for i in range(tree.n_otu):
            engine.set_tip_codes(i, patterns.codes[i])
        engine.set_model(model)

    # ---------------------------------------------------------------- flags (utilities.c:11614+)
    def Set_Both_Sides(self, yesno):
        self.both_sides = YES if yesno else NO
        self.tree.both_sides = bool(yesno)

    def Set_Use_Eigen_Lr(self, yesno):
        self.use_eigen_lr = YES if yesno else NO

    def Set_Update_Eigen_Lr(self, yesno):
        self.update_eigen_lr = YES if yesno else NO

    def Set_Model(self, model: Model):
        """Model parameters changed (the reference re-runs Update_RAS/Efrq/Eigen in Lk(NULL))."""
        self.mod = model
        self.eng.set_model(model)

    # ---------------------------------------------------------------- K0
    def Update_PMat_At_Given_Edge(self, b: int):
        self._flush()
        self.eng.update_pmats([b], [float(self.tree.l[b])])

    def Update_All_PMats(self):
        self.eng.update_pmats(list(range(self.tree.n_edges)), [float(x) for x in self.tree.l])

    # ---------------------------------------------------------------- K1
    def Update_Partial_Lk(self, b: int, d: int):
        if self.tree.is_tip(d):
            return  # lk.c:1297
        self._queue.append(self.tree.partial_op(b, d))

    def Post_Order_Lk(self, a: int, d: int):
        self._queue.extend(self.tree.post_order_ops(a, d))

    def Pre_Order_Lk(self, a: int, d: int):
        self._queue.extend(self.tree.pre_order_ops(a, d))

    def Update_All_Partial_Lk(self):
        a = self.tree.tip_root
        d = self.tree.adj[a][0][1]
        self.Post_Order_Lk(a, d)
        if self.both_sides:
            self.Pre_Order_Lk(a, d)

    def _flush(self):
        if self._queue:
            self.eng.update_partials(self._queue)
            self._queue = []
            self.n_flush += 1

    # ---------------------------------------------------------------- reductions
    def _reduce(self, vals):
        if self.reduce_fn is None:
            return list(vals)
        return list(self.reduce_fn(list(vals)))

    # ---------------------------------------------------------------- K3
    def Update_Eigen_Lr(self, b: int):
        self._flush()
        left, rght = self.tree.edge_sides(b)
        self.eng.eigen_lr(left, rght)
        self._eigen_edge = b

    # ---------------------------------------------------------------- Lk (lk.c:443-649)
    def Lk(self, b: Optional[int] = None) -> float:
        self.numerical_warning = NO
        self.old_lnL = self.c_lnL
        if b is None:
            self.eng.set_model(self.mod)          # Update_RAS/Efrq/Eigen results (lk.c:489-495)
            self.Update_All_PMats()               # lk.c:500-505
            self.Update_All_Partial_Lk()          # lk.c:562-564
            b = self.tree.root_edge               # lk.c:578-579
            if not self.update_eigen_lr and not self.use_eigen_lr and self._queue:
                # traversal + site loop at the root edge as ONE engine call (plk_traverse_edge_lnl)
                ops, self._queue = self._queue, []
                self.n_flush += 1
                left, rght = self.tree.edge_sides(b)
                lnl = self.eng.traverse_edge_lnl(ops, left, rght, b)      # lk.c:562-645
                (self.c_lnL,) = self._reduce([lnl])
                return self.c_lnL
        elif self.use_eigen_lr == NO:
            self.Update_PMat_At_Given_Edge(b)     # lk.c:515-527
        self._flush()
        if self.update_eigen_lr:
            self.Update_Eigen_Lr(b)               # lk.c:590
        if self.use_eigen_lr:
            lnl = self.eng.lnl_eigen(float(self.tree.l[b]))   # lk.c:592-603, 625-629
        else:
            left, rght = self.tree.edge_sides(b)
            lnl = self.eng.edge_lnl(left, rght, b)            # lk.c:605-645
        (self.c_lnL,) = self._reduce([lnl])
        return self.c_lnL

    # ---------------------------------------------------------------- dLk (lk.c:655-753)
    def dLk(self, l: float, b: int):
        """Returns (clamped l, lnL); sets c_lnL and c_dlnL like the reference."""
        self.numerical_warning = NO
        if self.update_eigen_lr:
            self.Update_Eigen_Lr(b)
        lc, lnl, dlnl = self.eng.lnl_dlnl(float(l))
        self.c_lnL, self.c_dlnL = self._reduce([lnl, dlnl])
        return lc, self.c_lnL

    # ---------------------------------------------------------------- Check_Lk_At_Given_Edge (lk.c:2642)
    def Check_Lk_At_Given_Edge(self, tol: float = 1e-2) -> np.ndarray:
        assert self.both_sides, "needs both_sides == YES"
        vals = np.array([self.Lk(e) for e in range(self.tree.n_edges)])
        if np.abs(vals - vals[0]).max() > tol:
            raise AssertionError("lnL differs across edges: pulley principle violated")
        return vals

    # ---------------------------------------------------------------- Br_Len_Opt (optimiz.c:607-664)
    def Br_Len_Opt(self, b: int, tol: float = 1e-6, max_iter: int = 50) -> float:
        """Optimise one branch length on the eigen-basis kernels, as optimiz.c:Br_Len_Opt does:
        one Lk(b) that projects both CLVs (K3), then repeated dLk (K4) inside a safeguarded
        Newton / bisection on d lnL/dl = 0 (the reference uses a spline search, Br_Len_Spline
        optimiz.c:2244; the engine calls are the same)."""
        self.Set_Update_Eigen_Lr(YES)
        self.Set_Use_Eigen_Lr(NO)
        self.Lk(b)
        self.Set_Update_Eigen_Lr(NO)
        self.Set_Use_Eigen_Lr(YES)
        lo, hi = self.mod.l_min, self.mod.l_max
        l = min(max(float(self.tree.l[b]), lo), hi)
        best_l, best_lnl = l, -np.inf
        for _ in range(max_iter):
            l, lnl = self.dLk(l, b)
            if lnl > best_lnl:
                best_l, best_lnl = l, lnl
            g = self.c_dlnL
            if g > 0:
                lo = l
            else:
                hi = l
            if hi - lo < tol * max(l, 1e-8):
                break
            # secant-free safeguarded step: geometric bisection of the bracket
            nl = np.sqrt(lo * hi) if lo > 0 else 0.5 * (lo + hi)
            if not (lo < nl < hi):
                nl = 0.5 * (lo + hi)
            l = nl
        self.tree.l[b] = best_l
        self.Set_Use_Eigen_Lr(NO)
        self.Update_PMat_At_Given_Edge(b)
        self.c_lnL = best_lnl
        return best_lnl
