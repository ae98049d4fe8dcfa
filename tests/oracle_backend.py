"""Test-only backend: the CPU oracle (oracle/liboracle.so) behind the same interface as
phyml_b200.engine.Engine, so host logic (traversal scheduling, sharding) can be tested on CPU and
the CUDA engine can be checked against it.  TEST INFRASTRUCTURE: never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class _Side(C.Structure):
    _fields_ = [("clv", C.c_void_p), ("scale", C.c_void_p), ("tipvec", C.c_void_p),
                ("d_state", C.c_void_p), ("is_ambigu", C.c_void_p)]


_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "plk_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
        _lib = C.CDLL(so)
        _lib.plk_oracle_edge_lnl.restype = C.c_double
        _lib.plk_oracle_lnl_dlnl.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


class OracleBackend:
    """Same methods as phyml_b200.engine.Engine, computed by oracle/plk_oracle.c."""

    def __init__(self, n_tips, n_pattern, ns, ncatg, n_clv, n_pmat, **_):
        self.lib = oracle_lib()
        self.n_tips, self.P, self.ns, self.ncatg = n_tips, n_pattern, ns, ncatg
        self.clv = {}
        self.scale = {}
        self.pm = np.zeros((n_pmat, ncatg, ns, ns))
        self.tipvec = [None] * n_tips
        self.d_state = [None] * n_tips
        self.is_ambigu = [None] * n_tips
        self.table = None
        self.apply_scaling = 1
        self.dot_prod = np.zeros((n_pattern, ncatg, ns))
        self.fact = np.zeros(n_pattern, dtype=np.int32)
        self.site = {}
        self.numerical_warning = 0

    # ---- uploads
    def set_weights(self, wght, invar):
        self.wght = np.ascontiguousarray(wght, dtype=np.float64)
        self.invar = np.ascontiguousarray(invar, dtype=np.int16)

    def set_tip_table(self, table):
        self.table = np.ascontiguousarray(table, dtype=np.float64)

    def set_tip_codes(self, tip, codes):
        vec = np.ascontiguousarray(self.table[np.asarray(codes, dtype=np.int64)])
        self.set_tip_vectors(tip, vec)

    def set_tip_vectors(self, tip, vec, d_state=None, is_ambigu=None):
        vec = np.ascontiguousarray(vec, dtype=np.float64).reshape(self.P, self.ns)
        self.tipvec[tip] = vec
        cnt = (vec > 0).sum(axis=1)
        self.is_ambigu[tip] = np.ascontiguousarray(
            (cnt != 1).astype(np.int16) if is_ambigu is None else is_ambigu, dtype=np.int16)
        self.d_state[tip] = np.ascontiguousarray(
            vec.argmax(axis=1).astype(np.int16) if d_state is None else d_state, dtype=np.int16)

    def set_model(self, m):
        self.m = m
        self.U = np.ascontiguousarray(m.U, dtype=np.float64)
        self.V = np.ascontiguousarray(m.V, dtype=np.float64)
        self.lam = np.ascontiguousarray(m.lam, dtype=np.float64)
        self.pi = np.ascontiguousarray(m.pi, dtype=np.float64)
        self.rates = np.ascontiguousarray(m.rates, dtype=np.float64)
        self.probs = np.ascontiguousarray(m.rate_probs, dtype=np.float64)

    # ---- K0
    def update_pmats(self, handles, lengths):
        m = self.m
        for h, l in zip(handles, lengths):
            out = np.zeros((self.ncatg, self.ns, self.ns))
            self.lib.plk_oracle_pmat(self.ns, self.ncatg, C.c_double(float(l)), _dp(self.rates),
                                     C.c_double(m.br_len_mult), C.c_double(m.l_min), C.c_double(m.l_max),
                                     _dp(self.U), _dp(self.V), _dp(self.lam), _dp(out))
            self.pm[h] = out

    def set_pmat(self, h, P):
        self.pm[h] = np.asarray(P).reshape(self.ncatg, self.ns, self.ns)

    def get_pmat(self, h):
        return self.pm[h].copy()

    # ---- operands
    def _side(self, s, keep):
        sd = _Side()
        if s.is_tip:
            sd.tipvec, sd.d_state, sd.is_ambigu = _p(self.tipvec[s.tip]), _p(self.d_state[s.tip]), _p(self.is_ambigu[s.tip])
        else:
            sd.clv, sd.scale = _p(self.clv[s.clv]), _p(self.scale[s.clv])
        keep.append(sd)
        return C.byref(sd)

    # ---- K1
    def update_partials(self, ops):
        for op in ops:
            if op.dst not in self.clv:
                self.clv[op.dst] = np.zeros((self.P, self.ncatg, self.ns))
                self.scale[op.dst] = np.zeros(self.P, dtype=np.int32)
            keep = []
            p1 = np.ascontiguousarray(self.pm[op.pmat1])
            p2 = np.ascontiguousarray(self.pm[op.pmat2])
            self.lib.plk_oracle_update_partial(self.ns, self.ncatg, self.P, _dp(self.wght), self.apply_scaling,
                                               _dp(self.clv[op.dst]), _p(self.scale[op.dst]),
                                               self._side(op.c1, keep), _dp(p1), self._side(op.c2, keep), _dp(p2))

    # ---- K1 + K2 (plk_traverse_edge_lnl)
    def traverse_edge_lnl(self, ops, left, rght, pmat):
        self.update_partials(ops)
        return self.edge_lnl(left, rght, pmat)

    # ---- K2
    def edge_lnl(self, left, rght, pmat):
        keep = []
        P = self.P
        self.site = {"site_lnl": np.zeros(P), "site_lk": np.zeros(P), "site_lk_cat": np.zeros((P, self.ncatg)),
                     "fact_sum_scale": np.zeros(P, dtype=np.int32)}
        warn = C.c_int(0)
        pm = np.ascontiguousarray(self.pm[pmat])
        lnl = self.lib.plk_oracle_edge_lnl(self.ns, self.ncatg, P, _dp(self.wght), _p(self.invar),
                                           int(self.m.invar), C.c_double(self.m.pinv), _dp(self.pi), _dp(self.probs),
                                           self._side(left, keep), self._side(rght, keep), _dp(pm),
                                           _dp(self.site["site_lnl"]), _dp(self.site["site_lk"]),
                                           _dp(self.site["site_lk_cat"]), _p(self.site["fact_sum_scale"]),
                                           C.byref(warn))
        self.fact = self.site["fact_sum_scale"]
        self.numerical_warning = warn.value
        return lnl

    # ---- K3 / K4
    def eigen_lr(self, left, rght):
        keep = []
        self.lib.plk_oracle_eigen_lr(self.ns, self.ncatg, self.P, _dp(self.wght), _dp(self.U), _dp(self.V),
                                     _dp(self.pi), self._side(left, keep), self._side(rght, keep),
                                     _dp(self.dot_prod))
        ls = self.scale[left.clv] if not left.is_tip else 0
        rs = self.scale[rght.clv] if not rght.is_tip else 0
        self.fact = np.ascontiguousarray(np.zeros(self.P, dtype=np.int32) + ls + rs, dtype=np.int32)

    def _k4(self, l, deriv):
        m = self.m
        lc = C.c_double(float(l))
        d = C.c_double(0.0)
        warn = C.c_int(0)
        lnl = self.lib.plk_oracle_lnl_dlnl(self.ns, self.ncatg, self.P, _dp(self.wght), _p(self.invar), int(m.invar),
                                           C.c_double(m.pinv), _dp(self.pi), _dp(self.rates), _dp(self.probs),
                                           C.c_double(m.br_len_mult), C.c_double(m.l_min), C.c_double(m.l_max),
                                           _dp(self.lam), _dp(self.dot_prod), _p(self.fact), C.byref(lc), deriv,
                                           C.byref(d), C.byref(warn))
        return lc.value, lnl, d.value

    def lnl_dlnl(self, l):
        return self._k4(l, 1)

    def lnl_eigen(self, l):
        return self._k4(l, 0)[1]

    # ---- read-backs
    def get_clv(self, h):
        return self.clv[h].copy(), self.scale[h].copy()

    def set_clv(self, h, clv, scale):
        self.clv[h] = np.ascontiguousarray(clv, dtype=np.float64).reshape(self.P, self.ncatg, self.ns).copy()
        self.scale[h] = np.ascontiguousarray(scale, dtype=np.int32).copy()

    def get_site_lnl(self):
        return {k: v.copy() for k, v in self.site.items()}

    def get_dot_prod(self):
        return self.dot_prod.copy()

    # ---- batched SPR candidates: the sequential composition the batched call stands for (spr.c:589-650)
    def spr_candidates(self, prune, l_prune, link_on_left, cands):
        from phyml_b200.tree import PartialOp, Side
        keep_pm = self.pm
        self.pm = np.concatenate([keep_pm, np.zeros((3,) + keep_pm.shape[1:])])
        ha, hb, hp = len(keep_pm), len(keep_pm) + 1, len(keep_pm) + 2
        tmp = -12345
        lnl, warn = np.zeros(len(cands)), np.zeros(len(cands), dtype=np.int32)
        for i, (a, la, b, lb) in enumerate(cands):
            self.update_pmats([ha, hb, hp], [la, lb, l_prune])
            self.update_partials([PartialOp(dst=tmp, c1=a, pmat1=ha, c2=b, pmat2=hb)])
            x = Side(clv=tmp)
            lnl[i] = self.edge_lnl(x, prune, hp) if link_on_left else self.edge_lnl(prune, x, hp)
            warn[i] = self.numerical_warning
        del self.clv[tmp], self.scale[tmp]
        self.pm = keep_pm
        return lnl, warn

    # ---- parsimony (oracle restatement of src/pars.c)
    def pars_create(self, n_buffers, step_mat=None):
        self.step_mat = None if step_mat is None else np.ascontiguousarray(step_mat, dtype=np.int32)
        self.p_ui, self.p_pars, self.p_pp = {}, {}, {}
        self.site_pars = np.zeros(self.P, dtype=np.int32)

    def pars_set_buffer(self, buf, ui=None, pars=None, p_pars=None):
        if ui is not None:
            self.p_ui[buf] = np.ascontiguousarray(ui, dtype=np.int32).copy()
            self.p_pars[buf] = np.ascontiguousarray(pars, dtype=np.int32).copy()
        if p_pars is not None:
            self.p_pp[buf] = np.ascontiguousarray(p_pars, dtype=np.int32).reshape(self.P, self.ns).copy()

    def pars_get_buffer(self, buf, general=False):
        return self.p_pp[buf].copy() if general else (self.p_ui[buf].copy(), self.p_pars[buf].copy())

    def pars_update(self, ops, general=False):
        for dst, c1, c2 in np.asarray(ops, dtype=np.int64).reshape(-1, 3).tolist():
            if general:
                out = np.zeros((self.P, self.ns), dtype=np.int32)
                self.lib.plk_oracle_pars_update_general(self.ns, self.P, _p(self.step_mat), _p(out),
                                                        _p(self.p_pp[c1]), _p(self.p_pp[c2]))
                self.p_pp[dst] = out
            else:
                ui, pars = np.zeros(self.P, dtype=np.int32), np.zeros(self.P, dtype=np.int32)
                self.lib.plk_oracle_pars_update(self.P, _p(ui), _p(pars), _p(self.p_ui[c1]), _p(self.p_pars[c1]),
                                                _p(self.p_ui[c2]), _p(self.p_pars[c2]))
                self.p_ui[dst], self.p_pars[dst] = ui, pars

    def pars_edge(self, left, rght, general=False):
        g = lambda d, k: _p(d[k]) if k in d else None
        return int(self.lib.plk_oracle_pars_edge(int(general), self.ns, self.P, _dp(self.wght), _p(self.step_mat),
                                                 g(self.p_ui, left), g(self.p_pars, left), g(self.p_pp, left),
                                                 g(self.p_ui, rght), g(self.p_pars, rght), g(self.p_pp, rght),
                                                 _p(self.site_pars)))

    def pars_traverse_edge(self, ops, left, rght, general=False):
        self.pars_update(ops, general)
        return self.pars_edge(left, rght, general)

    def get_site_pars(self):
        return self.site_pars.copy()

    def sync(self):
        pass

    def close(self):
        pass
