"""ctypes binding of the C-ABI engine library (include/phyml_b200.h -> libphyml_b200.so).

This is the only compute backend of the package.  There is no CPU fallback: if the library has not
been built (``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C phyml_b200/csrc``)
or no CUDA device is present, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Optional, Sequence

import numpy as np

from .tree import PartialOp, Side

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libphyml_b200.so")


class EngineError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("n_tips", C.c_int), ("n_patterns", C.c_int), ("ns", C.c_int), ("ncatg", C.c_int),
                ("n_clv", C.c_int), ("n_pmat", C.c_int), ("device", C.c_int), ("flags", C.c_int)]


class _Side(C.Structure):
    _fields_ = [("tip", C.c_int), ("clv", C.c_int)]


class _Op(C.Structure):
    _fields_ = [("dst", C.c_int), ("c1", _Side), ("pmat1", C.c_int), ("c2", _Side), ("pmat2", C.c_int)]


class _SprCand(C.Structure):
    _fields_ = [("a", _Side), ("l_a", C.c_double), ("b", _Side), ("l_b", C.c_double)]


OP_DTYPE = np.dtype([("dst", "<i4"), ("c1_tip", "<i4"), ("c1_clv", "<i4"), ("pmat1", "<i4"),
                     ("c2_tip", "<i4"), ("c2_clv", "<i4"), ("pmat2", "<i4")])
assert OP_DTYPE.itemsize == C.sizeof(_Op)

EXPORTS = [
    "plk_create", "plk_destroy", "plk_last_error", "plk_sync", "plk_set_pattern_weights",
    "plk_set_tip_table", "plk_set_tip_codes", "plk_set_all_tip_codes", "plk_set_all_tip_codes_packed4", "plk_set_tip_vectors", "plk_set_model", "plk_update_pmats",
    "plk_set_pmat", "plk_get_pmat", "plk_update_partials", "plk_edge_lnl", "plk_traverse_edge_lnl", "plk_lk_full", "plk_eigen_lr",
    "plk_edge_lnl_dlnl", "plk_edge_lnl_eigen", "plk_get_clv", "plk_set_clv", "plk_get_site_lnl",
    "plk_get_dot_prod", "plk_comm_unique_id", "plk_comm_init", "plk_comm_set_allreduce",
    "plk_comm_p2p_export", "plk_comm_p2p_init", "plk_create_sharded", "plk_n_shards",
    "plk_launch_count", "plk_device_bytes", "plk_stream", "plk_version",
    "plk_pars_create", "plk_pars_set_buffer", "plk_pars_get_buffer", "plk_pars_update", "plk_pars_edge",
    "plk_pars_traverse_edge", "plk_get_site_pars", "plk_spr_candidates", "plk_lk_full_begin", "plk_lk_wait",
]

_lib = None


def load_library() -> C.CDLL:
    """Load libphyml_b200.so; raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: build it with `make -C phyml_b200/csrc` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.plk_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
    lib.plk_create_sharded.argtypes = [C.POINTER(_Config), C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
    lib.plk_n_shards.argtypes = [vp]
    lib.plk_destroy.argtypes = [vp]
    lib.plk_destroy.restype = None
    lib.plk_last_error.argtypes = [vp]
    lib.plk_last_error.restype = C.c_char_p
    lib.plk_sync.argtypes = [vp]
    lib.plk_set_pattern_weights.argtypes = [vp, vp, vp]
    lib.plk_set_tip_table.argtypes = [vp, C.c_int, vp]
    lib.plk_set_tip_codes.argtypes = [vp, C.c_int, vp]
    lib.plk_set_all_tip_codes.argtypes = [vp, vp, C.c_size_t]
    lib.plk_set_all_tip_codes_packed4.argtypes = [vp, vp, C.c_size_t]
    lib.plk_set_tip_vectors.argtypes = [vp, C.c_int, vp]
    lib.plk_set_model.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_double, C.c_int, C.c_double, C.c_double,
                                  C.c_double]
    lib.plk_update_pmats.argtypes = [vp, C.c_int, vp, vp]
    lib.plk_set_pmat.argtypes = [vp, C.c_int, vp]
    lib.plk_get_pmat.argtypes = [vp, C.c_int, vp]
    lib.plk_update_partials.argtypes = [vp, C.c_int, vp]
    lib.plk_edge_lnl.argtypes = [vp, _Side, _Side, C.c_int, dp, ip]
    lib.plk_traverse_edge_lnl.argtypes = [vp, C.c_int, vp, _Side, _Side, C.c_int, dp, ip]
    lib.plk_lk_full.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, _Side, _Side, C.c_int, dp, ip]
    lib.plk_lk_full_begin.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, _Side, _Side, C.c_int]
    lib.plk_lk_wait.argtypes = [vp, dp, ip]
    lib.plk_eigen_lr.argtypes = [vp, _Side, _Side]
    lib.plk_edge_lnl_dlnl.argtypes = [vp, dp, dp, dp, ip]
    lib.plk_edge_lnl_eigen.argtypes = [vp, C.c_double, dp, ip]
    lib.plk_get_clv.argtypes = [vp, C.c_int, vp, vp]
    lib.plk_set_clv.argtypes = [vp, C.c_int, vp, vp]
    lib.plk_get_site_lnl.argtypes = [vp, vp, vp, vp, vp]
    lib.plk_get_dot_prod.argtypes = [vp, vp]
    lib.plk_comm_unique_id.argtypes = [vp]
    lib.plk_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.plk_comm_set_allreduce.argtypes = [vp, C.c_int]
    lib.plk_comm_p2p_export.argtypes = [vp, C.c_int, vp]
    lib.plk_comm_p2p_init.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.plk_launch_count.argtypes = [vp]
    lib.plk_launch_count.restype = C.c_longlong
    lib.plk_device_bytes.argtypes = [vp]
    lib.plk_device_bytes.restype = C.c_size_t
    lib.plk_stream.argtypes = [vp]
    lib.plk_stream.restype = vp
    lib.plk_version.restype = C.c_char_p
    lib.plk_pars_create.argtypes = [vp, C.c_int, vp]
    lib.plk_pars_set_buffer.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.plk_pars_get_buffer.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.plk_pars_update.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.plk_pars_edge.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip]
    lib.plk_pars_traverse_edge.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, ip]
    lib.plk_get_site_pars.argtypes = [vp, vp]
    lib.plk_spr_candidates.argtypes = [vp, _Side, C.c_double, C.c_int, C.c_int, vp, vp, vp]
    _lib = lib
    return lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def pack_ops(ops: Sequence[PartialOp]) -> np.ndarray:
    """List of PartialOp -> contiguous array of C ``plk_op`` records."""
    arr = np.empty(len(ops), dtype=OP_DTYPE)
    for i, o in enumerate(ops):
        arr[i] = (o.dst, o.c1.tip, o.c1.clv, o.pmat1, o.c2.tip, o.c2.clv, o.pmat2)
    return arr


def pack_codes4(codes: np.ndarray) -> np.ndarray:
    """[n_tips, P] codes < 16 -> [n_tips, (P + 1) // 2] bytes, low nibble = the even pattern."""
    c = np.ascontiguousarray(codes, dtype=np.uint8)
    assert c.ndim == 2 and (c < 16).all()
    if c.shape[1] & 1:
        c = np.concatenate([c, np.zeros((c.shape[0], 1), dtype=np.uint8)], axis=1)
    return np.ascontiguousarray(c[:, 0::2] | (c[:, 1::2] << 4))


class Engine:
    """One device instance (``plk_instance``): the B200 engine behind one tree."""

    def __init__(self, n_tips: int, n_pattern: int, ns: int, ncatg: int, n_clv: int, n_pmat: int,
                 device: int = 0, apply_scaling: bool = True, devices: Optional[Sequence[int]] = None):
        """``devices``: site-shard the instance over these CUDA devices inside this process
        (``plk_create_sharded``; a device may be listed more than once)."""
        self.lib = load_library()
        self.n_tips, self.P, self.ns, self.ncatg = n_tips, n_pattern, ns, ncatg
        self.n_clv, self.n_pmat = n_clv, n_pmat
        cfg = _Config(n_tips, n_pattern, ns, ncatg, n_clv, n_pmat, device, 0 if apply_scaling else 1)
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.plk_create(C.byref(cfg), C.byref(h))
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.plk_create_sharded(C.byref(cfg), len(devices), arr, C.byref(h))
        if rc != 0:
            raise EngineError(f"plk_create failed ({rc}): {self.lib.plk_last_error(None).decode()}")
        self.h = h
        self.numerical_warning = 0

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc: int):
        if rc != 0:
            raise EngineError(f"engine call failed ({rc}): {self.lib.plk_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.plk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._ck(self.lib.plk_sync(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.plk_launch_count(self.h))

    @property
    def n_shards(self) -> int:
        return int(self.lib.plk_n_shards(self.h))

    @property
    def device_bytes(self) -> int:
        return int(self.lib.plk_device_bytes(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.plk_stream(self.h) or 0)

    # ------------------------------------------------------------------ uploads
    def set_weights(self, wght, invar=None):
        w = np.ascontiguousarray(wght, dtype=np.float64)
        assert w.shape == (self.P,)
        iv = None if invar is None else np.ascontiguousarray(invar, dtype=np.int16)
        self._ck(self.lib.plk_set_pattern_weights(self.h, _ptr(w), None if iv is None else _ptr(iv)))

    def set_tip_table(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.ndim == 2 and t.shape[1] == self.ns
        self._ck(self.lib.plk_set_tip_table(self.h, t.shape[0], _ptr(t)))

    def set_tip_codes(self, tip: int, codes):
        c = np.ascontiguousarray(codes, dtype=np.uint8)
        assert c.shape == (self.P,)
        self._ck(self.lib.plk_set_tip_codes(self.h, tip, _ptr(c)))

    def set_all_tip_codes(self, codes):
        """All tips in one 2-D copy; ``codes`` is a [n_tips, P] uint8 array (numpy, or anything
        exposing data_ptr()/stride() such as a pinned torch tensor)."""
        if hasattr(codes, "data_ptr"):
            assert tuple(codes.shape) == (self.n_tips, self.P)
            ptr, stride = codes.data_ptr(), codes.stride(0)
        else:
            c = np.ascontiguousarray(codes, dtype=np.uint8)
            assert c.shape == (self.n_tips, self.P)
            ptr, stride = c.ctypes.data, c.strides[0]
        self._ck(self.lib.plk_set_all_tip_codes(self.h, C.c_void_p(ptr), stride))

    def set_all_tip_codes_packed4(self, packed):
        """All tips as 4-bit codes, two patterns per byte (``pack_codes4``): [n_tips, (P + 1) // 2] uint8."""
        if hasattr(packed, "data_ptr"):
            assert tuple(packed.shape) == (self.n_tips, (self.P + 1) // 2)
            ptr, stride = packed.data_ptr(), packed.stride(0)
        else:
            c = np.ascontiguousarray(packed, dtype=np.uint8)
            assert c.shape == (self.n_tips, (self.P + 1) // 2)
            ptr, stride = c.ctypes.data, c.strides[0]
        self._ck(self.lib.plk_set_all_tip_codes_packed4(self.h, C.c_void_p(ptr), stride))

    def set_weights_ptr(self, wght_ptr: int, invar_ptr: int = 0):
        """Weights / invar from raw host pointers (pinned buffers)."""
        self._ck(self.lib.plk_set_pattern_weights(self.h, C.c_void_p(wght_ptr),
                                                  C.c_void_p(invar_ptr) if invar_ptr else None))

    def get_site_lnl_ptr(self, site_lnl_ptr: int):
        """Per-site lnL (c_lnL_sorted) into a raw host pointer."""
        self._ck(self.lib.plk_get_site_lnl(self.h, C.c_void_p(site_lnl_ptr), None, None, None))

    def set_tip_vectors(self, tip: int, vec, d_state=None, is_ambigu=None):
        """Reference-format tip (fp64 0/1 vectors, p_lk_tip_r).  d_state/is_ambigu are accepted for
        interface parity with the oracle backend; the engine derives both from the vectors."""
        v = np.ascontiguousarray(vec, dtype=np.float64).reshape(self.P, self.ns)
        self._ck(self.lib.plk_set_tip_vectors(self.h, tip, _ptr(v)))

    def set_model(self, m):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (m.U, m.V, m.lam, m.pi, m.rates, m.rate_probs)]
        assert a[0].shape == (self.ns, self.ns) and a[4].shape == (self.ncatg,)
        self._ck(self.lib.plk_set_model(self.h, *[_ptr(x) for x in a], float(m.pinv), int(bool(m.invar)),
                                        float(m.l_min), float(m.l_max), float(m.br_len_mult)))

    # ------------------------------------------------------------------ K0
    def update_pmats(self, handles: Iterable[int], lengths: Iterable[float]):
        h = np.ascontiguousarray(handles if isinstance(handles, np.ndarray) else list(handles), dtype=np.int32)
        l = np.ascontiguousarray(lengths if isinstance(lengths, np.ndarray) else list(lengths), dtype=np.float64)
        assert h.shape == l.shape
        self._ck(self.lib.plk_update_pmats(self.h, len(h), _ptr(h), _ptr(l)))

    def set_pmat(self, handle: int, P):
        p = np.ascontiguousarray(P, dtype=np.float64).reshape(self.ncatg, self.ns, self.ns)
        self._ck(self.lib.plk_set_pmat(self.h, handle, _ptr(p)))

    def get_pmat(self, handle: int) -> np.ndarray:
        out = np.empty((self.ncatg, self.ns, self.ns))
        self._ck(self.lib.plk_get_pmat(self.h, handle, _ptr(out)))
        return out

    # ------------------------------------------------------------------ K1
    def update_partials(self, ops):
        arr = ops if isinstance(ops, np.ndarray) else pack_ops(ops)
        self._ck(self.lib.plk_update_partials(self.h, len(arr), _ptr(arr)))

    # ------------------------------------------------------------------ K2
    def edge_lnl(self, left: Side, rght: Side, pmat: int) -> float:
        out = C.c_double(0.0)
        warn = C.c_int(0)
        self._ck(self.lib.plk_edge_lnl(self.h, _Side(left.tip, left.clv), _Side(rght.tip, rght.clv), pmat,
                                       C.byref(out), C.byref(warn)))
        self.numerical_warning = warn.value
        return out.value

    def traverse_edge_lnl(self, ops, left: Side, rght: Side, pmat: int) -> float:
        """Post_Order_Lk + the edge reduction in one call (one launch for 4-state, 4-category data)."""
        arr = ops if isinstance(ops, np.ndarray) else pack_ops(ops)
        out = C.c_double(0.0)
        warn = C.c_int(0)
        self._ck(self.lib.plk_traverse_edge_lnl(self.h, len(arr), _ptr(arr) if len(arr) else None,
                                                _Side(left.tip, left.clv), _Side(rght.tip, rght.clv), pmat,
                                                C.byref(out), C.byref(warn)))
        self.numerical_warning = warn.value
        return out.value

    def lk_full_call(self, handles, lengths, ops, left: Side, rght: Side, pmat: int):
        """Bind the arguments of ``plk_lk_full`` once (arrays are kept alive by the closure, ``lengths`` may be
        edited in place between calls) and return a zero-argument callable: one full-tree evaluation per call
        = all P-matrices + Post_Order_Lk + the site loop at the root edge."""
        h = np.ascontiguousarray(handles, dtype=np.int32)
        l = np.ascontiguousarray(lengths, dtype=np.float64)
        arr = ops if isinstance(ops, np.ndarray) else pack_ops(ops)
        out, warn = C.c_double(0.0), C.c_int(0)
        args = (self.h, len(h), _ptr(h), _ptr(l), len(arr), _ptr(arr), _Side(left.tip, left.clv), _Side(rght.tip, rght.clv),
                pmat, C.byref(out), C.byref(warn))
        fn, ck = self.lib.plk_lk_full, self._ck

        def call(_keep=(h, l, arr)):
            ck(fn(*args))
            self.numerical_warning = warn.value
            return out.value

        return call

    def lk_full_begin_call(self, handles, lengths, ops, left: Side, rght: Side, pmat: int):
        """``plk_lk_full_begin`` with bound arguments (see ``lk_full_call``): enqueues one full-tree evaluation and
        returns; ``lk_wait()`` returns its lnL.  Uploads of the next inputs may be issued in between."""
        h = np.ascontiguousarray(handles, dtype=np.int32)
        l = np.ascontiguousarray(lengths, dtype=np.float64)
        arr = ops if isinstance(ops, np.ndarray) else pack_ops(ops)
        args = (self.h, len(h), _ptr(h), _ptr(l), len(arr), _ptr(arr), _Side(left.tip, left.clv), _Side(rght.tip, rght.clv), pmat)
        fn, ck = self.lib.plk_lk_full_begin, self._ck

        def call(_keep=(h, l, arr)):
            ck(fn(*args))

        return call

    def lk_wait(self) -> float:
        out, warn = C.c_double(0.0), C.c_int(0)
        self._ck(self.lib.plk_lk_wait(self.h, C.byref(out), C.byref(warn)))
        self.numerical_warning = warn.value
        return out.value

    # ------------------------------------------------------------------ K3 / K4
    def eigen_lr(self, left: Side, rght: Side):
        self._ck(self.lib.plk_eigen_lr(self.h, _Side(left.tip, left.clv), _Side(rght.tip, rght.clv)))

    def lnl_dlnl(self, l: float):
        lc, lnl, d = C.c_double(float(l)), C.c_double(0.0), C.c_double(0.0)
        warn = C.c_int(0)
        self._ck(self.lib.plk_edge_lnl_dlnl(self.h, C.byref(lc), C.byref(lnl), C.byref(d), C.byref(warn)))
        self.numerical_warning = warn.value
        return lc.value, lnl.value, d.value

    def lnl_eigen(self, l: float) -> float:
        lnl = C.c_double(0.0)
        warn = C.c_int(0)
        self._ck(self.lib.plk_edge_lnl_eigen(self.h, float(l), C.byref(lnl), C.byref(warn)))
        self.numerical_warning = warn.value
        return lnl.value

    # ------------------------------------------------------------------ read-backs
    def get_clv(self, handle: int):
        clv = np.empty((self.P, self.ncatg, self.ns))
        sc = np.empty(self.P, dtype=np.int32)
        self._ck(self.lib.plk_get_clv(self.h, handle, _ptr(clv), _ptr(sc)))
        return clv, sc

    def set_clv(self, handle: int, clv, scale):
        c = np.ascontiguousarray(clv, dtype=np.float64).reshape(self.P, self.ncatg, self.ns)
        s = np.ascontiguousarray(scale, dtype=np.int32)
        self._ck(self.lib.plk_set_clv(self.h, handle, _ptr(c), _ptr(s)))

    def get_site_lnl(self):
        P = self.P
        out = {"site_lnl": np.empty(P), "site_lk": np.empty(P), "site_lk_cat": np.empty((P, self.ncatg)),
               "fact_sum_scale": np.empty(P, dtype=np.int32)}
        self._ck(self.lib.plk_get_site_lnl(self.h, _ptr(out["site_lnl"]), _ptr(out["site_lk"]),
                                           _ptr(out["site_lk_cat"]), _ptr(out["fact_sum_scale"])))
        return out

    def get_dot_prod(self) -> np.ndarray:
        out = np.empty((self.P, self.ncatg, self.ns))
        self._ck(self.lib.plk_get_dot_prod(self.h, _ptr(out)))
        return out

    # ------------------------------------------------------------------ batched SPR candidates
    @staticmethod
    def pack_spr_cands(cands):
        """Sequence of (Side a, l_a, Side b, l_b) -> C array of ``plk_spr_cand`` (reusable across calls)."""
        arr = (_SprCand * max(1, len(cands)))()
        for i, (a, la, b, lb) in enumerate(cands):
            arr[i] = _SprCand(_Side(a.tip, a.clv), float(la), _Side(b.tip, b.clv), float(lb))
        arr._n = len(cands)
        return arr

    def spr_candidates(self, prune: Side, l_prune: float, link_on_left: bool, cands):
        """Scores of many regraft positions of one pruned subtree in one call (``plk_spr_candidates``).
        ``cands``: sequence of (Side a, l_a, Side b, l_b) or the result of ``pack_spr_cands``.
        Returns (lnl array, warning flags)."""
        arr = cands if hasattr(cands, "_n") else self.pack_spr_cands(cands)
        n = arr._n
        lnl = np.zeros(n)
        warn = np.zeros(n, dtype=np.int32)
        self._ck(self.lib.plk_spr_candidates(self.h, _Side(prune.tip, prune.clv), float(l_prune), int(bool(link_on_left)),
                                             n, C.cast(arr, C.c_void_p), _ptr(lnl), _ptr(warn)))
        return lnl, warn

    # ------------------------------------------------------------------ parsimony (src/pars.c)
    def pars_create(self, n_buffers: int, step_mat=None):
        sm = None if step_mat is None else np.ascontiguousarray(step_mat, dtype=np.int32).reshape(self.ns, self.ns)
        self._ck(self.lib.plk_pars_create(self.h, n_buffers, None if sm is None else _ptr(sm)))

    def pars_set_buffer(self, buf: int, ui=None, pars=None, p_pars=None):
        a = [None if x is None else np.ascontiguousarray(x, dtype=np.int32) for x in (ui, pars, p_pars)]
        assert all(x is None or x.size == n for x, n in zip(a, (self.P, self.P, self.P * self.ns)))
        self._ck(self.lib.plk_pars_set_buffer(self.h, buf, *[None if x is None else _ptr(x) for x in a]))

    def pars_get_buffer(self, buf: int, general: bool = False):
        if general:
            pp = np.empty((self.P, self.ns), dtype=np.int32)
            self._ck(self.lib.plk_pars_get_buffer(self.h, buf, None, None, _ptr(pp)))
            return pp
        ui, pars = np.empty(self.P, dtype=np.int32), np.empty(self.P, dtype=np.int32)
        self._ck(self.lib.plk_pars_get_buffer(self.h, buf, _ptr(ui), _ptr(pars), None))
        return ui, pars

    @staticmethod
    def _pars_ops(ops) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(ops, dtype=np.int32).reshape(-1, 3))

    def pars_update(self, ops, general: bool = False):
        arr = self._pars_ops(ops)
        self._ck(self.lib.plk_pars_update(self.h, int(general), len(arr), _ptr(arr)))

    def pars_edge(self, left: int, rght: int, general: bool = False) -> int:
        out = C.c_int(0)
        self._ck(self.lib.plk_pars_edge(self.h, int(general), left, rght, C.byref(out)))
        return out.value

    def pars_traverse_edge(self, ops, left: int, rght: int, general: bool = False) -> int:
        arr = self._pars_ops(ops)
        out = C.c_int(0)
        self._ck(self.lib.plk_pars_traverse_edge(self.h, int(general), len(arr), _ptr(arr) if len(arr) else None,
                                                 left, rght, C.byref(out)))
        return out.value

    def get_site_pars(self) -> np.ndarray:
        out = np.empty(self.P, dtype=np.int32)
        self._ck(self.lib.plk_get_site_pars(self.h, _ptr(out)))
        return out

    # ------------------------------------------------------------------ site sharding
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.plk_comm_unique_id(buf)
        if rc != 0:
            raise EngineError(f"plk_comm_unique_id failed ({rc}): {lib.plk_last_error(None).decode()}")
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        assert len(unique_id) == 128
        self._ck(self.lib.plk_comm_init(self.h, rank, world, C.c_char_p(unique_id)))

    def comm_p2p_export(self, world: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.plk_comm_p2p_export(self.h, world, buf))
        return buf.raw

    def comm_p2p_init(self, rank: int, world: int, handles: Sequence[bytes]):
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        self._ck(self.lib.plk_comm_p2p_init(self.h, rank, world, C.c_char_p(blob)))

    def comm_set_allreduce(self, enable: bool):
        self._ck(self.lib.plk_comm_set_allreduce(self.h, int(enable)))
