"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the two CPU checkers.

  * `Oracle`  -> oracle/libquafforacle.so, the portable C restatement (quaff_oracle.c)
  * `Ref`     -> oracle/_ref/libquaffref.so, the unmodified ihh/quaff sources behind C taps
                 (ref_harness.cpp); present only where it was built from /root/reference.

May be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never from quaff_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libquafforacle.so")
REF_SO = os.path.join(HERE, "_ref", "libquaffref.so")
REF_QUAFF = os.path.join(HERE, "_ref", "quaff")

NQ = 94
NQ1 = 95
c_double_p = C.POINTER(C.c_double)
c_u8_p = C.POINTER(C.c_uint8)
c_u32_p = C.POINTER(C.c_uint32)


def build_oracle() -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def ref_available() -> bool:
    return os.path.exists(REF_SO)


class SymQual(C.Structure):
    _fields_ = [("p", C.c_double), ("q", C.c_double), ("r", C.c_double)]


class CParams(C.Structure):
    _fields_ = [("match_k", C.c_int), ("gap_k", C.c_int), ("ref_base", C.c_double * 4),
                ("begin_insert", c_double_p), ("begin_delete", c_double_p),
                ("extend_insert", C.c_double), ("extend_delete", C.c_double),
                ("insert", SymQual * 4), ("match", C.POINTER(SymQual))]


class CNull(C.Structure):
    _fields_ = [("null_emit", C.c_double), ("null", SymQual * 4)]


class CScores(C.Structure):
    _fields_ = [("match_k", C.c_int), ("gap_k", C.c_int), ("match", c_double_p), ("insert", c_double_p),
                ("m2m", c_double_p), ("m2i", c_double_p), ("m2d", c_double_p), ("m2e", c_double_p),
                ("d2d", C.c_double), ("d2m", C.c_double), ("i2i", C.c_double), ("i2m", C.c_double)]


class CConfig(C.Structure):
    _fields_ = [("sparse", C.c_int), ("kmer_len", C.c_int), ("kmer_threshold", C.c_int),
                ("band_size", C.c_int), ("local", C.c_int), ("max_size", C.c_uint64)]


class CSeq(C.Structure):
    _fields_ = [("tok", c_u8_p), ("qual", c_u8_p), ("len", C.c_uint32)]


class COverlapScores(C.Structure):
    _fields_ = [("match_k", C.c_int), ("gap_k", C.c_int), ("y_complemented", C.c_int),
                ("m2m", c_double_p), ("m2i", c_double_p), ("m2d", c_double_p),
                ("i2m", C.c_double), ("i2i", C.c_double), ("i2d", C.c_double),
                ("d2m", C.c_double), ("d2i", C.c_double), ("d2d", C.c_double),
                ("pair", c_double_p), ("x_only", c_double_p), ("y_only", c_double_p), ("none", c_double_p),
                ("insert", C.c_double * (4 * NQ1))]


def _dp(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def make_config(sparse=True, kmer_len=6, kmer_threshold=14, band_size=64, local=True, max_size=0) -> CConfig:
    return CConfig(int(sparse), int(kmer_len), int(kmer_threshold), int(band_size), int(local), int(max_size))


class Scores:
    """Flat QuaffScores tables (numpy-owned) + the C view of them."""

    def __init__(self, match_k: int, gap_k: int):
        nK, nG = 4 ** match_k, 4 ** gap_k
        self.match_k, self.gap_k = match_k, gap_k
        self.match = np.zeros((4, nK, NQ1))
        self.insert = np.zeros((4, NQ1))
        self.m2m = np.zeros(nG); self.m2i = np.zeros(nG); self.m2d = np.zeros(nG); self.m2e = np.zeros(nG)
        self.d2d = self.d2m = self.i2i = self.i2m = 0.0

    def c(self) -> CScores:
        return CScores(self.match_k, self.gap_k, _dp(self.match), _dp(self.insert), _dp(self.m2m), _dp(self.m2i),
                       _dp(self.m2d), _dp(self.m2e), self.d2d, self.d2m, self.i2i, self.i2m)


class OverlapScores:
    def __init__(self, match_k: int, gap_k: int, y_complemented: bool):
        nK, nG = 4 ** match_k, 4 ** gap_k
        self.match_k, self.gap_k, self.y_complemented = match_k, gap_k, bool(y_complemented)
        self.m2m = np.zeros((nG, nG)); self.m2i = np.zeros((nG, nG)); self.m2d = np.zeros((nG, nG))
        self.pair = np.zeros((nK, nK, NQ, NQ))
        self.x_only = np.zeros((nK, nK, NQ)); self.y_only = np.zeros((nK, nK, NQ)); self.none = np.zeros((nK, nK))
        self.cs = COverlapScores(match_k, gap_k, int(y_complemented), _dp(self.m2m), _dp(self.m2i), _dp(self.m2d),
                                 0, 0, 0, 0, 0, 0, _dp(self.pair), _dp(self.x_only), _dp(self.y_only), _dp(self.none))

    @property
    def scal6(self):
        return np.array([self.cs.i2m, self.cs.i2i, self.cs.i2d, self.cs.d2m, self.cs.d2i, self.cs.d2d])

    @property
    def insert(self):
        return np.array(list(self.cs.insert)).reshape(4, NQ1)


class SeqBuf:
    """Keeps numpy buffers alive next to their CSeq view."""

    def __init__(self, tok: np.ndarray, qual: Optional[np.ndarray]):
        self.tok = np.ascontiguousarray(tok, dtype=np.uint8)
        self.qual = None if qual is None else np.ascontiguousarray(qual, dtype=np.uint8)
        self.c = CSeq(self.tok.ctypes.data_as(c_u8_p),
                      self.qual.ctypes.data_as(c_u8_p) if self.qual is not None else None, len(self.tok))


def _params_c(qp) -> Tuple[CParams, list]:
    """qp: quaff_b200.params.QuaffParams (duck-typed)."""
    nK = 4 ** qp.match_k
    bi = np.ascontiguousarray(qp.begin_insert, dtype=np.float64)
    bd = np.ascontiguousarray(qp.begin_delete, dtype=np.float64)
    marr = (SymQual * (4 * nK))()
    for i in range(4):
        for j in range(nK):
            d = qp.match[i][j]
            marr[i * nK + j] = SymQual(d.p, d.q, d.r)
    cp = CParams()
    cp.match_k, cp.gap_k = qp.match_k, qp.gap_k
    for i in range(4):
        cp.ref_base[i] = qp.ref_base[i]
        cp.insert[i] = SymQual(qp.insert[i].p, qp.insert[i].q, qp.insert[i].r)
    cp.begin_insert, cp.begin_delete = _dp(bi), _dp(bd)
    cp.extend_insert, cp.extend_delete = qp.extend_insert, qp.extend_delete
    cp.match = C.cast(marr, C.POINTER(SymQual))
    return cp, [bi, bd, marr]


def _null_c(np_) -> CNull:
    cn = CNull()
    cn.null_emit = np_.null_emit
    for i in range(4):
        cn.null[i] = SymQual(np_.null[i].p, np_.null[i].q, np_.null[i].r)
    return cn


def path_from_rows(xrow: str, yrow: str) -> np.ndarray:
    """align rows -> op codes 0=M 1=I (gap in ref row) 2=D (gap in read row)."""
    x = np.frombuffer(xrow.encode(), dtype=np.uint8)
    y = np.frombuffer(yrow.encode(), dtype=np.uint8)
    ops = np.zeros(len(x), dtype=np.uint8)
    ops[x == ord("-")] = 1
    ops[y == ord("-")] = 2
    return ops


# =================================================================================================
class Oracle:
    """The C restatement."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "quaff_oracle.c")):
            build_oracle()
        L = self.L = C.CDLL(ORACLE_SO)
        L.qo_lse.restype = C.c_double; L.qo_lse.argtypes = [C.c_double, C.c_double]
        L.qo_lse_unary.restype = C.c_double; L.qo_lse_unary.argtypes = [C.c_double]
        L.qo_lse_table.restype = c_double_p
        L.qo_log_negbinom.restype = C.c_double; L.qo_log_negbinom.argtypes = [C.c_int, C.c_double, C.c_double]
        L.qo_null_loglike.restype = C.c_double
        L.qo_counts_size.restype = C.c_size_t; L.qo_counts_size.argtypes = [C.c_int, C.c_int]
        L.qo_free.argtypes = [C.c_void_p]

    def lse(self, a, b): return self.L.qo_lse(a, b)
    def lse_unary(self, x): return self.L.qo_lse_unary(x)

    def lse_table(self) -> np.ndarray:
        n = C.c_int()
        p = self.L.qo_lse_table(C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def context_kmers(self, tok: np.ndarray, k: int) -> np.ndarray:
        tok = np.ascontiguousarray(tok, dtype=np.uint8)
        out = np.zeros(len(tok), dtype=np.uint32)
        self.L.qo_context_kmers(tok.ctypes.data_as(c_u8_p), C.c_uint32(len(tok)), C.c_int(k), out.ctypes.data_as(c_u32_p))
        return out

    def scores(self, qp) -> Scores:
        cp, keep = _params_c(qp)
        s = Scores(qp.match_k, qp.gap_k)
        cs = s.c()
        self.L.qo_scores_from_params(C.byref(cp), C.byref(cs))
        s.d2d, s.d2m, s.i2i, s.i2m = cs.d2d, cs.d2m, cs.i2i, cs.i2m
        return s

    def null_loglike(self, np_, seq: SeqBuf) -> float:
        cn = _null_c(np_)
        return self.L.qo_null_loglike(C.byref(cn), C.byref(seq.c))

    def envelope(self, x: SeqBuf, y: SeqBuf, cfg: CConfig, cell_size: int = 24):
        d = C.POINTER(C.c_int32)(); cu = C.c_uint64()
        n = self.L.qo_envelope(C.byref(x.c), C.byref(y.c), C.byref(cfg), C.c_uint64(cell_size), C.byref(d), C.byref(cu))
        if n < 0:
            raise RuntimeError("qo_envelope failed (sequence shorter than k?)")
        out = np.ctypeslib.as_array(d, shape=(n,)).copy()
        self.L.qo_free(d)
        return out, cu.value

    def _cells(self, p, n):
        a = np.ctypeslib.as_array(p, shape=(n.value, 3)).copy() if n.value else np.zeros((0, 3))
        self.L.qo_free(p)
        return a

    def viterbi(self, x: SeqBuf, y: SeqBuf, s: Scores, cfg: CConfig, want_cells=False):
        res = C.c_double(); xs = C.c_uint32(); xe = C.c_uint32()
        path = c_u8_p(); plen = C.c_uint32(); cells = c_double_p(); ncells = C.c_uint64()
        cs = s.c()
        rc = self.L.qo_viterbi(C.byref(x.c), C.byref(y.c), C.byref(cs), C.byref(cfg), C.byref(res), C.byref(xs), C.byref(xe),
                               C.byref(path), C.byref(plen), C.byref(cells) if want_cells else None, C.byref(ncells))
        if rc != 0:
            raise RuntimeError(f"qo_viterbi rc={rc}")
        p = np.ctypeslib.as_array(path, shape=(plen.value,)).copy() if plen.value else np.zeros(0, np.uint8)
        if plen.value:
            self.L.qo_free(path)
        out = dict(result=res.value, x_start=xs.value, x_end=xe.value, path=p)
        if want_cells:
            out["cells"] = self._cells(cells, ncells)
        return out

    def forward(self, x, y, s: Scores, cfg, want_cells=False):
        res = C.c_double(); cells = c_double_p(); ncells = C.c_uint64()
        cs = s.c()
        rc = self.L.qo_forward(C.byref(x.c), C.byref(y.c), C.byref(cs), C.byref(cfg), C.byref(res),
                               C.byref(cells) if want_cells else None, C.byref(ncells))
        if rc != 0:
            raise RuntimeError(f"qo_forward rc={rc}")
        out = dict(result=res.value)
        if want_cells:
            out["cells"] = self._cells(cells, ncells)
        return out

    def counts_size(self, match_k, gap_k) -> int:
        return self.L.qo_counts_size(match_k, gap_k)

    def backward(self, x, y, s: Scores, cfg, want_cells=False):
        f = C.c_double(); b = C.c_double(); cells = c_double_p(); ncells = C.c_uint64()
        counts = np.zeros(self.counts_size(s.match_k, s.gap_k))
        cs = s.c()
        rc = self.L.qo_backward(C.byref(x.c), C.byref(y.c), C.byref(cs), C.byref(cfg), C.byref(f), C.byref(b), _dp(counts),
                                C.byref(cells) if want_cells else None, C.byref(ncells))
        if rc != 0:
            raise RuntimeError(f"qo_backward rc={rc}")
        out = dict(fwd=f.value, back=b.value, counts=counts)
        if want_cells:
            out["cells"] = self._cells(cells, ncells)
        return out

    def estep(self, xs: Sequence[SeqBuf], ys: Sequence[SeqBuf], s: Scores, np_, use_null: bool, cfg,
              sort_order: Optional[List[List[int]]] = None):
        nx, ny = len(xs), len(ys)
        xa = (CSeq * nx)(*[v.c for v in xs]); ya = (CSeq * ny)(*[v.c for v in ys])
        so = np.zeros((ny, nx), dtype=np.uint32); sl = np.zeros(ny, dtype=np.uint32)
        for m in range(ny):
            o = list(range(nx)) if sort_order is None else sort_order[m]
            so[m, :len(o)] = o; sl[m] = len(o)
        ll = np.zeros(ny); counts = np.zeros(self.counts_size(s.match_k, s.gap_k))
        cs = s.c(); cn = _null_c(np_)
        rc = self.L.qo_estep(xa, nx, ya, ny, C.byref(cs), C.byref(cn), int(use_null), C.byref(cfg),
                             so.ctypes.data_as(c_u32_p), sl.ctypes.data_as(c_u32_p), _dp(ll), _dp(counts))
        if rc != 0:
            raise RuntimeError(f"qo_estep rc={rc}")
        return dict(loglike=ll, counts=counts, sort_order=[list(map(int, so[m, :sl[m]])) for m in range(ny)])

    def overlap_scores(self, qp, y_complemented: bool) -> OverlapScores:
        cp, keep = _params_c(qp)
        o = OverlapScores(qp.match_k, qp.gap_k, y_complemented)
        self.L.qo_overlap_scores_from_params(C.byref(cp), int(y_complemented), C.byref(o.cs))
        return o

    def overlap_viterbi(self, x, y, o: OverlapScores, cfg, want_cells=False):
        res = C.c_double(); co = (C.c_uint32 * 4)(); xr = C.c_char_p(); yr = C.c_char_p()
        xrp = C.c_void_p(); yrp = C.c_void_p(); cells = c_double_p(); ncells = C.c_uint64()
        rc = self.L.qo_overlap_viterbi(C.byref(x.c), C.byref(y.c), C.byref(o.cs), C.byref(cfg), C.byref(res), co,
                                       C.byref(xrp), C.byref(yrp), C.byref(cells) if want_cells else None, C.byref(ncells))
        if rc != 0:
            raise RuntimeError(f"qo_overlap_viterbi rc={rc}")
        out = dict(result=res.value, coords=tuple(co), xrow="", yrow="")
        if xrp.value:
            out["xrow"] = C.string_at(xrp.value).decode(); out["yrow"] = C.string_at(yrp.value).decode()
            self.L.qo_free(xrp); self.L.qo_free(yrp)
        if want_cells:
            out["cells"] = self._cells(cells, ncells)
        return out


# =================================================================================================
class Ref:
    """The unmodified reference behind oracle/ref_harness.cpp (only where oracle/_ref was built)."""

    def __init__(self):
        if not ref_available():
            raise FileNotFoundError(REF_SO)
        L = self.L = C.CDLL(REF_SO)
        for f in ("qref_params_from_json", "qref_null_from_json", "qref_seq_new", "qref_seq_revcomp"):
            getattr(L, f).restype = C.c_void_p
        for f in ("qref_params_to_json", "qref_seq_bases", "qref_seq_quals"):
            getattr(L, f).restype = C.c_void_p
        L.qref_null_loglike.restype = C.c_double; L.qref_null_loglike.argtypes = [C.c_void_p, C.c_void_p]
        L.qref_lse.restype = C.c_double; L.qref_lse.argtypes = [C.c_double, C.c_double]
        L.qref_lse_unary.restype = C.c_double; L.qref_lse_unary.argtypes = [C.c_double]
        L.qref_free.argtypes = [C.c_void_p]
        L.qref_seq_new.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.qref_params_from_json.argtypes = [C.c_char_p]
        L.qref_null_from_json.argtypes = [C.c_char_p]

    def params(self, qp) -> int:
        return self.L.qref_params_from_json(qp.to_json().encode())

    def null(self, np_) -> int:
        return self.L.qref_null_from_json(np_.to_json().encode())

    def params_as_parsed(self, hp, qp):
        """copy of `qp` carrying the values exactly as the reference parsed them"""
        import copy
        out = copy.deepcopy(qp)
        nK, nG = 4 ** qp.match_k, 4 ** qp.gap_k
        bi = np.zeros(nG); bd = np.zeros(nG); ext = np.zeros(2); ins = np.zeros(12); mat = np.zeros(12 * nK); rb = np.zeros(4)
        self.L.qref_params_values(C.c_void_p(hp), _dp(bi), _dp(bd), _dp(ext), _dp(ins), _dp(mat), _dp(rb))
        out.begin_insert, out.begin_delete = bi, bd
        out.extend_insert, out.extend_delete = float(ext[0]), float(ext[1])
        out.ref_base = [float(v) for v in rb]
        for i in range(4):
            out.insert[i].p, out.insert[i].q, out.insert[i].r = map(float, ins[3*i:3*i+3])
            for j in range(nK):
                d = out.match[i][j]
                d.p, d.q, d.r = map(float, mat[3*(i*nK+j):3*(i*nK+j)+3])
        return out

    def null_fit(self, hseqs):
        """QuaffNullParams fitted from reads, as `quaff align/overlap/count` do without -null"""
        arr = (C.c_void_p * len(hseqs))(*hseqs)
        self.L.qref_null_fit.restype = C.c_void_p
        return self.L.qref_null_fit(arr, len(hseqs))

    def null_as_parsed(self, hn, np_):
        import copy
        out = copy.deepcopy(np_)
        ne = C.c_double(); pqr = np.zeros(12)
        self.L.qref_null_values(C.c_void_p(hn), C.byref(ne), _dp(pqr))
        out.null_emit = ne.value
        for i in range(4):
            out.null[i].p, out.null[i].q, out.null[i].r = map(float, pqr[3*i:3*i+3])
        return out

    def seq(self, fs) -> int:
        """fs: quaff_b200.seqs.FastSeq"""
        return self.L.qref_seq_new(fs.name.encode(), fs.seq.encode(), fs.qual.encode("latin-1") if fs.has_qual() else None)

    def lse(self, a, b): return self.L.qref_lse(a, b)

    def scores(self, hp, match_k, gap_k) -> Scores:
        s = Scores(match_k, gap_k)
        scal = np.zeros(4)
        self.L.qref_scores(C.c_void_p(hp), _dp(s.match), _dp(s.insert), _dp(s.m2m), _dp(s.m2i), _dp(s.m2d), _dp(s.m2e), _dp(scal))
        s.d2d, s.d2m, s.i2i, s.i2m = map(float, scal)
        return s

    def null_loglike(self, hn, hs) -> float:
        return self.L.qref_null_loglike(hn, hs)

    def kmers(self, hs, n, k) -> np.ndarray:
        out = np.zeros(n, dtype=np.uint64)
        self.L.qref_seq_kmers(C.c_void_p(hs), k, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out

    def envelope(self, hx, hy, cfg: CConfig, cell_size=24):
        d = C.POINTER(C.c_int)(); ts = C.c_uint64(); cu = C.c_uint64()
        n = self.L.qref_envelope(C.c_void_p(hx), C.c_void_p(hy), C.byref(cfg), C.c_uint64(cell_size), C.byref(d), C.byref(ts), C.byref(cu))
        out = np.ctypeslib.as_array(d, shape=(n,)).astype(np.int32).copy()
        self.L.qref_free(d)
        return out, cu.value, ts.value

    def _cells(self, p, n):
        a = np.ctypeslib.as_array(p, shape=(n.value, 3)).copy() if n.value else np.zeros((0, 3))
        self.L.qref_free(p)
        return a

    def _rows(self, xr, yr):
        if not xr.value:
            return "", ""
        a, b = C.string_at(xr.value).decode(), C.string_at(yr.value).decode()
        self.L.qref_free(xr); self.L.qref_free(yr)
        return a, b

    def viterbi(self, hx, hy, hp, cfg, want_cells=False):
        res = C.c_double(); xs = C.c_uint32(); xe = C.c_uint32(); xr = C.c_void_p(); yr = C.c_void_p()
        cells = c_double_p(); ncells = C.c_uint64(); secs = (C.c_double * 3)()
        self.L.qref_viterbi(C.c_void_p(hx), C.c_void_p(hy), C.c_void_p(hp), C.byref(cfg), C.byref(res), C.byref(xs), C.byref(xe),
                            C.byref(xr), C.byref(yr), C.byref(cells) if want_cells else None, C.byref(ncells), secs)
        xrow, yrow = self._rows(xr, yr)
        out = dict(result=res.value, x_start=xs.value, x_end=xe.value, xrow=xrow, yrow=yrow,
                   path=path_from_rows(xrow, yrow), seconds=tuple(secs))
        if want_cells:
            out["cells"] = self._cells(cells, ncells)
        return out

    def forward(self, hx, hy, hp, cfg, want_cells=False):
        res = C.c_double(); cells = c_double_p(); ncells = C.c_uint64(); secs = (C.c_double * 2)()
        self.L.qref_forward(C.c_void_p(hx), C.c_void_p(hy), C.c_void_p(hp), C.byref(cfg), C.byref(res),
                            C.byref(cells) if want_cells else None, C.byref(ncells), secs)
        out = dict(result=res.value, seconds=tuple(secs))
        if want_cells:
            out["cells"] = self._cells(cells, ncells)
        return out

    def backward(self, hx, hy, hp, cfg, n_counts, want_cells=False):
        f = C.c_double(); b = C.c_double(); cells = c_double_p(); ncells = C.c_uint64(); secs = (C.c_double * 2)()
        counts = np.zeros(n_counts)
        self.L.qref_backward(C.c_void_p(hx), C.c_void_p(hy), C.c_void_p(hp), C.byref(cfg), C.byref(f), C.byref(b), _dp(counts),
                             C.byref(cells) if want_cells else None, C.byref(ncells), secs)
        out = dict(fwd=f.value, back=b.value, counts=counts, seconds=tuple(secs))
        if want_cells:
            out["cells"] = self._cells(cells, ncells)
        return out

    def overlap(self, hx, hy, hp, cfg, y_complemented, want_cells=False):
        res = C.c_double(); co = (C.c_uint32 * 4)(); xr = C.c_void_p(); yr = C.c_void_p()
        cells = c_double_p(); ncells = C.c_uint64()
        self.L.qref_overlap(C.c_void_p(hx), C.c_void_p(hy), C.c_void_p(hp), C.byref(cfg), int(y_complemented), C.byref(res), co,
                            C.byref(xr), C.byref(yr), C.byref(cells) if want_cells else None, C.byref(ncells))
        xrow, yrow = self._rows(xr, yr)
        out = dict(result=res.value, coords=tuple(co), xrow=xrow, yrow=yrow)
        if want_cells:
            out["cells"] = self._cells(cells, ncells)
        return out

    def overlap_scores(self, hp, match_k, gap_k, y_complemented) -> OverlapScores:
        o = OverlapScores(match_k, gap_k, y_complemented)
        scal = np.zeros(6)
        self.L.qref_overlap_scores(C.c_void_p(hp), int(y_complemented), _dp(scal), _dp(o.m2m), _dp(o.m2i), _dp(o.m2d),
                                   _dp(o.pair), _dp(o.x_only), _dp(o.y_only), _dp(o.none))
        o.cs.i2m, o.cs.i2i, o.cs.i2d, o.cs.d2m, o.cs.d2i, o.cs.d2d = map(float, scal)
        return o

    def estep(self, hxs, hys, hp, hn, use_null, cfg, n_counts, sort_order=None):
        nx, ny = len(hxs), len(hys)
        xa = (C.c_void_p * nx)(*hxs); ya = (C.c_void_p * ny)(*hys)
        so = np.zeros((ny, nx), dtype=np.uint32); sl = np.zeros(ny, dtype=np.uint32)
        for m in range(ny):
            o = list(range(nx)) if sort_order is None else sort_order[m]
            so[m, :len(o)] = o; sl[m] = len(o)
        ll = np.zeros(ny); counts = np.zeros(n_counts)
        self.L.qref_estep(xa, nx, ya, ny, C.c_void_p(hp), C.c_void_p(hn), int(use_null), C.byref(cfg),
                          so.ctypes.data_as(c_u32_p), sl.ctypes.data_as(c_u32_p), _dp(ll), _dp(counts))
        return dict(loglike=ll, counts=counts, sort_order=[list(map(int, so[m, :sl[m]])) for m in range(ny)])
