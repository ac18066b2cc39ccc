"""ctypes binding of libquaffgpu.so (include/quaffgpu.h) and the host-side mirror of the reference's
three driver seams (QuaffAligner::align, QuaffOverlapAligner::align, QuaffTrainer::getCounts).

There is no CPU path: if the CUDA library is missing or no device is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .params import QuaffNullParams, QuaffParams
from .seqs import FastSeq

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "libquaffgpu.so")

QG_REFS, QG_READS = 0, 1
NQ, NQ1 = 94, 95
OP_MATCH, OP_INSERT, OP_DELETE = 0, 1, 2
ERR_NAMES = {1: "QG_ERR_CUDA", 2: "QG_ERR_INVALID", 3: "QG_ERR_UNSUPPORTED", 4: "QG_ERR_NO_DEVICE", 5: "QG_ERR_STATE", 6: "QG_ERR_PRECONDITION"}

c_double_p = C.POINTER(C.c_double)
c_u8_p = C.POINTER(C.c_uint8)
c_u32_p = C.POINTER(C.c_uint32)
c_u64_p = C.POINTER(C.c_uint64)

# every symbol include/quaffgpu.h declares (checked by tests/test_abi.py against the header text)
ABI_SYMBOLS = [
    "qg_counts_size", "qg_create", "qg_set_option", "qg_destroy", "qg_last_error", "qg_free", "qg_abi_version", "qg_set_seqs",
    "qg_set_align_model", "qg_scores_from_params", "qg_null_loglike", "qg_envelopes", "qg_viterbi", "qg_forward",
    "qg_backward_counts", "qg_align_reads", "qg_align_reads_range", "qg_estep", "qg_set_overlap_model", "qg_overlap_viterbi", "qg_overlap_rows",
    "qg_overlap_reads", "qg_get_stats",
    "qg_device_count", "qg_init_devices", "qg_pool_create", "qg_pool_destroy", "qg_pool_last_error", "qg_pool_size", "qg_pool_context", "qg_pool_set_refs",
    "qg_pool_set_align_model", "qg_pool_set_option", "qg_pool_align_reads", "qg_pool_estep", "qg_pool_overlap_reads",
]


class QuaffGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class DPConfig(C.Structure):
    """qg_dpconfig == the QuaffDPConfig members the DP reads (qmodel.h:280-290)."""
    _fields_ = [("sparse", C.c_int32), ("kmer_len", C.c_int32), ("kmer_threshold", C.c_int32),
                ("band_size", C.c_int32), ("local", C.c_int32), ("max_size", C.c_uint64)]


def dp_config(sparse=True, kmer_len=6, kmer_threshold=14, band_size=64, local=True, max_size=0) -> DPConfig:
    return DPConfig(int(sparse), int(kmer_len), int(kmer_threshold), int(band_size), int(local), int(max_size))


class _AlignModel(C.Structure):
    _fields_ = [("match_k", C.c_int32), ("gap_k", C.c_int32), ("match", c_double_p), ("insert", c_double_p),
                ("m2m", c_double_p), ("m2i", c_double_p), ("m2d", c_double_p), ("m2e", c_double_p),
                ("d2d", C.c_double), ("d2m", C.c_double), ("i2i", C.c_double), ("i2m", C.c_double)]


class _Params(C.Structure):
    _fields_ = [("match_k", C.c_int32), ("gap_k", C.c_int32), ("ref_base", C.c_double * 4),
                ("begin_insert", c_double_p), ("begin_delete", c_double_p),
                ("extend_insert", C.c_double), ("extend_delete", C.c_double),
                ("insert_pqr", c_double_p), ("match_pqr", c_double_p)]


class _OverlapModel(C.Structure):
    _fields_ = [("match_k", C.c_int32), ("gap_k", C.c_int32), ("match", c_double_p), ("insert", c_double_p),
                ("log_ref_base", C.c_double * 4), ("begin_insert", c_double_p), ("begin_delete", c_double_p),
                ("extend_insert", C.c_double), ("extend_delete", C.c_double)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ms_seed", "ms_envelope", "ms_prep", "ms_viterbi", "ms_traceback", "ms_forward",
                                           "ms_backward", "ms_overlap", "ms_h2d", "ms_d2h")] + \
               [(n, C.c_uint64) for n in ("kernel_launches", "cell_updates", "kmer_hits", "trace_bytes", "fwd_store_bytes",
                                           "n_pairs", "n_segments")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def _dp(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


@dataclass
class AlignScores:
    """Flat QuaffScores tables (qmodel.h:181-191)."""
    match_k: int
    gap_k: int
    match: np.ndarray      # [4][4^K][95]
    insert: np.ndarray     # [4][95]
    m2m: np.ndarray
    m2i: np.ndarray
    m2d: np.ndarray
    m2e: np.ndarray
    d2d: float
    d2m: float
    i2i: float
    i2m: float


def load_library(path: Optional[str] = None) -> C.CDLL:
    path = path or os.environ.get("QUAFF_GPU_LIB") or DEFAULT_LIB      # QUAFF_GPU_LIB: development builds of the same library
    if not os.path.exists(path):
        raise QuaffGpuError(4, f"{path} not found: build it with `python -m quaff_b200.build` (nvcc, sm_100a); there is no CPU fallback")
    L = C.CDLL(path)
    L.qg_last_error.restype = C.c_char_p
    L.qg_last_error.argtypes = [C.c_void_p]
    L.qg_counts_size.restype = C.c_size_t
    L.qg_counts_size.argtypes = [C.c_int, C.c_int]
    L.qg_null_loglike.restype = C.c_double
    L.qg_null_loglike.argtypes = [C.c_double, c_double_p, c_u8_p, c_u8_p, C.c_uint64]
    L.qg_free.argtypes = [C.c_void_p]
    L.qg_destroy.argtypes = [C.c_void_p]
    L.qg_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.qg_pool_last_error.restype = C.c_char_p
    L.qg_pool_last_error.argtypes = [C.c_void_p]
    L.qg_pool_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int]
    L.qg_pool_destroy.argtypes = [C.c_void_p]
    L.qg_pool_context.restype = C.c_void_p
    L.qg_pool_context.argtypes = [C.c_void_p, C.c_int]
    L.qg_pool_size.argtypes = [C.c_void_p]
    return L


def scores_from_params(qp: QuaffParams, lib: Optional[C.CDLL] = None) -> AlignScores:
    """QuaffScores (qmodel.cpp:296-325) through the library's host helper."""
    L = lib or load_library()
    nK, nG = qp.n_match_kmers, qp.n_gap_kmers
    bi = np.ascontiguousarray(qp.begin_insert, dtype=np.float64)
    bd = np.ascontiguousarray(qp.begin_delete, dtype=np.float64)
    ipqr = np.ascontiguousarray(qp.insert_pqr()); mpqr = np.ascontiguousarray(qp.match_pqr())
    cp = _Params(qp.match_k, qp.gap_k, (C.c_double * 4)(*qp.ref_base), _dp(bi), _dp(bd), qp.extend_insert, qp.extend_delete, _dp(ipqr), _dp(mpqr))
    match = np.zeros((4, nK, NQ1)); insert = np.zeros((4, NQ1))
    m2m = np.zeros(nG); m2i = np.zeros(nG); m2d = np.zeros(nG); m2e = np.zeros(nG); scal = np.zeros(4)
    rc = L.qg_scores_from_params(C.byref(cp), _dp(match), _dp(insert), _dp(m2m), _dp(m2i), _dp(m2d), _dp(m2e), _dp(scal))
    if rc != 0:
        raise QuaffGpuError(rc, "qg_scores_from_params")
    return AlignScores(qp.match_k, qp.gap_k, match, insert, m2m, m2i, m2d, m2e, float(scal[0]), float(scal[1]), float(scal[2]), float(scal[3]))


def null_loglike(np_: QuaffNullParams, seq: FastSeq, lib: Optional[C.CDLL] = None) -> float:
    L = lib or load_library()
    pqr = np.ascontiguousarray(np_.pqr())
    tok = seq.tokens(); q = seq.qual_scores()
    return L.qg_null_loglike(np_.null_emit, _dp(pqr), tok.ctypes.data_as(c_u8_p), q.ctypes.data_as(c_u8_p) if q is not None else None, len(tok))


def _flatten(seqs: Sequence[FastSeq], want_qual: bool):
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    tok = np.concatenate([s.tokens() for s in seqs]) if len(seqs) else np.zeros(0, np.uint8)
    qual = None
    if want_qual:
        qual = np.concatenate([s.qual_scores() for s in seqs]) if len(seqs) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(tok, dtype=np.uint8), (None if qual is None else np.ascontiguousarray(qual, dtype=np.uint8)), off


class QuaffGPU:
    """One context = one GPU.  Mirrors the call sequence of the reference's drivers."""

    def __init__(self, device: int = 0, lib_path: Optional[str] = None):
        self.L = load_library(lib_path)
        self.ctx = C.c_void_p()
        rc = self.L.qg_create(C.byref(self.ctx), int(device))
        if rc != 0:
            raise QuaffGpuError(rc, (self.L.qg_last_error(None) or b"").decode())
        self.n = [0, 0]
        self.lens = [None, None]
        self.match_k, self.gap_k = 1, 0

    def close(self):
        if self.ctx:
            self.L.qg_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise QuaffGpuError(rc, (self.L.qg_last_error(self.ctx) or b"").decode())

    def set_option(self, option: int, value: int):
        self._check(self.L.qg_set_option(self.ctx, int(option), C.c_int64(int(value))))

    def set_fb_exact(self, exact: bool):
        """True: log-space Forward/Backward with the reference's table log-sum-exp (bit-exact Forward);
        False (default): probability-space kernels"""
        self.set_option(1, 1 if exact else 0)

    # ---- inputs ---------------------------------------------------------------------------------
    def set_seqs_raw(self, which: int, tok: np.ndarray, qual: Optional[np.ndarray], off: np.ndarray):
        tok = np.ascontiguousarray(tok, dtype=np.uint8); off = np.ascontiguousarray(off, dtype=np.uint64)
        if qual is not None:
            qual = np.ascontiguousarray(qual, dtype=np.uint8)
        self._check(self.L.qg_set_seqs(self.ctx, which, C.c_size_t(len(off) - 1), tok.ctypes.data_as(c_u8_p),
                                       qual.ctypes.data_as(c_u8_p) if qual is not None else None, off.ctypes.data_as(c_u64_p)))
        self.n[which] = len(off) - 1
        self.lens[which] = np.diff(off).astype(np.int64)

    def set_refs(self, refs: Sequence[FastSeq]):
        tok, _, off = _flatten(refs, False)
        self.set_seqs_raw(QG_REFS, tok, None, off)

    def set_reads(self, reads: Sequence[FastSeq], use_quals: bool = True):
        want = use_quals and all(r.has_qual() for r in reads) and len(reads) > 0
        tok, qual, off = _flatten(reads, want)
        self.set_seqs_raw(QG_READS, tok, qual, off)

    def set_align_scores(self, s: AlignScores):
        self._keep = s
        m = _AlignModel(s.match_k, s.gap_k, _dp(np.ascontiguousarray(s.match)), _dp(np.ascontiguousarray(s.insert)),
                        _dp(s.m2m), _dp(s.m2i), _dp(s.m2d), _dp(s.m2e), s.d2d, s.d2m, s.i2i, s.i2m)
        self._check(self.L.qg_set_align_model(self.ctx, C.byref(m)))
        self.match_k, self.gap_k = s.match_k, s.gap_k

    def set_params(self, qp: QuaffParams):
        self.set_align_scores(scores_from_params(qp, self.L))

    # ---- per-object entry points --------------------------------------------------------------------
    @staticmethod
    def _pairs(xi, yi):
        xi = np.ascontiguousarray(xi, dtype=np.uint32); yi = np.ascontiguousarray(yi, dtype=np.uint32)
        assert len(xi) == len(yi)
        return xi, yi

    def envelopes(self, cfg: DPConfig, xi, yi, cell_size: int = 24, x_set: int = QG_REFS):
        xi, yi = self._pairs(xi, yi)
        n = len(xi)
        d = C.POINTER(C.c_int32)(); off = np.zeros(n + 1, dtype=np.uint64); cu = np.zeros(n, dtype=np.uint64)
        self._check(self.L.qg_envelopes(self.ctx, C.byref(cfg), C.c_uint64(cell_size), x_set, C.c_size_t(n), xi.ctypes.data_as(c_u32_p),
                                        yi.ctypes.data_as(c_u32_p), C.byref(d), off.ctypes.data_as(c_u64_p), cu.ctypes.data_as(c_u64_p)))
        total = int(off[n])
        flat = np.ctypeslib.as_array(d, shape=(max(total, 1),))[:total].copy()
        self.L.qg_free(d)
        return [flat[int(off[p]):int(off[p + 1])] for p in range(n)], cu

    def _take_paths(self, ptr, off, n):
        total = int(off[n])
        flat = np.ctypeslib.as_array(ptr, shape=(max(total, 1),))[:total].copy() if ptr else np.zeros(0, np.uint8)
        if ptr:
            self.L.qg_free(ptr)
        return [flat[int(off[p]):int(off[p + 1])] for p in range(n)]

    def viterbi(self, cfg: DPConfig, xi, yi, want_path=None):
        xi, yi = self._pairs(xi, yi)
        n = len(xi)
        score = np.zeros(n); xs = np.zeros(n, np.uint32); xe = np.zeros(n, np.uint32)
        path = c_u8_p(); off = np.zeros(n + 1, dtype=np.uint64)
        wp = None if want_path is None else np.ascontiguousarray(want_path, dtype=np.uint8)
        self._check(self.L.qg_viterbi(self.ctx, C.byref(cfg), C.c_size_t(n), xi.ctypes.data_as(c_u32_p), yi.ctypes.data_as(c_u32_p),
                                      wp.ctypes.data_as(c_u8_p) if wp is not None else None, _dp(score), xs.ctypes.data_as(c_u32_p),
                                      xe.ctypes.data_as(c_u32_p), C.byref(path), off.ctypes.data_as(c_u64_p)))
        return dict(score=score, x_start=xs, x_end=xe, paths=self._take_paths(path, off, n))

    def forward(self, cfg: DPConfig, xi, yi) -> np.ndarray:
        xi, yi = self._pairs(xi, yi)
        ll = np.zeros(len(xi))
        self._check(self.L.qg_forward(self.ctx, C.byref(cfg), C.c_size_t(len(xi)), xi.ctypes.data_as(c_u32_p), yi.ctypes.data_as(c_u32_p), _dp(ll)))
        return ll

    def counts_size(self) -> int:
        return self.L.qg_counts_size(self.match_k, self.gap_k)

    def backward_counts(self, cfg: DPConfig, xi, yi, weights=None, per_pair=False):
        xi, yi = self._pairs(xi, yi)
        n = len(xi); nc = self.counts_size()
        f = np.zeros(n); b = np.zeros(n); csum = np.zeros(nc)
        cpp = np.zeros((n, nc)) if per_pair else None
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self._check(self.L.qg_backward_counts(self.ctx, C.byref(cfg), C.c_size_t(n), xi.ctypes.data_as(c_u32_p), yi.ctypes.data_as(c_u32_p),
                                              _dp(w) if w is not None else None, _dp(f), _dp(b), _dp(csum), _dp(cpp) if per_pair else None))
        return dict(fwd=f, back=b, counts=csum, counts_per_pair=cpp)

    # ---- seam A: QuaffAligner::align (qmodel.cpp:2624) -----------------------------------------------
    def align_reads(self, cfg: DPConfig, null_ll: np.ndarray, first: int = 0, count: Optional[int] = None, split_paths: bool = True):
        """best reference per read of READS[first : first + count] (default: the whole set)"""
        n = self.n[QG_READS] - first if count is None else int(count)
        null_ll = np.ascontiguousarray(null_ll, dtype=np.float64)
        assert len(null_ll) == n
        best = np.zeros(n, np.uint32); score = np.zeros(n); xs = np.zeros(n, np.uint32); xe = np.zeros(n, np.uint32)
        path = c_u8_p(); off = np.zeros(n + 1, dtype=np.uint64)
        self._check(self.L.qg_align_reads_range(self.ctx, C.byref(cfg), C.c_size_t(first), C.c_size_t(n), _dp(null_ll),
                                                best.ctypes.data_as(c_u32_p), _dp(score), xs.ctypes.data_as(c_u32_p),
                                                xe.ctypes.data_as(c_u32_p), C.byref(path), off.ctypes.data_as(c_u64_p)))
        if split_paths:
            paths = self._take_paths(path, off, n)
        else:
            total = int(off[n])
            paths = np.ctypeslib.as_array(path, shape=(max(total, 1),))[:total].copy() if path else np.zeros(0, np.uint8)
            if path:
                self.L.qg_free(path)
        return dict(best_ref=best, score=score, x_start=xs, x_end=xe, paths=paths, path_offsets=off)

    # ---- seam C: QuaffTrainer::getCounts (qmodel.cpp:2005) --------------------------------------------
    def estep(self, cfg: DPConfig, use_null: bool, null_ll: np.ndarray, sort_order: Optional[List[List[int]]] = None):
        ny, nx = self.n[QG_READS], self.n[QG_REFS]
        so = np.zeros((ny, nx), dtype=np.uint32); sl = np.zeros(ny, dtype=np.uint32)
        for m in range(ny):
            o = list(range(nx)) if sort_order is None else sort_order[m]
            so[m, :len(o)] = o; sl[m] = len(o)
        null_ll = np.ascontiguousarray(null_ll, dtype=np.float64)
        yll = np.zeros(ny); counts = np.zeros(self.counts_size()); tot = C.c_double()
        self._check(self.L.qg_estep(self.ctx, C.byref(cfg), int(use_null), _dp(null_ll), so.ctypes.data_as(c_u32_p), sl.ctypes.data_as(c_u32_p),
                                    _dp(yll), _dp(counts), C.byref(tot)))
        return dict(y_loglike=yll, counts=counts, loglike=tot.value, sort_order=[list(map(int, so[m, :sl[m]])) for m in range(ny)])

    # ---- overlap ---------------------------------------------------------------------------------------
    def set_overlap_params(self, qp: QuaffParams):
        s = scores_from_params(qp, self.L)
        self._okeep = (s, np.ascontiguousarray(qp.begin_insert, dtype=np.float64), np.ascontiguousarray(qp.begin_delete, dtype=np.float64))
        m = _OverlapModel(qp.match_k, qp.gap_k, _dp(np.ascontiguousarray(s.match)), _dp(np.ascontiguousarray(s.insert)),
                          (C.c_double * 4)(*[float(np.log(v)) for v in qp.ref_base]), _dp(self._okeep[1]), _dp(self._okeep[2]),
                          qp.extend_insert, qp.extend_delete)
        self._check(self.L.qg_set_overlap_model(self.ctx, C.byref(m)))

    def overlap_viterbi(self, cfg: DPConfig, xi, yi, y_complemented, want_path=None):
        xi, yi = self._pairs(xi, yi)
        n = len(xi)
        yc = np.ascontiguousarray(y_complemented, dtype=np.uint8)
        wp = None if want_path is None else np.ascontiguousarray(want_path, dtype=np.uint8)
        score = np.zeros(n); co = np.zeros((n, 4), np.uint32); path = c_u8_p(); off = np.zeros(n + 1, dtype=np.uint64)
        self._check(self.L.qg_overlap_viterbi(self.ctx, C.byref(cfg), C.c_size_t(n), xi.ctypes.data_as(c_u32_p), yi.ctypes.data_as(c_u32_p),
                                              yc.ctypes.data_as(c_u8_p), wp.ctypes.data_as(c_u8_p) if wp is not None else None,
                                              _dp(score), co.ctypes.data_as(c_u32_p), C.byref(path), off.ctypes.data_as(c_u64_p)))
        return dict(score=score, coords=co, paths=self._take_paths(path, off, n))

    def overlap_rows(self, x_tok: np.ndarray, y_tok: np.ndarray, coords4: np.ndarray, path: np.ndarray) -> Tuple[str, str]:
        x_tok = np.ascontiguousarray(x_tok, dtype=np.uint8); y_tok = np.ascontiguousarray(y_tok, dtype=np.uint8)
        co = np.ascontiguousarray(coords4, dtype=np.uint32); path = np.ascontiguousarray(path, dtype=np.uint8)
        xr = C.c_void_p(); yr = C.c_void_p()
        rc = self.L.qg_overlap_rows(x_tok.ctypes.data_as(c_u8_p), y_tok.ctypes.data_as(c_u8_p), co.ctypes.data_as(c_u32_p),
                                    path.ctypes.data_as(c_u8_p), C.c_uint64(len(path)), C.byref(xr), C.byref(yr))
        if rc != 0:
            raise QuaffGpuError(rc, "qg_overlap_rows")
        a, b = C.string_at(xr.value).decode(), C.string_at(yr.value).decode()
        self.L.qg_free(xr); self.L.qg_free(yr)
        return a, b

    # ---- instrumentation -------------------------------------------------------------------------------
    def stats(self, reset: bool = False) -> dict:
        st = Stats()
        self._check(self.L.qg_get_stats(self.ctx, C.byref(st), int(reset)))
        return st.as_dict()


class QuaffGPUPool:
    """Several contexts on ONE GPU, driven by one host thread each.  Every context owns a stream and its scratch, so the
    host-side work of one (pair lists, plans, result copies) and its kernels overlap with the kernels of the others: the
    seeding kernel (shared-memory atomics) and the DP fills (FP64 issue) share an SM well.  Reads are split contiguously
    over the contexts; the reference and the model are replicated (a few tens of MB)."""

    def __init__(self, device: int = 0, n_ctx: int = 2, lib_path: Optional[str] = None):
        from concurrent.futures import ThreadPoolExecutor
        self.ctxs = [QuaffGPU(device=device, lib_path=lib_path) for _ in range(n_ctx)]
        self.pool = ThreadPoolExecutor(max_workers=n_ctx)
        self.bounds: List[Tuple[int, int]] = []

    @property
    def L(self):
        return self.ctxs[0].L

    def close(self):
        self.pool.shutdown(wait=True)
        for g in self.ctxs:
            g.close()

    def _each(self, fn):
        return [f.result() for f in [self.pool.submit(fn, k, g) for k, g in enumerate(self.ctxs)]]

    def set_refs(self, refs):
        self._each(lambda k, g: g.set_refs(refs))

    def set_params(self, qp):
        s = scores_from_params(qp, self.L)
        self._each(lambda k, g: g.set_align_scores(s))

    def _split(self, n: int):
        w = len(self.ctxs)
        base, extra = divmod(n, w)
        self.bounds = []
        lo = 0
        for k in range(w):
            hi = lo + base + (1 if k < extra else 0)
            self.bounds.append((lo, hi)); lo = hi

    def set_reads_raw(self, tok: np.ndarray, qual: Optional[np.ndarray], off: np.ndarray):
        """host buffers -> each context uploads its contiguous share of the reads"""
        n = len(off) - 1
        self._split(n)

        def up(k, g):
            lo, hi = self.bounds[k]
            if hi == lo:
                return
            o = off[lo:hi + 1]
            b0, b1 = int(o[0]), int(o[-1])
            g.set_seqs_raw(QG_READS, tok[b0:b1], None if qual is None else qual[b0:b1], o - o[0])
        self._each(up)

    def set_reads(self, reads, use_quals: bool = True):
        want = use_quals and all(r.has_qual() for r in reads) and len(reads) > 0
        tok, qual, off = _flatten(reads, want)
        self.set_reads_raw(tok, qual, off)

    def set_read_batches(self, batches):
        """batches: list of (tok, qual, off) host arrays.  Every context receives its contiguous share of EVERY batch, laid
        out batch after batch, so that a later align_batch(b) only touches resident data."""
        w = len(self.ctxs)
        self.batch_ranges = []            # [batch][ctx] -> (first read in the context's set, count, global lo, global hi)
        per_ctx = [([], [], [np.zeros(1, dtype=np.uint64)]) for _ in range(w)]
        counts = [0] * w
        for tok, qual, off in batches:
            n = len(off) - 1
            base, extra = divmod(n, w)
            lo = 0; row = []
            for k in range(w):
                hi = lo + base + (1 if k < extra else 0)
                o = off[lo:hi + 1]; b0, b1 = int(o[0]), int(o[-1])
                t, q, offs = per_ctx[k]
                t.append(tok[b0:b1])
                if qual is not None:
                    q.append(qual[b0:b1])
                offs.append(o[1:] - o[0] + offs[-1][-1])
                row.append((counts[k], hi - lo, lo, hi)); counts[k] += hi - lo
                lo = hi
            self.batch_ranges.append(row)

        def up(k, g):
            t, q, offs = per_ctx[k]
            g.set_seqs_raw(QG_READS, np.concatenate(t), np.concatenate(q) if q else None, np.concatenate(offs))
        self._each(up)

    @staticmethod
    def _empty():
        return dict(best_ref=np.zeros(0, np.uint32), score=np.zeros(0), x_start=np.zeros(0, np.uint32), x_end=np.zeros(0, np.uint32),
                    paths=np.zeros(0, np.uint8), path_offsets=np.zeros(1, np.uint64))

    def _merge(self, parts):
        out = {k: np.concatenate([p[k] for p in parts]) for k in ("best_ref", "score", "x_start", "x_end", "paths")}
        base = np.uint64(0); cat = []
        for p in parts:
            o = np.asarray(p["path_offsets"], dtype=np.uint64)
            cat.append(o[:-1] + base); base = base + o[-1]
        out["path_offsets"] = np.concatenate(cat + [np.array([base], dtype=np.uint64)])
        return out

    def align_batch(self, cfg: DPConfig, b: int, null_ll: np.ndarray):
        """seam A over resident batch b (see set_read_batches); null_ll is the batch's, in read order"""
        null_ll = np.ascontiguousarray(null_ll, dtype=np.float64)

        def run(k, g):
            first, count, lo, hi = self.batch_ranges[b][k]
            if count == 0:
                return self._empty()
            return g.align_reads(cfg, null_ll[lo:hi], first=first, count=count, split_paths=False)
        return self._merge(self._each(run))

    def align_batches(self, cfg: DPConfig, steps):
        """steps: list of (resident batch index, its null_ll).  Each context works through its share of every step on its own
        (no join between steps: the contexts drift apart and keep the GPU busy while the others do host work); returns
        one merged result per step."""
        steps = [(b, np.ascontiguousarray(nl, dtype=np.float64)) for b, nl in steps]

        def run(k, g):
            outs = []
            for b, nl in steps:
                first, count, lo, hi = self.batch_ranges[b][k]
                outs.append(g.align_reads(cfg, nl[lo:hi], first=first, count=count, split_paths=False) if count else self._empty())
            return outs
        per_ctx = self._each(run)
        return [self._merge([per_ctx[k][t] for k in range(len(self.ctxs))]) for t in range(len(steps))]

    def align_stream(self, cfg: DPConfig, steps):
        """steps: list of (tok, qual, off, null_ll) HOST buffers.  Each context uploads its share of a step's reads and aligns
        them, step after step, independently of the other contexts; returns one merged result per step."""
        w = len(self.ctxs)
        prepared = []
        for tok, qual, off, nl in steps:
            n = len(off) - 1
            base, extra = divmod(n, w)
            bounds = []; lo = 0
            for k in range(w):
                hi = lo + base + (1 if k < extra else 0); bounds.append((lo, hi)); lo = hi
            prepared.append((tok, qual, off, np.ascontiguousarray(nl, dtype=np.float64), bounds))

        def run(k, g):
            outs = []
            for tok, qual, off, nl, bounds in prepared:
                lo, hi = bounds[k]
                if hi == lo:
                    outs.append(self._empty()); continue
                o = off[lo:hi + 1]
                b0, b1 = int(o[0]), int(o[-1])
                g.set_seqs_raw(QG_READS, tok[b0:b1], None if qual is None else qual[b0:b1], o - o[0])
                outs.append(g.align_reads(cfg, nl[lo:hi], split_paths=False))
            return outs
        per_ctx = self._each(run)
        return [self._merge([per_ctx[k][t] for k in range(w)]) for t in range(len(steps))]

    def align_reads(self, cfg: DPConfig, null_ll: np.ndarray):
        """seam A over everything uploaded with set_reads / set_reads_raw"""
        null_ll = np.ascontiguousarray(null_ll, dtype=np.float64)

        def run(k, g):
            lo, hi = self.bounds[k]
            if hi == lo:
                return self._empty()
            return g.align_reads(cfg, null_ll[lo:hi], split_paths=False)
        return self._merge(self._each(run))

    def stats(self, reset: bool = False) -> List[dict]:
        return [g.stats(reset) for g in self.ctxs]


CHUNK_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, c_u32_p, c_double_p, c_u32_p, c_u32_p, c_u8_p, c_u64_p)


class QuaffPool:
    """The native multi-context / multi-GPU handle (include/quaffgpu.h, qg_pool_*): one process drives every device; host
    threads, chunking and the E-step count sum live in the library.  `quaff ... -gpu 0,1,..` binds the same calls."""

    def __init__(self, devices: Sequence[int] = (0,), contexts_per_device: int = 2, lib_path: Optional[str] = None):
        self.L = load_library(lib_path)
        self.pool = C.c_void_p()
        dev = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = self.L.qg_pool_create(C.byref(self.pool), dev, len(devices), int(contexts_per_device))
        if rc != 0:
            raise QuaffGpuError(rc, (self.L.qg_pool_last_error(None) or b"").decode())
        self.n_refs = 0
        self.match_k, self.gap_k = 1, 0

    def close(self):
        if self.pool:
            self.L.qg_pool_destroy(self.pool)
            self.pool = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise QuaffGpuError(rc, (self.L.qg_pool_last_error(self.pool) or b"").decode())

    def size(self) -> int:
        return int(self.L.qg_pool_size(self.pool))

    def set_refs(self, refs: Sequence[FastSeq]):
        tok, _, off = _flatten(refs, False)
        self._check(self.L.qg_pool_set_refs(self.pool, C.c_size_t(len(refs)), tok.ctypes.data_as(c_u8_p), off.ctypes.data_as(c_u64_p)))
        self.n_refs = len(refs)

    def set_params(self, qp: QuaffParams):
        s = scores_from_params(qp, self.L)
        self._keep = s
        m = _AlignModel(s.match_k, s.gap_k, _dp(np.ascontiguousarray(s.match)), _dp(np.ascontiguousarray(s.insert)),
                        _dp(s.m2m), _dp(s.m2i), _dp(s.m2d), _dp(s.m2e), s.d2d, s.d2m, s.i2i, s.i2m)
        self._check(self.L.qg_pool_set_align_model(self.pool, C.byref(m)))
        self.match_k, self.gap_k = s.match_k, s.gap_k

    def set_fb_exact(self, exact: bool):
        self._check(self.L.qg_pool_set_option(self.pool, 1, C.c_int64(1 if exact else 0)))

    def align_reads_raw(self, cfg: DPConfig, tok: np.ndarray, qual: Optional[np.ndarray], off: np.ndarray, null_ll: np.ndarray, chunk_reads: int = 0):
        """seam A over host buffers; returns per-read arrays and the concatenated op paths in read order"""
        n = len(off) - 1
        null_ll = np.ascontiguousarray(null_ll, dtype=np.float64)
        best = np.zeros(n, np.uint32); score = np.zeros(n); xs = np.zeros(n, np.uint32); xe = np.zeros(n, np.uint32)
        plen = np.zeros(n + 1, np.uint64)
        chunks = {}

        def on_chunk(user, worker, first, cnt, b, sc, x0, x1, paths, poff):
            first, cnt = int(first), int(cnt)
            best[first:first + cnt] = np.ctypeslib.as_array(b, shape=(cnt,))
            score[first:first + cnt] = np.ctypeslib.as_array(sc, shape=(cnt,))
            xs[first:first + cnt] = np.ctypeslib.as_array(x0, shape=(cnt,))
            xe[first:first + cnt] = np.ctypeslib.as_array(x1, shape=(cnt,))
            o = np.ctypeslib.as_array(poff, shape=(cnt + 1,))
            plen[first + 1:first + cnt + 1] = np.diff(o)
            total = int(o[cnt])
            chunks[first] = np.ctypeslib.as_array(paths, shape=(max(total, 1),))[:total].copy() if total else np.zeros(0, np.uint8)
        cb = CHUNK_FN(on_chunk)
        self._check(self.L.qg_pool_align_reads(self.pool, C.byref(cfg), C.c_size_t(n), tok.ctypes.data_as(c_u8_p),
                                               qual.ctypes.data_as(c_u8_p) if qual is not None else None, off.ctypes.data_as(c_u64_p),
                                               _dp(null_ll), C.c_size_t(chunk_reads), cb, None))
        paths = np.concatenate([chunks[k] for k in sorted(chunks)]) if chunks else np.zeros(0, np.uint8)
        return dict(best_ref=best, score=score, x_start=xs, x_end=xe, paths=paths, path_offsets=np.cumsum(plen).astype(np.uint64))

    def align_reads(self, cfg: DPConfig, reads: Sequence[FastSeq], null_ll: np.ndarray, chunk_reads: int = 0, use_quals: bool = True):
        want = use_quals and len(reads) > 0 and all(r.has_qual() for r in reads)
        tok, qual, off = _flatten(reads, want)
        return self.align_reads_raw(cfg, tok, qual, off, null_ll, chunk_reads)

    def estep(self, cfg: DPConfig, use_null: bool, reads: Sequence[FastSeq], null_ll: np.ndarray, sort_order: Optional[List[List[int]]] = None):
        return self.estep_raw(cfg, use_null, *_flatten(reads, True), null_ll=null_ll, sort_order=sort_order)

    def estep_raw(self, cfg: DPConfig, use_null: bool, tok: np.ndarray, qual: np.ndarray, off: np.ndarray, null_ll: np.ndarray,
                  sort_order: Optional[List[List[int]]] = None):
        """seam C over host buffers (tokens, qualities, offsets of the whole read set)"""
        ny, nx = len(off) - 1, self.n_refs
        so = np.zeros((ny, nx), dtype=np.uint32); sl = np.zeros(ny, dtype=np.uint32)
        for m in range(ny):
            o = list(range(nx)) if sort_order is None else sort_order[m]
            so[m, :len(o)] = o; sl[m] = len(o)
        null_ll = np.ascontiguousarray(null_ll, dtype=np.float64)
        nc = int(self.L.qg_counts_size(self.match_k, self.gap_k))
        yll = np.zeros(ny); counts = np.zeros(nc); tot = C.c_double()
        self._check(self.L.qg_pool_estep(self.pool, C.byref(cfg), int(use_null), C.c_size_t(nx), C.c_size_t(ny), tok.ctypes.data_as(c_u8_p),
                                         qual.ctypes.data_as(c_u8_p), off.ctypes.data_as(c_u64_p), _dp(null_ll), so.ctypes.data_as(c_u32_p),
                                         sl.ctypes.data_as(c_u32_p), _dp(yll), _dp(counts), C.byref(tot)))
        return dict(y_loglike=yll, counts=counts, loglike=tot.value, sort_order=[list(map(int, so[m, :sl[m]])) for m in range(ny)])

    def stats(self, reset: bool = False) -> List[dict]:
        out = []
        for i in range(self.size()):
            st = Stats()
            self.L.qg_get_stats(C.c_void_p(self.L.qg_pool_context(self.pool, i)), C.byref(st), 1 if reset else 0)
            out.append(st.as_dict())
        return out
