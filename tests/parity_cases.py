"""Shared parity checks: the CUDA library (real device, or the CPU-thread emulation in the not-gpu
suite) against the oracle on the same seeded inputs."""
from __future__ import annotations

import numpy as np

from oracle import pyoracle as po
from quaff_b200 import api
from quaff_b200.params import QuaffNullParams, QuaffParams, random_params
from quaff_b200.seqs import FastSeq, add_revcomps
from quaff_b200.synth import random_ref, sample_reads

DEFAULT_PARAMS_JSON = None


def default_params() -> QuaffParams:
    """The reference's built-in parameters (data/defaultparams.json), committed as a golden fixture."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    return QuaffParams.load(os.path.join(here, "golden", "defaultparams.json"))


def make_workload(ref_len, n_reads, read_len, seed, both_strands=True, n_refs=1):
    refs = [random_ref(ref_len, seed + 100 * r, name=f"ref{r}") for r in range(n_refs)]
    reads = []
    for r, rf in enumerate(refs):
        rd, _, _ = sample_reads(rf, n_reads, read_len, seed + 1 + 100 * r, name_prefix=f"r{r}_")
        reads += rd
    x = add_revcomps(refs) if both_strands else refs
    return x, reads


def all_pairs(nx, ny):
    xi = np.tile(np.arange(nx, dtype=np.uint32), ny)
    yi = np.repeat(np.arange(ny, dtype=np.uint32), nx)
    return xi, yi


def oracle_cfg(cfg: api.DPConfig) -> po.CConfig:
    return po.make_config(cfg.sparse, cfg.kmer_len, cfg.kmer_threshold, cfg.band_size, cfg.local, cfg.max_size)


def seqbufs(x, reads, use_quals=True):
    xs = [po.SeqBuf(s.tokens(), None) for s in x]
    ys = [po.SeqBuf(s.tokens(), s.qual_scores() if use_quals else None) for s in reads]
    return xs, ys


def check_envelopes(G, O, x, reads, cfg, xi, yi, cell_size=24):
    xs, ys = seqbufs(x, reads)
    env, cu = G.envelopes(cfg, xi, yi, cell_size=cell_size)
    oc = oracle_cfg(cfg)
    for p, (a, b) in enumerate(zip(xi, yi)):
        d, c = O.envelope(xs[a], ys[b], oc, cell_size)
        assert np.array_equal(d, env[p]), f"pair {p}: envelope differs ({len(d)} vs {len(env[p])} diagonals)"
        assert c == cu[p], f"pair {p}: cell updates {c} vs {cu[p]}"
    return cu


def check_viterbi(G, O, x, reads, s_or, cfg, xi, yi, use_quals=True):
    """bit-exact: score, interval and the whole state path"""
    xs, ys = seqbufs(x, reads, use_quals)
    v = G.viterbi(cfg, xi, yi)
    oc = oracle_cfg(cfg)
    for p, (a, b) in enumerate(zip(xi, yi)):
        o = O.viterbi(xs[a], ys[b], s_or, oc)
        assert o["result"] == v["score"][p] or (np.isinf(o["result"]) and np.isinf(v["score"][p])), (p, o["result"], v["score"][p])
        assert o["x_start"] == v["x_start"][p] and o["x_end"] == v["x_end"][p], (p, o["x_start"], o["x_end"], v["x_start"][p], v["x_end"][p])
        assert np.array_equal(o["path"], v["paths"][p]), f"pair {p}: path differs"
    return v


FAST_LL_TOL = 1e-5      # nats per pair, probability-space kernels vs the reference's table log-sum-exp (stated bar: 1e-3)
FAST_COUNT_RTOL = 1e-6  # stated bar: 1e-4 relative


def check_forward(G, O, x, reads, s_or, cfg, xi, yi, use_quals=True, exact=True):
    """exact=True: log-space kernels, bit-exact (same table log-sum-exp, same fold order).
    exact=False: probability-space kernels, within FAST_LL_TOL nats (the stated tolerance is 1e-3 nats per read)"""
    xs, ys = seqbufs(x, reads, use_quals)
    G.set_fb_exact(exact)
    f = G.forward(cfg, xi, yi)
    oc = oracle_cfg(cfg)
    for p, (a, b) in enumerate(zip(xi, yi)):
        o = O.forward(xs[a], ys[b], s_or, oc)
        if np.isinf(o["result"]):
            assert np.isinf(f[p]) and f[p] < 0, (p, f[p])
        elif exact:
            assert o["result"] == f[p], (p, o["result"], f[p])
        else:
            assert abs(o["result"] - f[p]) <= FAST_LL_TOL, (p, o["result"], f[p])
    return f


def check_backward(G, O, x, reads, s_or, cfg, xi, yi, rel=1e-9, exact=True):
    """counts within 1e-4 relative is the stated bar; the log-space kernels hold 1e-9 (only exp() differs), the
    probability-space kernels FAST_COUNT_RTOL"""
    xs, ys = seqbufs(x, reads)
    G.set_fb_exact(exact)
    if not exact:
        return _check_backward_fast(G, O, xs, ys, s_or, cfg, xi, yi)
    r = G.backward_counts(cfg, xi, yi, per_pair=True)
    oc = oracle_cfg(cfg)
    tot = np.zeros_like(r["counts"])
    for p, (a, b) in enumerate(zip(xi, yi)):
        o = O.backward(xs[a], ys[b], s_or, oc)
        assert o["fwd"] == r["fwd"][p] or (np.isinf(o["fwd"]) and np.isinf(r["fwd"][p]))
        if np.isfinite(o["fwd"]):
            assert abs(o["back"] - r["back"][p]) <= 1e-9 * abs(o["back"]), (p, o["back"], r["back"][p])
            np.testing.assert_allclose(r["counts_per_pair"][p], o["counts"], rtol=rel, atol=1e-12)
            tot += o["counts"]
    np.testing.assert_allclose(r["counts"], tot, rtol=rel, atol=1e-12)
    return r


def check_estep(G, O, x, reads, s_or, nullp, cfg, use_null=True, n_iter=2, rel=1e-9, exact=True):
    """QuaffCountingTask::run replayed over `n_iter` E-steps (sortOrder pruning carried across)."""
    xs, ys = seqbufs(x, reads)
    G.set_fb_exact(exact)
    if not exact:
        rel = FAST_COUNT_RTOL
    null_ll = np.array([api.null_loglike(nullp, r, G.L) for r in reads])
    so_g = so_o = None
    for it in range(n_iter):
        g = G.estep(cfg, use_null, null_ll, so_g)
        o = O.estep(xs, ys, s_or, nullp, use_null, oracle_cfg(cfg), so_o)
        np.testing.assert_allclose(g["y_loglike"], o["loglike"], rtol=1e-12 if exact else 0, atol=0 if exact else FAST_LL_TOL)
        assert g["sort_order"] == o["sort_order"], (it, g["sort_order"], o["sort_order"])
        np.testing.assert_allclose(g["counts"], o["counts"], rtol=rel, atol=1e-12 if exact else 1e-9 * max(1.0, float(np.abs(o["counts"]).max())))
        assert abs(g["loglike"] - o["loglike"].sum()) <= 1e-9 * abs(g["loglike"])
        so_g, so_o = g["sort_order"], o["sort_order"]
    return g


def _check_backward_fast(G, O, xs, ys, s_or, cfg, xi, yi):
    r = G.backward_counts(cfg, xi, yi, per_pair=True)
    oc = oracle_cfg(cfg)
    tot = np.zeros_like(r["counts"])
    for p, (a, b) in enumerate(zip(xi, yi)):
        o = O.backward(xs[a], ys[b], s_or, oc)
        if not np.isfinite(o["fwd"]):
            assert np.isinf(r["fwd"][p])
            continue
        assert abs(o["fwd"] - r["fwd"][p]) <= FAST_LL_TOL and abs(o["back"] - r["back"][p]) <= FAST_LL_TOL, (p, o["fwd"], r["fwd"][p], o["back"], r["back"][p])
        scale = max(1.0, float(np.abs(o["counts"]).max()))
        np.testing.assert_allclose(r["counts_per_pair"][p], o["counts"], rtol=FAST_COUNT_RTOL, atol=1e-9 * scale)
        tot += o["counts"]
    np.testing.assert_allclose(r["counts"], tot, rtol=FAST_COUNT_RTOL, atol=1e-9 * max(1.0, float(np.abs(tot).max())))
    return r


def overlap_pairs(n_originals, n_total):
    """QuaffOverlapScheduler order (qoverlap.cpp:473-478, 528-547)."""
    xi, yi, yc = [], [], []
    for nx in range(max(n_originals - 1, 0)):
        for ny in range(nx + 1, n_total):
            xi.append(nx); yi.append(ny); yc.append(1 if ny >= n_originals else 0)
    return np.array(xi, np.uint32), np.array(yi, np.uint32), np.array(yc, np.uint8)


def check_overlap(G, O, seqs, n_originals, qp, cfg, use_quals=True):
    """bit-exact: result, end coordinates and both gapped rows (after the reference's indel squashing)"""
    bufs = [po.SeqBuf(s.tokens(), s.qual_scores() if use_quals else None) for s in seqs]
    os_ = [O.overlap_scores(qp, False), O.overlap_scores(qp, True)]
    xi, yi, yc = overlap_pairs(n_originals, len(seqs))
    r = G.overlap_viterbi(cfg, xi, yi, yc)
    oc = oracle_cfg(cfg)
    n_finite = 0
    for p in range(len(xi)):
        o = O.overlap_viterbi(bufs[xi[p]], bufs[yi[p]], os_[yc[p]], oc)
        assert o["result"] == r["score"][p] or (np.isinf(o["result"]) and np.isinf(r["score"][p])), (p, o["result"], r["score"][p])
        if np.isfinite(o["result"]):
            n_finite += 1
            assert tuple(int(v) for v in r["coords"][p]) == tuple(o["coords"]), (p, r["coords"][p], o["coords"])
            xr, yr = G.overlap_rows(bufs[xi[p]].tok, bufs[yi[p]].tok, r["coords"][p], r["paths"][p])
            assert xr == o["xrow"] and yr == o["yrow"], f"pair {p}: rows differ"
    return r, n_finite
