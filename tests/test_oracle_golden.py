"""The oracle (oracle/quaff_oracle.c) against the committed reference vectors -- runs anywhere (no GPU, no
/root/reference).  Bit-exact except where noted."""
import json
import os

import numpy as np
import pytest

import golden_cases as gc
import parity_cases as pc
from oracle import pyoracle as po
from quaff_b200.seqs import FastSeq, read_fastseqs


@pytest.mark.parametrize("name", gc.SYNTH_CASES)
def test_oracle_matches_reference_vectors(oracle, name):
    meta, arr = gc.load_case(name)
    x, reads, qp = gc.case_inputs(meta)
    cfg = po.make_config(**meta["cfg"])
    s = oracle.scores(qp)
    xs = [po.SeqBuf(v.tokens(), None) for v in x]; ys = [po.SeqBuf(v.tokens(), v.qual_scores()) for v in reads]
    for rec in meta["pairs"]:
        n, m = rec["x"], rec["y"]; key = f"p{n}_{m}"
        d24, cu = oracle.envelope(xs[n], ys[m], cfg, 24); d48, _ = oracle.envelope(xs[n], ys[m], cfg, 48)
        assert np.array_equal(d24, arr[key + "_diag24"]) and np.array_equal(d48, arr[key + "_diag48"]) and cu == rec["cu"]
        v = oracle.viterbi(xs[n], ys[m], s, cfg)
        assert v["result"] == rec["viterbi"] and (v["x_start"], v["x_end"]) == (rec["x_start"], rec["x_end"])
        assert np.array_equal(v["path"], gc.unrle(arr[key + "_path"]))
        assert oracle.forward(xs[n], ys[m], s, cfg)["result"] == rec["forward"]
        if "backward" in rec:
            b = oracle.backward(xs[n], ys[m], s, cfg)
            assert (b["back"] == rec["backward"] or np.isnan(rec["backward"])) and np.array_equal(b["counts"], arr[key + "_counts"], equal_nan=True)
    if "estep" in meta:
        nullp = gc.null_from_payload(meta["null"])
        assert [oracle.null_loglike(nullp, y) for y in ys] == meta["null_loglike"]
        so = None
        for it, e in enumerate(meta["estep"]):
            r = oracle.estep(xs, ys, s, nullp, True, cfg, so)
            so = r["sort_order"]
            assert list(r["loglike"]) == e["loglike"] and so == e["sort_order"]
            assert np.array_equal(r["counts"], arr[f"estep{it}_counts"])


def test_oracle_overlap_vectors(oracle):
    meta, _ = gc.load_case("overlap_default")
    seqs = [FastSeq(f"s{n}", s, q) for n, (s, q) in enumerate(meta["seqs"])]
    qp = gc.params_from_payload(meta["params"])
    cfg = po.make_config(**meta["cfg"])
    bufs = [po.SeqBuf(v.tokens(), v.qual_scores()) for v in seqs]
    os_ = [oracle.overlap_scores(qp, False), oracle.overlap_scores(qp, True)]
    for rec in meta["pairs"]:
        o = oracle.overlap_viterbi(bufs[rec["x"]], bufs[rec["y"]], os_[rec["yc"]], cfg)
        if rec["result"] is None:
            assert np.isinf(o["result"])
        else:
            assert o["result"] == rec["result"] and list(o["coords"]) == rec["coords"]
            assert o["xrow"] == rec["xrow"] and o["yrow"] == rec["yrow"]


def test_reference_own_golden_files(oracle):
    """data/c8f30-self-{align,overlap,counts}.json (Makefile:146-156: -kmatchmb 10 -fwdstrand): the numbers those
    files pin, to the 6 significant digits they print."""
    with open(os.path.join(gc.GOLD, "c8f30_reference_goldens.json")) as fh:
        g = json.load(fh)
    c8 = read_fastseqs(os.path.join(gc.GOLD, "c8f30.fastq.gz"))[0]
    qp = gc.params_from_payload(g["params"]); nullp = gc.null_from_payload(g["null_fitted"])
    s = oracle.scores(qp)
    y = po.SeqBuf(c8.tokens(), c8.qual_scores()); x = po.SeqBuf(c8.tokens(), None)
    null_ll = oracle.null_loglike(nullp, y)
    assert null_ll == g["null_loglike"]
    cfg = po.make_config(kmer_threshold=-1, max_size=g["max_size"])
    d, _ = oracle.envelope(x, y, cfg, 24)
    assert list(d) == [0]                                  # the budget is smaller than one band: diagonal 0 only
    v = oracle.viterbi(x, y, s, cfg)
    assert float(f"{v['result'] - null_ll:.6g}") == g["align_score"]
    xq = po.SeqBuf(c8.tokens(), c8.qual_scores())
    o = oracle.overlap_viterbi(xq, y, oracle.overlap_scores(qp, False), cfg)
    null2 = oracle.null_loglike(gc.null_from_payload(g["null_fitted_overlap"]), y)     # fitted from both copies of the read
    assert null2 == g["null_loglike_overlap"]
    assert float(f"{o['result'] - 2 * null2:.6g}") == g["overlap_score"]
    od = oracle.overlap_viterbi(xq, y, oracle.overlap_scores(qp, False), po.make_config(kmer_threshold=14))
    assert od["result"] == g["overlap_default"]["result"] and list(od["coords"]) == g["overlap_default"]["coords"]
    # quaff count: QuaffParamCounts JSON (6 significant digits)
    e = oracle.estep([x], [y], s, nullp, True, po.make_config(kmer_threshold=-1, max_size=g["max_size"]))
    c = e["counts"]; nK = 4; off = 0
    cj = g["counts"]
    letters = "ACGT"
    def r6(v): return float(f"{v:.6g}")
    for i in range(4):
        for j in range(nK):
            exp = cj["match"][""][letters[i]][letters[j]]
            got = c[off:off + 94]; off += 94
            assert [r6(a) for a in got] == [r6(b) for b in exp]
    for i in range(4):
        exp = cj["insert"][letters[i]]; got = c[off:off + 94]; off += 94
        assert [r6(a) for a in got] == [r6(b) for b in exp]
    for key in ("beginInsertNo", "beginInsertYes", "beginDeleteNo", "beginDeleteYes"):
        assert r6(c[off]) == r6(cj[key][""]); off += 1
    for key in ("extendInsertNo", "extendInsertYes", "extendDeleteNo", "extendDeleteYes"):
        assert r6(c[off]) == r6(cj[key]); off += 1
