"""Generates tests/golden/*.npz|json from the UNMODIFIED reference (oracle/_ref/libquaffref.so, built from
/root/reference by oracle/Makefile) and from the reference's own golden files.  Runs only where
/root/reference exists; the outputs are committed so that the oracle and the CUDA library can be checked
against the reference on machines that do not have it (the GPU box).

    python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as po                      # noqa: E402
from quaff_b200.params import QuaffNullParams, QuaffParams, random_params   # noqa: E402
from quaff_b200.seqs import FastSeq, add_revcomps, read_fastseqs             # noqa: E402
import parity_cases as pc                               # noqa: E402

REFDATA = "/root/reference/data"
R = po.Ref()


def rle(path):
    """run-length encoding of an op path: [(op, len), ...] flattened"""
    out = []
    for op in path:
        if out and out[-2] == op:
            out[-1] += 1
        else:
            out += [int(op), 1]
    return np.array(out, dtype=np.int32)


def params_payload(qp):
    return dict(match_k=qp.match_k, gap_k=qp.gap_k, begin_insert=list(map(float, qp.begin_insert)), begin_delete=list(map(float, qp.begin_delete)),
                extend_insert=qp.extend_insert, extend_delete=qp.extend_delete, ref_base=list(qp.ref_base),
                insert=[[d.p, d.q, d.r] for d in qp.insert], match=[[[d.p, d.q, d.r] for d in row] for row in qp.match])


def null_payload(n):
    return dict(null_emit=n.null_emit, null=[[d.p, d.q, d.r] for d in n.null])


def synthetic_case(name, qp, x, reads, cfg_kw, with_counts=True, n_estep=2, nullp=None):
    hp = R.params(qp); qp = R.params_as_parsed(hp, qp)
    cfg = po.make_config(**cfg_kw)
    hx = [R.seq(s) for s in x]; hy = [R.seq(s) for s in reads]
    nC = 4 * 4 ** qp.match_k * 94 + 4 * 94 + 4 * 4 ** qp.gap_k + 4
    out = dict(cfg=cfg_kw, params=params_payload(qp), x=[s.seq for s in x], reads=[[s.seq, s.qual] for s in reads], pairs=[])
    arrays = {}
    for m, hyy in enumerate(hy):
        for n, hxx in enumerate(hx):
            key = f"p{n}_{m}"
            d24, cu, _ = R.envelope(hxx, hyy, cfg, 24)
            d48, _, _ = R.envelope(hxx, hyy, cfg, 48)
            v = R.viterbi(hxx, hyy, hp, cfg)
            f = R.forward(hxx, hyy, hp, cfg)
            rec = dict(x=n, y=m, cu=int(cu), viterbi=v["result"], x_start=int(v["x_start"]), x_end=int(v["x_end"]), forward=f["result"])
            arrays[key + "_diag24"] = d24; arrays[key + "_diag48"] = d48; arrays[key + "_path"] = rle(v["path"])
            if with_counts and reads[m].has_qual():
                b = R.backward(hxx, hyy, hp, cfg, nC)
                rec["backward"] = b["back"]
                arrays[key + "_counts"] = b["counts"]
            out["pairs"].append(rec)
    if nullp is not None and n_estep:
        hn = R.null(nullp); nullp = R.null_as_parsed(hn, nullp)
        out["null"] = null_payload(nullp)
        out["null_loglike"] = [R.null_loglike(hn, h) for h in hy]
        so = None; out["estep"] = []
        for it in range(n_estep):
            e = R.estep(hx, hy, hp, hn, True, cfg, nC, so)
            so = e["sort_order"]
            out["estep"].append(dict(loglike=list(map(float, e["loglike"])), sort_order=so))
            arrays[f"estep{it}_counts"] = e["counts"]
    with open(os.path.join(HERE, name + ".json"), "w") as fh:
        json.dump(out, fh)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(name, len(out["pairs"]), "pairs")


def overlap_case(name, qp, seqs, n_orig, cfg_kw, nullp=None):
    hp = R.params(qp); qp = R.params_as_parsed(hp, qp)
    cfg = po.make_config(**cfg_kw)
    hs = [R.seq(s) for s in seqs]
    xi, yi, yc = pc.overlap_pairs(n_orig, len(seqs))
    out = dict(cfg=cfg_kw, params=params_payload(qp), seqs=[[s.seq, s.qual] for s in seqs], n_originals=n_orig, pairs=[])
    if nullp is not None:
        hn = R.null(nullp); nullp = R.null_as_parsed(hn, nullp)
        out["null"] = null_payload(nullp); out["null_loglike"] = [R.null_loglike(hn, h) for h in hs]
    for p in range(len(xi)):
        o = R.overlap(hs[xi[p]], hs[yi[p]], hp, cfg, bool(yc[p]))
        out["pairs"].append(dict(x=int(xi[p]), y=int(yi[p]), yc=int(yc[p]), result=o["result"] if np.isfinite(o["result"]) else None,
                                 coords=list(map(int, o["coords"])), xrow=o["xrow"], yrow=o["yrow"]))
    with open(os.path.join(HERE, name + ".json"), "w") as fh:
        json.dump(out, fh)
    print(name, len(out["pairs"]), "pairs")


def reference_own_goldens():
    """The three golden files the reference's `make test` diffs against (Makefile:146-156), reduced to the numbers
    they pin, plus the auto-fitted null model those runs used."""
    c8 = read_fastseqs(os.path.join(REFDATA, "c8f30.fastq.gz"))[0]
    hq = R.seq(c8)
    hn = R.null_fit([hq])
    nullp = R.null_as_parsed(hn, QuaffNullParams())
    qp = QuaffParams.load(os.path.join(REFDATA, "defaultparams.json")); hp = R.params(qp); qp = R.params_as_parsed(hp, qp)
    align_txt = open(os.path.join(REFDATA, "c8f30-self-align.json")).read()
    overlap_txt = open(os.path.join(REFDATA, "c8f30-self-overlap.json")).read()
    counts = json.load(open(os.path.join(REFDATA, "c8f30-self-counts.json")))
    out = dict(flags="-kmatchmb 10 -fwdstrand", max_size=10 << 20,
               align_score=float(re.search(r"#=GF Score (\S+)", align_txt).group(1)),
               overlap_score=float(re.search(r"#=GF Score (\S+)", overlap_txt).group(1)),
               params=params_payload(qp), null_fitted=null_payload(nullp), null_loglike=R.null_loglike(hn, hq),
               counts=counts)
    # the same read under the stated DEFAULT overlap flags (-kmatch 6 -kmatchn 14 -kmatchband 64), from the reference build
    copy = FastSeq("copy", c8.seq, c8.qual)
    hc = R.seq(copy)
    cfg = po.make_config(kmer_threshold=14)
    o = R.overlap(hq, hc, hp, cfg, False)
    d, cu, ts = R.envelope(hq, hc, cfg, 24)
    hn2 = R.null_fit([hq, hc])                               # `quaff overlap c8f30 copy-of-c8f30` fits its null model from both reads
    out["null_fitted_overlap"] = null_payload(R.null_as_parsed(hn2, QuaffNullParams()))
    out["null_loglike_overlap"] = R.null_loglike(hn2, hq)
    out["overlap_default"] = dict(result=o["result"], coords=list(map(int, o["coords"])), n_diagonals=int(len(d)), storage_cells=int(ts),
                                  score=o["result"] - 2 * R.null_loglike(hn2, hq))
    with open(os.path.join(HERE, "c8f30_reference_goldens.json"), "w") as fh:
        json.dump(out, fh)
    print("c8f30 goldens: align", out["align_score"], "overlap", out["overlap_score"], "default-flags overlap", out["overlap_default"]["score"])


def cfg5_case(name="cfg5_full"):
    """BASELINE config 5 at its stated shape: one 10 kb read against a 50 kb reference, `-kmatchoff -fwdstrand` (full DP,
    5e8 cells; the reference needs 12 GB for the matrix).  Only the numbers and the run-length path are kept; the inputs are
    regenerated from the seeds (reference seed 3, reads seed 4)."""
    from quaff_b200.synth import random_ref, sample_reads
    ref = random_ref(50_000, 3)
    reads, starts, strands = sample_reads(ref, 1, 10_000, 4, both_strands=False)
    qp = pc.default_params(); hp = R.params(qp); qp = R.params_as_parsed(hp, qp)
    cfg_kw = dict(sparse=False)
    cfg = po.make_config(**cfg_kw)
    hx, hy = R.seq(ref), R.seq(reads[0])
    v = R.viterbi(hx, hy, hp, cfg)
    f = R.forward(hx, hy, hp, cfg)
    out = dict(cfg=cfg_kw, params=params_payload(qp), ref_len=50_000, ref_seed=3, read_len=10_000, read_seed=4, read_actual_len=len(reads[0]),
               viterbi=v["result"], x_start=int(v["x_start"]), x_end=int(v["x_end"]), forward=f["result"], path_len=int(len(v["path"])))
    with open(os.path.join(HERE, name + ".json"), "w") as fh:
        json.dump(out, fh)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), path=rle(v["path"]))
    print(name, out["viterbi"], out["forward"], out["x_start"], out["x_end"], out["path_len"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cfg5":
        cfg5_case()
        sys.exit(0)
    reference_own_goldens()
    cfg5_case()
    nullp = QuaffNullParams.load(os.path.join(HERE, "testquaffnullparams.json"))
    x, reads = pc.make_workload(ref_len=9000, n_reads=3, read_len=800, seed=41)
    synthetic_case("synth_default", pc.default_params(), x, reads, dict(kmer_threshold=14), nullp=nullp)
    synthetic_case("synth_global_mem", pc.default_params(), x, reads[:2], dict(kmer_threshold=-1, max_size=500_000, local=False), n_estep=0)
    x2, reads2 = pc.make_workload(ref_len=2500, n_reads=2, read_len=500, seed=43)
    synthetic_case("synth_order2", random_params(np.random.default_rng(44), match_k=3, gap_k=2), x2, reads2, dict(kmer_threshold=8), nullp=nullp, n_estep=1)
    x3, reads3 = pc.make_workload(ref_len=400, n_reads=2, read_len=150, seed=45)
    synthetic_case("synth_full", pc.default_params(), x3, [FastSeq(r.name, r.seq) for r in reads3], dict(sparse=False), with_counts=False, n_estep=0)
    from quaff_b200.synth import random_ref, sample_reads
    rf = random_ref(1500, 46)
    rd, _, _ = sample_reads(rf, 3, 800, 47)
    overlap_case("overlap_default", pc.default_params(), add_revcomps(rd), 3, dict(kmer_threshold=14), nullp=nullp)
