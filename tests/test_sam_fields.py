"""not-gpu: quaff_b200.sam (the SAM fields bench.py compares) against the text the UNMODIFIED reference CLI prints
(oracle/_ref/quaff align ... -format sam), with seam A served by the oracle -- both strands, so that the reverse-strand
POS arithmetic of Alignment::revcomp (qmodel.cpp:611-622, fastseq.cpp:218-230) is pinned."""
import os
import subprocess

import numpy as np
import pytest

import golden_cases as gc
import parity_cases as pc
from oracle import pyoracle as po
from quaff_b200 import api, sam
from quaff_b200.params import QuaffNullParams


def test_cigar_and_round():
    assert sam.cigar_string(np.array([0, 0, 1, 2, 2, 0], np.uint8)) == "M2I1D2M1"
    assert sam.cigar_string(np.zeros(0, np.uint8)) == ""
    assert (sam.c_round(2.5), sam.c_round(-2.5), sam.c_round(2.4999)) == (3, -3, 2)


def test_sam_fields_match_reference_cli(oracle, tmp_path):
    if not os.path.exists(po.REF_QUAFF):
        pytest.skip("oracle/_ref/quaff not built here (needs /root/reference)")
    x, reads = pc.make_workload(ref_len=6000, n_reads=6, read_len=500, seed=77)
    qp = pc.default_params()
    nullp = QuaffNullParams.load(os.path.join(gc.GOLD, "testquaffnullparams.json"))
    fa, fq, pj, nj = (str(tmp_path / n) for n in ("ref.fa", "reads.fq", "params.json", "null.json"))
    open(fa, "w").write(f">{x[0].name}\n{x[0].seq}\n")
    open(fq, "w").write("".join(f"@{r.name}\n{r.seq}\n+\n{r.qual}\n" for r in reads))
    # the SAME model text on both sides: quaff_b200.params reads JSON numbers the way the reference's parser does
    open(pj, "w").write(open(os.path.join(gc.GOLD, "defaultparams.json")).read()); open(nj, "w").write(open(os.path.join(gc.GOLD, "testquaffnullparams.json")).read())
    res = subprocess.run([po.REF_QUAFF, "align", fa, fq, "-params", pj, "-null", nj, "-kmatchn", "10", "-format", "sam", "-threads", "1"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-500:]
    # seam A from the oracle: best strand per read (earliest on ties), null-adjusted score, interval, path
    s = oracle.scores(qp)
    cfg = po.make_config(kmer_threshold=10)
    xs, ys = pc.seqbufs(x, reads)
    best, score, xst, xen, paths = [], [], [], [], []
    for m, r in enumerate(reads):
        vs = [oracle.viterbi(xs[n], ys[m], s, cfg) for n in range(len(x))]
        b = max(range(len(x)), key=lambda n: (vs[n]["result"], -n))
        best.append(b); score.append(vs[b]["result"] - api.null_loglike(nullp, r)); xst.append(vs[b]["x_start"]); xen.append(vs[b]["x_end"])
        paths.append(vs[b]["path"])
    off = np.concatenate(([0], np.cumsum([len(p) for p in paths]))).astype(np.uint64)
    result = dict(best_ref=np.array(best, np.uint32), score=np.array(score), x_start=np.array(xst, np.uint32), x_end=np.array(xen, np.uint32),
                  paths=np.concatenate(paths), path_offsets=off)
    assert len(set(best)) == 2                                   # both strands occur
    assert sam.compare_batch(res.stdout, reads, x[0].name, len(x[0]), result) == len(reads)
    # and a corrupted result is caught
    result["x_start"][0] += 1
    with pytest.raises(AssertionError):
        sam.compare_batch(res.stdout, reads, x[0].name, len(x[0]), result)


def test_gason_number_parsing():
    """JSON numbers as the reference reads them (gason string2double, src/gason.cpp:73-117): not correctly rounded"""
    from quaff_b200.params import gason_double
    assert gason_double("0.1") == 0.1 and gason_double("47") == 47.0 and gason_double("-2.5e1") == -25.0
    assert gason_double("0.3") == 3 * 0.1 and gason_double("0.3") != 0.3
    assert gason_double("1e-2") == 0.1 * 0.1 and gason_double("12.5E+1") == 125.0


def test_params_load_equals_reference_parse(ref):
    from quaff_b200.params import QuaffParams, random_params
    texts = [open(os.path.join(gc.GOLD, n)).read() for n in ("defaultparams.json", "testquaffparams.json")]
    texts.append(random_params(np.random.default_rng(1), 2, 1).to_json())
    for text in texts:
        qp = QuaffParams.from_json(text)
        hp = ref.L.qref_params_from_json(text.encode())
        qr = ref.params_as_parsed(hp, qp)
        assert np.array_equal(qp.begin_insert, qr.begin_insert) and np.array_equal(qp.begin_delete, qr.begin_delete)
        assert (qp.extend_insert, qp.extend_delete) == (qr.extend_insert, qr.extend_delete)
        for a, b in list(zip(qp.insert, qr.insert)) + [(a, b) for ra, rb in zip(qp.match, qr.match) for a, b in zip(ra, rb)]:
            assert (a.p, a.q, a.r) == (b.p, b.q, b.r)
