"""-m gpu: the CUDA library against the committed REFERENCE vectors (tests/golden, generated from the unmodified
reference by make_golden.py) -- no oracle in the loop.  Integer work, Viterbi and Forward are bit-exact;
Backward / counts are held to 1e-9 relative (the stated bar is 1e-4; only exp() differs from the host libm)."""
import json
import os

import numpy as np
import pytest

import golden_cases as gc
import parity_cases as pc
from quaff_b200 import api
from quaff_b200.seqs import FastSeq, read_fastseqs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", gc.SYNTH_CASES)
def test_gpu_matches_reference_vectors(gpu, name):
    meta, arr = gc.load_case(name)
    x, reads, qp = gc.case_inputs(meta)
    cfg = api.dp_config(**meta["cfg"])
    use_quals = all(r.has_qual() for r in reads)
    gpu.set_refs(x); gpu.set_reads(reads, use_quals=use_quals); gpu.set_params(qp)
    gpu.set_fb_exact(True)                      # the log-space kernels: Forward is bit-exact against the reference vectors
    xi = np.array([r["x"] for r in meta["pairs"]], np.uint32); yi = np.array([r["y"] for r in meta["pairs"]], np.uint32)
    env24, cu = gpu.envelopes(cfg, xi, yi, cell_size=24)
    env48, _ = gpu.envelopes(cfg, xi, yi, cell_size=48)
    v = gpu.viterbi(cfg, xi, yi)
    f = gpu.forward(cfg, xi, yi)
    b = gpu.backward_counts(cfg, xi, yi, per_pair=True) if "backward" in meta["pairs"][0] else None
    for p, rec in enumerate(meta["pairs"]):
        key = f"p{rec['x']}_{rec['y']}"
        assert np.array_equal(env24[p], arr[key + "_diag24"]) and np.array_equal(env48[p], arr[key + "_diag48"]) and cu[p] == rec["cu"]
        assert v["score"][p] == rec["viterbi"] and (v["x_start"][p], v["x_end"][p]) == (rec["x_start"], rec["x_end"])
        assert np.array_equal(v["paths"][p], gc.unrle(arr[key + "_path"]))
        assert f[p] == rec["forward"]
        if b is not None and np.isfinite(rec["forward"]):
            assert abs(b["back"][p] - rec["backward"]) <= 1e-9 * abs(rec["backward"])
            np.testing.assert_allclose(b["counts_per_pair"][p], arr[key + "_counts"], rtol=1e-9, atol=1e-12)
    if "estep" in meta:
        nullp = gc.null_from_payload(meta["null"])
        null_ll = np.array([api.null_loglike(nullp, r, gpu.L) for r in reads])
        assert list(null_ll) == meta["null_loglike"]
        so = None
        for it, e in enumerate(meta["estep"]):
            r = gpu.estep(cfg, True, null_ll, so)
            so = r["sort_order"]
            np.testing.assert_allclose(r["y_loglike"], e["loglike"], rtol=1e-13)
            assert so == e["sort_order"]
            np.testing.assert_allclose(r["counts"], arr[f"estep{it}_counts"], rtol=1e-9, atol=1e-12)


def test_gpu_overlap_vectors(gpu):
    meta, _ = gc.load_case("overlap_default")
    seqs = [FastSeq(f"s{n}", s, q) for n, (s, q) in enumerate(meta["seqs"])]
    qp = gc.params_from_payload(meta["params"])
    cfg = api.dp_config(**meta["cfg"])
    gpu.set_reads(seqs); gpu.set_overlap_params(qp)
    xi = np.array([r["x"] for r in meta["pairs"]], np.uint32); yi = np.array([r["y"] for r in meta["pairs"]], np.uint32)
    yc = np.array([r["yc"] for r in meta["pairs"]], np.uint8)
    r = gpu.overlap_viterbi(cfg, xi, yi, yc)
    for p, rec in enumerate(meta["pairs"]):
        if rec["result"] is None:
            assert np.isinf(r["score"][p])
            continue
        assert r["score"][p] == rec["result"] and list(map(int, r["coords"][p])) == rec["coords"]
        xr, yr = gpu.overlap_rows(seqs[rec["x"]].tokens(), seqs[rec["y"]].tokens(), r["coords"][p], r["paths"][p])
        assert xr == rec["xrow"] and yr == rec["yrow"]


def test_gpu_reference_own_golden_files(gpu):
    """BASELINE configs 2/3 inputs: the reference's c8f30 goldens (align 7981.84, overlap 6876.76, counts JSON) and the
    default-flag overlap (197 diagonals, score 6900.30) on the full 6604-base read."""
    with open(os.path.join(gc.GOLD, "c8f30_reference_goldens.json")) as fh:
        g = json.load(fh)
    c8 = read_fastseqs(os.path.join(gc.GOLD, "c8f30.fastq.gz"))[0]
    qp = gc.params_from_payload(g["params"]); nullp = gc.null_from_payload(g["null_fitted"])
    null_ll = api.null_loglike(nullp, c8, gpu.L)
    assert null_ll == g["null_loglike"]
    cfg = api.dp_config(kmer_threshold=-1, max_size=g["max_size"])
    gpu.set_refs([c8]); gpu.set_reads([c8]); gpu.set_params(qp)
    gpu.set_fb_exact(True)
    env, _ = gpu.envelopes(cfg, [0], [0])
    assert list(env[0]) == [0]
    a = gpu.align_reads(cfg, np.array([null_ll]))
    assert float(f"{a['score'][0]:.6g}") == g["align_score"] and (a["x_start"][0], a["x_end"][0]) == (1, len(c8))
    # quaff count
    e = gpu.estep(cfg, True, np.array([null_ll]))
    c = e["counts"]; cj = g["counts"]; off = 0; letters = "ACGT"
    def r6(v): return float(f"{v:.6g}")
    for i in range(4):
        for j in range(4):
            assert [r6(t) for t in c[off:off + 94]] == [r6(t) for t in cj["match"][""][letters[i]][letters[j]]]; off += 94
    for i in range(4):
        assert [r6(t) for t in c[off:off + 94]] == [r6(t) for t in cj["insert"][letters[i]]]; off += 94
    for key in ("beginInsertNo", "beginInsertYes", "beginDeleteNo", "beginDeleteYes"):
        assert r6(c[off]) == r6(cj[key][""]); off += 1
    for key in ("extendInsertNo", "extendInsertYes", "extendDeleteNo", "extendDeleteYes"):
        assert r6(c[off]) == r6(cj[key]); off += 1
    # quaff overlap, golden flags and default flags
    copy = FastSeq("copy", c8.seq, c8.qual)
    gpu.set_reads([c8, copy]); gpu.set_overlap_params(qp)
    null2 = api.null_loglike(gc.null_from_payload(g["null_fitted_overlap"]), c8, gpu.L)
    o = gpu.overlap_viterbi(cfg, [0], [1], [0])
    assert float(f"{o['score'][0] - 2 * null2:.6g}") == g["overlap_score"]
    cfg_d = api.dp_config(kmer_threshold=14)
    envd, _ = gpu.envelopes(cfg_d, [0], [1], x_set=api.QG_READS)
    assert len(envd[0]) == g["overlap_default"]["n_diagonals"]
    od = gpu.overlap_viterbi(cfg_d, [0], [1], [0])
    assert od["score"][0] == g["overlap_default"]["result"] and list(map(int, od["coords"][0])) == g["overlap_default"]["coords"]
    assert abs((od["score"][0] - 2 * null2) - g["overlap_default"]["score"]) < 1e-9
