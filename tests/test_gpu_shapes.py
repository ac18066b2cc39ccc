"""-m gpu: parity at the shapes BASELINE.json names (not scaled-down stand-ins).

  cfg4  8 kb reads vs the 5 Mb seed-1 reference, both strands, -kmatch 6 -kmatchn 20 -kmatchband 64 (the bench workload):
        envelopes / Viterbi score, interval and path bit-exact, probability-space Forward within 1e-3 nats, E-step counts
        within 1e-4 relative -- against the oracle on the same seeded inputs (the oracle takes ~6 s per pair here).
  cfg5  one 10 kb read vs the 50 kb seed-3 reference, -kmatchoff -fwdstrand (5e8 cells, tiled wavefront): against the
        committed reference vector tests/golden/cfg5_full.* (make_golden.py cfg5; the reference needs 12 GB for it).
  cfg3  `quaff train tiny.fasta c8f30.fastq.gz -ref c8f30.fastq.gz -order 2 -force -maxiter 4` through the patched CLI
        against the unmodified reference CLI (the non-degenerate form of config 3).
  overlap with K = 2 match contexts and G = 1 gap contexts on the device.
"""
import json
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import golden_cases as gc
import parity_cases as pc
from quaff_b200 import api
from quaff_b200.params import QuaffNullParams, random_params
from quaff_b200.seqs import add_revcomps
from quaff_b200.synth import random_ref, sample_reads

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFQ = os.path.join(ROOT, "oracle", "_ref", "quaff")
GPUQ = os.path.join(ROOT, "host", "_build", "quaff-gpu")

CFG4_LL_TOL = 1e-3        # nats per pair: the tolerance north_star states for Forward log-likelihoods
CFG4_COUNT_RTOL = 1e-4    # relative: the tolerance north_star states for E-step counts


@pytest.fixture(scope="module")
def cfg4():
    ref = random_ref(5_000_000, 1)                                    # bench.py's reference (seed 1)
    reads, _, _ = sample_reads(ref, 4, 8000, 2, name_prefix="r")      # bench.py's first reads (seed 2)
    return add_revcomps([ref]), reads


def test_cfg4_envelopes_and_viterbi_bit_exact(gpu, oracle, cfg4):
    x, reads = cfg4
    qp = pc.default_params()
    gpu.set_refs(x); gpu.set_reads(reads); gpu.set_params(qp)
    s_or = oracle.scores(qp)
    xi, yi = pc.all_pairs(len(x), len(reads))
    cfg = api.dp_config(kmer_threshold=20, band_size=64, kmer_len=6)
    cu = pc.check_envelopes(gpu, oracle, x, reads, cfg, xi, yi)
    assert cu.min() > 8000 and cu.max() > 1_000_000                   # both the diagonal-0-only and the banded kind occur
    pc.check_viterbi(gpu, oracle, x, reads, s_or, cfg, xi, yi)
    # seam A on the same reads: best strand, null-adjusted score, interval and path of the traced pair
    nullp = QuaffNullParams.load(os.path.join(gc.GOLD, "testquaffnullparams.json"))
    null_ll = np.array([api.null_loglike(nullp, r, gpu.L) for r in reads])
    a = gpu.align_reads(cfg, null_ll)
    xs, ys = pc.seqbufs(x, reads)
    for m in range(len(reads)):
        os_ = [oracle.viterbi(xs[n], ys[m], s_or, pc.oracle_cfg(cfg)) for n in range(len(x))]
        best = max(range(len(x)), key=lambda n: (os_[n]["result"], -n))
        assert a["best_ref"][m] == best and a["score"][m] == os_[best]["result"] - null_ll[m]
        assert (a["x_start"][m], a["x_end"][m]) == (os_[best]["x_start"], os_[best]["x_end"])
        assert np.array_equal(a["paths"][m], os_[best]["path"])


def test_cfg4_forward_backward_estep(gpu, oracle, cfg4):
    """QG_LSE_CUT over 8000 rows, the closed-form isolated-diagonal Backward and the stored-Forward path at full length"""
    x, reads = cfg4
    qp = pc.default_params()
    gpu.set_refs(x); gpu.set_reads(reads); gpu.set_params(qp)
    s_or = oracle.scores(qp)
    xi, yi = pc.all_pairs(len(x), len(reads))
    cfg = api.dp_config(kmer_threshold=20, band_size=64, kmer_len=6)
    xs, ys = pc.seqbufs(x, reads)
    oc = pc.oracle_cfg(cfg)
    ob = [oracle.backward(xs[a], ys[b], s_or, oc) for a, b in zip(xi, yi)]
    # log-space kernels: Forward bit-exact, Backward / counts 1e-9
    gpu.set_fb_exact(True)
    f = gpu.forward(cfg, xi, yi)
    for p, o in enumerate(ob):
        assert o["fwd"] == f[p], (p, o["fwd"], f[p])
    # probability-space kernels (default): stated tolerances
    gpu.set_fb_exact(False)
    f = gpu.forward(cfg, xi, yi)
    r = gpu.backward_counts(cfg, xi, yi, per_pair=True)
    worst_ll = 0.0; worst_c = 0.0
    for p, o in enumerate(ob):
        worst_ll = max(worst_ll, abs(o["fwd"] - f[p]), abs(o["fwd"] - r["fwd"][p]), abs(o["back"] - r["back"][p]))
        scale = max(1.0, float(np.abs(o["counts"]).max()))
        worst_c = max(worst_c, float(np.max(np.abs(r["counts_per_pair"][p] - o["counts"]) / (np.abs(o["counts"]) + 1e-6 * scale))))
        np.testing.assert_allclose(r["counts_per_pair"][p], o["counts"], rtol=CFG4_COUNT_RTOL, atol=1e-6 * scale)
    print(f"cfg4 probability-space: worst |dLL| = {worst_ll:.3e} nats, worst relative count error = {worst_c:.3e}")
    assert worst_ll <= CFG4_LL_TOL
    nullp = QuaffNullParams.load(os.path.join(gc.GOLD, "testquaffnullparams.json"))
    null_ll = np.array([api.null_loglike(nullp, rd, gpu.L) for rd in reads])
    g = gpu.estep(cfg, True, null_ll)
    o = oracle.estep(xs, ys, s_or, nullp, True, oc, None)
    np.testing.assert_allclose(g["y_loglike"], o["loglike"], rtol=0, atol=CFG4_LL_TOL)
    assert g["sort_order"] == o["sort_order"]
    np.testing.assert_allclose(g["counts"], o["counts"], rtol=CFG4_COUNT_RTOL, atol=1e-6 * max(1.0, float(np.abs(o["counts"]).max())))
    gpu.set_fb_exact(True)


def test_cfg5_full_dp_reference_vector(gpu):
    meta, arr = gc.load_case("cfg5_full")
    ref = random_ref(meta["ref_len"], meta["ref_seed"])
    reads, _, _ = sample_reads(ref, 1, meta["read_len"], meta["read_seed"], both_strands=False)
    assert len(reads[0]) == meta["read_actual_len"]
    gpu.set_refs([ref]); gpu.set_reads(reads); gpu.set_params(gc.params_from_payload(meta["params"]))   # as the reference parsed them
    cfg = api.dp_config(**meta["cfg"])
    v = gpu.viterbi(cfg, [0], [0])
    assert v["score"][0] == meta["viterbi"] and (int(v["x_start"][0]), int(v["x_end"][0])) == (meta["x_start"], meta["x_end"])
    assert np.array_equal(v["paths"][0], gc.unrle(arr["path"])) and len(v["paths"][0]) == meta["path_len"]
    gpu.set_fb_exact(True)
    f = gpu.forward(cfg, [0], [0])
    assert f[0] == meta["forward"]


def test_overlap_order2_contexts(gpu, oracle):
    """K = 2 match contexts, G = 1 gap contexts: the device emission table (16^K x 94^2) and the [gx][gy] transitions"""
    ref = random_ref(3000, 51)
    reads, _, _ = sample_reads(ref, 4, 1500, 52, both_strands=True)
    seqs = add_revcomps(reads)
    qp = random_params(np.random.default_rng(53), match_k=2, gap_k=1)
    gpu.set_reads(seqs); gpu.set_overlap_params(qp)
    r, nf = pc.check_overlap(gpu, oracle, seqs, len(reads), qp, api.dp_config(kmer_threshold=10))
    assert nf >= 3


def test_overlap_order3_emission_per_cell(gpu, oracle, monkeypatch):
    """K = 3, G = 2: above the table limit the pair emission is evaluated per cell from per-position factors, with the
    fold order of qoverlap.cpp:60-72 (same bits as a table entry); the forced per-cell path at K = 2 must agree too"""
    ref = random_ref(3000, 71)
    reads, _, _ = sample_reads(ref, 4, 1500, 72, both_strands=True)
    seqs = add_revcomps(reads)
    qp = random_params(np.random.default_rng(73), match_k=3, gap_k=2)
    gpu.set_reads(seqs); gpu.set_overlap_params(qp)
    r, nf = pc.check_overlap(gpu, oracle, seqs, len(reads), qp, api.dp_config(kmer_threshold=10))
    assert nf >= 3
    monkeypatch.setenv("QG_OVERLAP_FLY", "1")
    qp = random_params(np.random.default_rng(74), match_k=2, gap_k=1)
    gpu.set_overlap_params(qp)
    r, nf = pc.check_overlap(gpu, oracle, seqs, len(reads), qp, api.dp_config(kmer_threshold=10))
    assert nf >= 3


def test_threshold_zero_envelopes(gpu, oracle, monkeypatch):
    """-kmatchn 0: every diagonal with at least one hit is a seed, none without (diagenv.cpp:33-46) -- on the
    shared-memory kernel and on the general path"""
    x, reads = pc.make_workload(ref_len=20000, n_reads=2, read_len=1200, seed=61)
    gpu.set_refs(x); gpu.set_reads(reads); gpu.set_params(pc.default_params())
    xi, yi = pc.all_pairs(len(x), len(reads))
    pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(kmer_threshold=0, band_size=4), xi, yi)
    pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(kmer_len=9, kmer_threshold=0, band_size=4), xi, yi)
    monkeypatch.setenv("QG_SEED_GENERAL", "1")
    pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(kmer_threshold=0, band_size=4), xi, yi)


def _run(binary, args, env=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    res = subprocess.run([binary] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e, timeout=1200, cwd=tempfile.gettempdir())
    assert res.returncode == 0, res.stderr[-800:]
    return res.stdout, res.stderr


def test_cfg3_train_order2_cli(tmp_path):
    if not (os.path.exists(REFQ) and os.path.exists(GPUQ)):
        pytest.skip("host/_build/quaff-gpu or oracle/_ref/quaff not built (they are built where /root/reference exists)")
    from test_host_cli import _assert_params_close
    tiny, c8 = os.path.join(gc.GOLD, "tiny.fasta"), os.path.join(gc.GOLD, "c8f30.fastq.gz")
    args = ["train", tiny, c8, "-ref", c8, "-order", "2", "-force", "-maxiter", "4", "-v2"]
    a_out, a_err = _run(REFQ, args + ["-threads", "1"])
    pat = re.compile(r"EM iteration (\d+): log-likelihood \((\S+)\)")
    a_ll = [float(m.group(2)) for m in pat.finditer(a_err)]
    assert len(a_ll) == 4 and abs(a_ll[0] + 26755.5) < 0.1 and abs(a_ll[-1] + 17415.6) < 0.1
    for env in ({"QUAFF_GPU_EXACT": "1"}, {}):
        b_out, b_err = _run(GPUQ, args + ["-gpu"], env=env)
        b_ll = [float(m.group(2)) for m in pat.finditer(b_err)]
        assert len(b_ll) == 4
        for u, v in zip(a_ll, b_ll):
            assert abs(u - v) <= 1e-4 * abs(u), (a_ll, b_ll)
        _assert_params_close(a_out, b_out, 1e-4)
