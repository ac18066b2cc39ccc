"""-m gpu: the CUDA library on a real B200, through the C ABI, against the oracle."""
import os

import numpy as np
import pytest

import parity_cases as pc
from quaff_b200 import api

pytestmark = pytest.mark.gpu


_CACHE = {}


@pytest.fixture
def workload(gpu, oracle):
    """the context is shared by the whole session: (re)upload this module's inputs before every test"""
    if "w" not in _CACHE:
        x, reads = pc.make_workload(ref_len=60000, n_reads=6, read_len=2500, seed=11)
        _CACHE["w"] = (x, reads, pc.default_params())
    x, reads, qp = _CACHE["w"]
    gpu.set_refs(x); gpu.set_reads(reads); gpu.set_params(qp)
    return x, reads, oracle.scores(qp)


def test_envelopes_bit_exact(gpu, oracle, workload):
    x, reads, _ = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    for thr in (20, 14, 6):
        pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(kmer_threshold=thr), xi, yi)
    pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(kmer_threshold=10, band_size=20, kmer_len=5), xi, yi)
    pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(kmer_threshold=10, band_size=33, kmer_len=7), xi, yi)


def test_viterbi_bit_exact(gpu, oracle, workload):
    x, reads, s_or = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    pc.check_viterbi(gpu, oracle, x, reads, s_or, api.dp_config(kmer_threshold=20), xi, yi)
    pc.check_viterbi(gpu, oracle, x, reads, s_or, api.dp_config(kmer_threshold=20, local=False), xi, yi)


def test_forward_bit_exact(gpu, oracle, workload):
    x, reads, s_or = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    pc.check_forward(gpu, oracle, x, reads, s_or, api.dp_config(kmer_threshold=20), xi, yi)


def test_full_dp_small(gpu, oracle):
    """-kmatchoff on short sequences: one wide run, the multi-warp kernel"""
    x, reads = pc.make_workload(ref_len=700, n_reads=3, read_len=300, seed=5)
    qp = pc.default_params()
    gpu.set_refs(x); gpu.set_reads(reads); gpu.set_params(qp)
    s_or = oracle.scores(qp)
    xi, yi = pc.all_pairs(len(x), len(reads))
    cfg = api.dp_config(sparse=False)
    pc.check_envelopes(gpu, oracle, x, reads, cfg, xi, yi)
    pc.check_viterbi(gpu, oracle, x, reads, s_or, cfg, xi, yi)
    pc.check_forward(gpu, oracle, x, reads, s_or, cfg, xi, yi)


def test_backward_counts(gpu, oracle, workload):
    x, reads, s_or = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    pc.check_backward(gpu, oracle, x, reads, s_or, api.dp_config(kmer_threshold=20), xi, yi)


def test_estep_and_align_reads(gpu, oracle, workload):
    import os
    from quaff_b200.params import QuaffNullParams
    x, reads, s_or = workload
    nullp = QuaffNullParams.load(os.path.join(os.path.dirname(__file__), "golden", "testquaffnullparams.json"))
    cfg = api.dp_config(kmer_threshold=20)
    pc.check_estep(gpu, oracle, x, reads, s_or, nullp, cfg, use_null=True, n_iter=2)
    pc.check_estep(gpu, oracle, x, reads, s_or, nullp, cfg, use_null=False, n_iter=1)
    null_ll = np.array([api.null_loglike(nullp, r, gpu.L) for r in reads])
    a = gpu.align_reads(cfg, null_ll)
    xs, ys = pc.seqbufs(x, reads)
    for m in range(len(reads)):
        best, bo = None, None
        for n in range(len(x)):
            o = oracle.viterbi(xs[n], ys[m], s_or, pc.oracle_cfg(cfg))
            if np.isfinite(o["result"]) and (bo is None or o["result"] > bo["result"]):
                best, bo = n, o
        assert a["best_ref"][m] == best and a["score"][m] == bo["result"] - null_ll[m]
        assert a["x_start"][m] == bo["x_start"] and a["x_end"][m] == bo["x_end"] and np.array_equal(a["paths"][m], bo["path"])


def test_memory_guided_envelopes(gpu, oracle, workload):
    x, reads, _ = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    for nbytes, cell in ((10 << 20, 24), (2_000_000, 24), (20_000_000, 48), (400_000, 24)):
        pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(kmer_threshold=-1, max_size=nbytes), xi, yi, cell_size=cell)


def test_order2_params(gpu, oracle):
    """-order 2: k-mer context for substitutions (K = 3) and gap opening (G = 2)"""
    from quaff_b200.params import random_params
    x, reads = pc.make_workload(ref_len=20000, n_reads=3, read_len=1500, seed=7)
    qp = random_params(np.random.default_rng(5), match_k=3, gap_k=2)
    gpu.set_refs(x); gpu.set_reads(reads); gpu.set_params(qp)
    s_or = oracle.scores(qp)
    xi, yi = pc.all_pairs(len(x), len(reads))
    cfg = api.dp_config(kmer_threshold=14)
    pc.check_viterbi(gpu, oracle, x, reads, s_or, cfg, xi, yi)
    pc.check_forward(gpu, oracle, x, reads, s_or, cfg, xi, yi)
    pc.check_backward(gpu, oracle, x, reads, s_or, cfg, xi, yi)


def test_noquals(gpu, oracle):
    x, reads = pc.make_workload(ref_len=20000, n_reads=3, read_len=1200, seed=9)
    qp = pc.default_params()
    gpu.set_refs(x); gpu.set_reads(reads, use_quals=False); gpu.set_params(qp)
    s_or = oracle.scores(qp)
    xi, yi = pc.all_pairs(len(x), len(reads))
    cfg = api.dp_config(kmer_threshold=14)
    pc.check_viterbi(gpu, oracle, x, reads, s_or, cfg, xi, yi, use_quals=False)
    pc.check_forward(gpu, oracle, x, reads, s_or, cfg, xi, yi, use_quals=False)


def test_overlap(gpu, oracle):
    from quaff_b200.synth import random_ref, sample_reads
    from quaff_b200.seqs import add_revcomps
    ref = random_ref(4000, 21)
    reads, _, _ = sample_reads(ref, 5, 2200, 22, both_strands=True)
    seqs = add_revcomps(reads)
    qp = pc.default_params()
    gpu.set_reads(seqs); gpu.set_overlap_params(qp)
    r, nf = pc.check_overlap(gpu, oracle, seqs, len(reads), qp, api.dp_config(kmer_threshold=14))
    assert nf >= 5
    r, nf = pc.check_overlap(gpu, oracle, seqs, len(reads), qp, api.dp_config(kmer_threshold=14), use_quals=True)


def test_probability_space_forward_backward(gpu, oracle, workload):
    """the default train kernels (probability space, reference-compatible log-sum-exp cut-off) against the oracle:
    Forward within 1e-5 nats (bar 1e-3), counts within 1e-6 relative (bar 1e-4), same gate / sortOrder decisions"""
    import os
    from quaff_b200.params import QuaffNullParams
    x, reads, s_or = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    cfg = api.dp_config(kmer_threshold=20)
    pc.check_forward(gpu, oracle, x, reads, s_or, cfg, xi, yi, exact=False)
    pc.check_backward(gpu, oracle, x, reads, s_or, cfg, xi, yi, exact=False)
    nullp = QuaffNullParams.load(os.path.join(os.path.dirname(__file__), "golden", "testquaffnullparams.json"))
    pc.check_estep(gpu, oracle, x, reads, s_or, nullp, cfg, use_null=True, n_iter=2, exact=False)
    gpu.set_fb_exact(True)


def test_probability_space_order2_and_global(gpu, oracle):
    from quaff_b200.params import random_params
    x, reads = pc.make_workload(ref_len=20000, n_reads=3, read_len=1500, seed=7)
    qp = random_params(np.random.default_rng(5), match_k=3, gap_k=2)
    gpu.set_refs(x); gpu.set_reads(reads); gpu.set_params(qp)
    s_or = oracle.scores(qp)
    xi, yi = pc.all_pairs(len(x), len(reads))
    for cfg in (api.dp_config(kmer_threshold=14), api.dp_config(kmer_threshold=14, local=False)):
        pc.check_forward(gpu, oracle, x, reads, s_or, cfg, xi, yi, exact=False)
        pc.check_backward(gpu, oracle, x, reads, s_or, cfg, xi, yi, exact=False)
    gpu.set_fb_exact(True)


def test_wide_runs_tiled(gpu, oracle):
    """-kmatchoff on a 9 kb reference: one run of > 8192 diagonals per pair goes through the i-space tile
    wavefront (qg_tile.cuh); Viterbi score / interval / path and Forward stay bit-exact, Backward + counts (tiles over the
    reversed matrix) within 1e-9, and the E-step (count / train -kmatchoff) runs on top of them"""
    x, reads = pc.make_workload(ref_len=9000, n_reads=2, read_len=1500, seed=23)
    qp = pc.default_params()
    gpu.set_refs(x); gpu.set_reads(reads); gpu.set_params(qp)
    s_or = oracle.scores(qp)
    xi, yi = pc.all_pairs(len(x), len(reads))
    for local in (True, False):
        cfg = api.dp_config(sparse=False, local=local)
        pc.check_viterbi(gpu, oracle, x, reads, s_or, cfg, xi, yi)
        pc.check_forward(gpu, oracle, x, reads, s_or, cfg, xi, yi)
        pc.check_backward(gpu, oracle, x, reads, s_or, cfg, xi, yi)
    from quaff_b200.params import QuaffNullParams
    nullp = QuaffNullParams.load(os.path.join(os.path.dirname(__file__), "golden", "testquaffnullparams.json"))
    pc.check_estep(gpu, oracle, x, reads, s_or, nullp, api.dp_config(sparse=False), n_iter=1)


def test_wide_and_narrow_pairs_mixed(gpu, oracle, workload, monkeypatch):
    """lowering the width threshold sends some pairs of a call through the tiles and the rest through the
    warp-per-run kernels: align_reads (best reference per read) must not change"""
    x, reads, s_or = workload
    cfg = api.dp_config(kmer_threshold=14)
    null_ll = np.zeros(len(reads))
    a0 = gpu.align_reads(cfg, null_ll)
    xi, yi = pc.all_pairs(len(x), len(reads))
    diags, _ = gpu.envelopes(cfg, xi, yi)

    def widest_run(d):
        cuts = np.flatnonzero(np.diff(d) != 1)
        return int(np.diff(np.concatenate(([-1], cuts, [len(d) - 1]))).max()) if len(d) else 0
    widths = sorted(widest_run(d) for d in diags)
    assert widths[0] < widths[-1]
    monkeypatch.setenv("QG_WIDE_MIN_DIAGS", str((widths[0] + widths[-1]) // 2))   # pairs on both sides of the threshold
    a1 = gpu.align_reads(cfg, null_ll)
    f1 = pc.check_forward(gpu, oracle, x, reads, s_or, cfg, xi, yi)
    pc.check_backward(gpu, oracle, x, reads, s_or, cfg, xi, yi)           # wide and banded pairs in one call: tiles + the banded kernels
    monkeypatch.delenv("QG_WIDE_MIN_DIAGS")
    assert np.array_equal(a0["best_ref"], a1["best_ref"]) and np.array_equal(a0["score"], a1["score"])
    assert np.array_equal(a0["x_start"], a1["x_start"]) and np.array_equal(a0["x_end"], a1["x_end"])
    for p0, p1 in zip(a0["paths"], a1["paths"]):
        assert np.array_equal(p0, p1)
    assert np.isfinite(f1).any()


def test_general_seeding_path(gpu, oracle, workload, monkeypatch):
    """-kmatch k outside 5..7 and reads beyond the shared-memory kernel's reach: sorted read index (CUB segmented
    radix sort) + HBM counters; envelopes bit-exact, then Viterbi on top of them"""
    x, reads, s_or = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    for k, thr in ((4, 60), (8, 6), (12, 3), (20, 1), (32, 1)):
        pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(kmer_len=k, kmer_threshold=thr), xi, yi)
    pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(kmer_len=9, kmer_threshold=-1, max_size=20_000_000), xi, yi)
    pc.check_viterbi(gpu, oracle, x, reads, s_or, api.dp_config(kmer_len=10, kmer_threshold=4), xi, yi)
    monkeypatch.setenv("QG_SEED_GENERAL", "1")
    pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(kmer_threshold=14), xi, yi)
    monkeypatch.setenv("QG_COUNT_BUDGET_MB", "1")
    pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(kmer_threshold=14), xi, yi)


def test_long_read_seeding(gpu, oracle):
    """a 70 kb read does not fit the shared-memory index: general path, bit-exact envelope"""
    x, reads = pc.make_workload(ref_len=200000, n_reads=1, read_len=70000, seed=31)
    gpu.set_refs(x); gpu.set_reads(reads); gpu.set_params(pc.default_params())
    xi, yi = pc.all_pairs(len(x), len(reads))
    pc.check_envelopes(gpu, oracle, x, reads, api.dp_config(), xi, yi)
    v = gpu.viterbi(api.dp_config(), xi, yi)
    assert np.isfinite(v["score"]).any()


def test_viterbi_matrix_edges_and_narrow_runs(gpu, oracle):
    """bands clipped by the matrix edges and runs of 1..4 diagonals (thread-per-run kernel): bit-exact paths"""
    qp = pc.default_params()
    s_or = oracle.scores(qp)
    for ref_len, read_len, seed in ((4300, 4000, 41), (3500, 4000, 42), (9000, 1500, 43)):
        x, reads = pc.make_workload(ref_len=ref_len, n_reads=2, read_len=read_len, seed=seed)
        gpu.set_refs(x); gpu.set_reads(reads); gpu.set_params(qp)
        xi, yi = pc.all_pairs(len(x), len(reads))
        for cfg in (api.dp_config(kmer_threshold=8, band_size=40), api.dp_config(kmer_threshold=5, band_size=2),
                    api.dp_config(kmer_threshold=10, band_size=90, local=False)):
            pc.check_viterbi(gpu, oracle, x, reads, s_or, cfg, xi, yi)


def test_empty_pair_lists_are_noops(gpu, oracle, workload):
    """zero pairs: every entry point returns empty results (empty grids are skipped, not launch errors)"""
    e = np.zeros(0, np.uint32)
    cfg = api.dp_config(kmer_threshold=14)
    d, cu = gpu.envelopes(cfg, e, e)
    assert len(d) == 0 and len(cu) == 0
    v = gpu.viterbi(cfg, e, e)
    assert len(v["score"]) == 0 and len(v["paths"]) == 0
    assert len(gpu.forward(cfg, e, e)) == 0
    b = gpu.backward_counts(cfg, e, e)
    assert np.all(np.asarray(b["counts"]) == 0) if "counts" in b else True
    a = gpu.align_reads(cfg, np.zeros(0), first=0, count=0)
    assert len(a["score"]) == 0


def test_forward_backward_kernel_choice_per_pair(gpu, oracle, workload):
    """a call whose pairs need different kernels (bands wider than 256 diagonals next to narrow runs) is split per pair:
    probability-space kernels for the pairs that fit one warp, log-space ones for the others -- each within its tolerance,
    whatever its batch-mates are; the E-step on top of such a call takes its two-pass form"""
    x, reads, s_or = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    cfg = api.dp_config(kmer_threshold=14, band_size=300)
    d, _ = gpu.envelopes(cfg, xi, yi, cell_size=48)
    widths = [len(v) for v in d]
    assert min(widths) <= 256 < max(widths), widths
    pc.check_forward(gpu, oracle, x, reads, s_or, cfg, xi, yi, exact=False)
    pc.check_backward(gpu, oracle, x, reads, s_or, cfg, xi, yi, exact=False)
    from quaff_b200.params import QuaffNullParams
    nullp = QuaffNullParams.load(os.path.join(os.path.dirname(__file__), "golden", "testquaffnullparams.json"))
    pc.check_estep(gpu, oracle, x, reads, s_or, nullp, cfg, n_iter=1, exact=False)
    gpu.set_fb_exact(True)


def test_pool_matches_single_context(gpu, oracle):
    """qg_pool_*: chunks of reads over three contexts on the device (host threads inside the library), results in read order;
    E-step count sum over contiguous read ranges -- both against the single-context entry points"""
    from quaff_b200.params import QuaffNullParams
    x, reads = pc.make_workload(ref_len=40000, n_reads=11, read_len=1200, seed=31)
    qp = pc.default_params()
    nullp = QuaffNullParams.load(os.path.join(os.path.dirname(__file__), "golden", "testquaffnullparams.json"))
    cfg = api.dp_config(kmer_threshold=14)
    null_ll = np.array([api.null_loglike(nullp, r, gpu.L) for r in reads])
    gpu.set_refs(x); gpu.set_reads(reads); gpu.set_params(qp)
    one = gpu.align_reads(cfg, null_ll, split_paths=False)
    P = api.QuaffPool(devices=[0], contexts_per_device=3)
    try:
        assert P.size() == 3
        P.set_refs(x); P.set_params(qp)
        many = P.align_reads(cfg, reads, null_ll, chunk_reads=2)          # 6 chunks over 3 contexts
        for k in ("best_ref", "score", "x_start", "x_end", "paths", "path_offsets"):
            assert np.array_equal(np.asarray(one[k]), np.asarray(many[k])), k
        gpu.set_fb_exact(True); P.set_fb_exact(True)
        e1 = gpu.estep(cfg, True, null_ll)
        e2 = P.estep(cfg, True, reads, null_ll)
        assert e1["sort_order"] == e2["sort_order"] and np.array_equal(e1["y_loglike"], e2["y_loglike"])
        np.testing.assert_allclose(e2["counts"], e1["counts"], rtol=1e-12, atol=1e-300)
        assert abs(e1["loglike"] - e2["loglike"]) <= 1e-12 * abs(e1["loglike"])
    finally:
        P.close()
