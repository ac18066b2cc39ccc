"""not-gpu: the same kernels compiled against the CPU-thread shim (tests/emu, test infrastructure only)
checked against the oracle.  Sizes are tiny: every CUDA thread is an OS thread here."""
import numpy as np
import pytest

import parity_cases as pc
from quaff_b200 import api


@pytest.fixture(scope="module")
def emu(emu_lib):
    g = api.QuaffGPU(lib_path=emu_lib)
    yield g
    g.close()


_CACHE = {}


@pytest.fixture
def workload(emu, oracle):
    if "w" not in _CACHE:
        x, reads = pc.make_workload(ref_len=3000, n_reads=2, read_len=250, seed=3)
        _CACHE["w"] = (x, reads, pc.default_params())
    x, reads, qp = _CACHE["w"]
    emu.set_refs(x); emu.set_reads(reads); emu.set_params(qp)
    return x, reads, oracle.scores(qp)


def test_emu_envelopes(emu, oracle, workload):
    x, reads, _ = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_threshold=6), xi, yi)


def test_emu_threshold_zero(emu, oracle, workload, monkeypatch):
    """-kmatchn 0: only diagonals that received a hit are seeds (diagenv.cpp:33-46), on both seeding paths"""
    x, reads, _ = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_threshold=0, band_size=4), xi, yi)
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_len=9, kmer_threshold=0, band_size=4), xi, yi)
    monkeypatch.setenv("QG_SEED_GENERAL", "1")
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_threshold=0, band_size=4), xi, yi)


def test_emu_tile_seeding_item_boundaries(emu, oracle, workload, monkeypatch):
    """the tile-sorted seeding kernel with several work items per pair: edge tiles, interior tiles, the ring offset of items
    that do not start on a 4-diagonal boundary; fixed threshold and memory-guided mode"""
    x, reads, _ = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    monkeypatch.setenv("QG_SEED_CHUNK", "1024")
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_threshold=6), xi, yi)
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_threshold=3, kmer_len=5, band_size=10), xi, yi)
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_threshold=-1, max_size=200_000), xi, yi)
    monkeypatch.setenv("QG_SEED_LEGACY", "1")                       # the round-1 kernel stays the fallback for long reads / k = 7
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_threshold=6), xi, yi)


def test_emu_viterbi_forward(emu, oracle, workload):
    x, reads, s_or = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    cfg = api.dp_config(kmer_threshold=6)
    pc.check_viterbi(emu, oracle, x, reads, s_or, cfg, xi, yi)
    pc.check_forward(emu, oracle, x, reads, s_or, cfg, xi, yi)


def test_emu_backward_counts(emu, oracle, workload):
    x, reads, s_or = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    pc.check_backward(emu, oracle, x, reads, s_or, api.dp_config(kmer_threshold=6), xi, yi)


def test_emu_align_reads_and_estep(emu, oracle, workload):
    from quaff_b200.params import QuaffNullParams
    import os
    x, reads, s_or = workload
    nullp = QuaffNullParams.load(os.path.join(os.path.dirname(__file__), "golden", "testquaffnullparams.json"))
    cfg = api.dp_config(kmer_threshold=6)
    pc.check_estep(emu, oracle, x, reads, s_or, nullp, cfg, use_null=True, n_iter=2)
    # seam A: best reference per read, null-adjusted score
    null_ll = np.array([api.null_loglike(nullp, r, emu.L) for r in reads])
    a = emu.align_reads(cfg, null_ll)
    xs, ys = pc.seqbufs(x, reads)
    for m in range(len(reads)):
        best, bo = None, None
        for n in range(len(x)):
            o = oracle.viterbi(xs[n], ys[m], s_or, pc.oracle_cfg(cfg))
            if np.isfinite(o["result"]) and (bo is None or o["result"] > bo["result"]):
                best, bo = n, o
        assert a["best_ref"][m] == best
        assert a["score"][m] == bo["result"] - null_ll[m]
        assert a["x_start"][m] == bo["x_start"] and a["x_end"][m] == bo["x_end"]
        assert np.array_equal(a["paths"][m], bo["path"])


def test_emu_memory_guided_envelopes(emu, oracle, workload):
    """-kmatchmb: tiers accepted while the storage diagonals fit (diagenv.cpp:62-96, kmerThreshold < 0)"""
    x, reads, _ = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    for mb_bytes, cell in ((1 << 20, 24), (300_000, 24), (3_000_000, 48), (50_000, 24), (1 << 30, 24)):
        cfg = api.dp_config(kmer_threshold=-1, max_size=mb_bytes)
        pc.check_envelopes(emu, oracle, x, reads, cfg, xi, yi, cell_size=cell)


def test_emu_overlap(emu, oracle):
    from quaff_b200.synth import random_ref, sample_reads
    from quaff_b200.seqs import add_revcomps
    ref = random_ref(420, 21)
    reads, _, _ = sample_reads(ref, 3, 260, 22, both_strands=True)
    seqs = add_revcomps(reads)
    qp = pc.default_params()
    emu.set_reads(seqs); emu.set_overlap_params(qp)
    r, nf = pc.check_overlap(emu, oracle, seqs, len(reads), qp, api.dp_config(kmer_threshold=5))
    assert nf > 0


def test_emu_overlap_emission_per_cell(emu, oracle, monkeypatch):
    """K = 3 match contexts (G = 2): no emission table, the pair emission is evaluated per cell (qoverlap.cpp:50-75);
    and the same path forced at K = 1, where the table path is the comparison"""
    from quaff_b200.synth import random_ref, sample_reads
    from quaff_b200.params import random_params
    from quaff_b200.seqs import add_revcomps
    ref = random_ref(420, 31)
    reads, _, _ = sample_reads(ref, 3, 260, 32, both_strands=True)
    seqs = add_revcomps(reads)
    qp = random_params(np.random.default_rng(33), match_k=3, gap_k=2)
    emu.set_reads(seqs); emu.set_overlap_params(qp)
    r, nf = pc.check_overlap(emu, oracle, seqs, len(reads), qp, api.dp_config(kmer_threshold=5))
    assert nf > 0
    monkeypatch.setenv("QG_OVERLAP_FLY", "1")
    qp = pc.default_params()
    emu.set_overlap_params(qp)
    r, nf = pc.check_overlap(emu, oracle, seqs, len(reads), qp, api.dp_config(kmer_threshold=5))
    assert nf > 0


def test_emu_probability_space_forward_backward(emu, oracle, workload):
    """the fast (default) train kernels: probability space, block-floating exponents"""
    import os
    from quaff_b200.params import QuaffNullParams
    x, reads, s_or = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    cfg = api.dp_config(kmer_threshold=6)
    pc.check_forward(emu, oracle, x, reads, s_or, cfg, xi, yi, exact=False)
    pc.check_backward(emu, oracle, x, reads, s_or, cfg, xi, yi, exact=False)
    nullp = QuaffNullParams.load(os.path.join(os.path.dirname(__file__), "golden", "testquaffnullparams.json"))
    pc.check_estep(emu, oracle, x, reads, s_or, nullp, cfg, use_null=True, n_iter=1, exact=False)
    emu.set_fb_exact(True)


def test_emu_wide_runs_tiled(emu_lib, oracle, monkeypatch):
    """runs wider than one CTA's 8192 diagonals take the i-space tile wavefront (qg_tile.cuh): -kmatchoff on an
    8.2 kb reference gives one run of > 8192 diagonals; Viterbi (score, interval, path) and Forward stay bit-exact.
    The second half lowers the width threshold so that sparse multi-run envelopes go through the tiles as well."""
    qp = pc.default_params()
    s_or = oracle.scores(qp)
    g = api.QuaffGPU(lib_path=emu_lib)
    try:
        x, reads = pc.make_workload(ref_len=8200, n_reads=1, read_len=120, seed=11)
        g.set_refs(x); g.set_reads(reads); g.set_params(qp)
        xi, yi = pc.all_pairs(len(x), len(reads))
        xi, yi = xi[:1], yi[:1]                       # one strand is enough at this size (32 OS threads per tile)
        pc.check_viterbi(g, oracle, x, reads, s_or, api.dp_config(sparse=False), xi, yi)
        pc.check_forward(g, oracle, x, reads, s_or, api.dp_config(sparse=False, local=False), xi, yi)
        monkeypatch.setenv("QG_WIDE_MIN_DIAGS", "40")
        x, reads = pc.make_workload(ref_len=1500, n_reads=2, read_len=300, seed=5)
        g.set_refs(x); g.set_reads(reads); g.set_params(qp)
        xi, yi = pc.all_pairs(len(x), len(reads))
        for cfg in (api.dp_config(kmer_threshold=3, band_size=20), api.dp_config(kmer_threshold=6, local=False)):
            pc.check_viterbi(g, oracle, x, reads, s_or, cfg, xi, yi)
            pc.check_forward(g, oracle, x, reads, s_or, cfg, xi, yi)
            pc.check_backward(g, oracle, x, reads, s_or, cfg, xi, yi)      # Backward + counts on tiles (qg_tile_backward_kernel)
    finally:
        g.close()


def test_emu_general_seeding_path(emu, oracle, workload, monkeypatch):
    """k-mer lengths outside the shared-memory kernel's 5..7 (and reads too long for it) take the general path:
    sorted read index + per-diagonal counters in HBM.  Envelopes stay bit-exact, in both threshold modes."""
    x, reads, _ = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    for k, thr in ((4, 30), (8, 4), (11, 2), (3, 60)):
        pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_len=k, kmer_threshold=thr), xi, yi)
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_len=9, kmer_threshold=-1, max_size=400_000), xi, yi)
    monkeypatch.setenv("QG_SEED_GENERAL", "1")
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_threshold=6), xi, yi)
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_threshold=-1, max_size=300_000), xi, yi)
    monkeypatch.setenv("QG_COUNT_BUDGET_MB", "0")     # one pair per sub-batch
    pc.check_envelopes(emu, oracle, x, reads, api.dp_config(kmer_threshold=6), xi, yi)


def test_emu_viterbi_matrix_edges(emu_lib, oracle):
    """bands that run into the matrix edges (reference barely longer / shorter than the read): the phase-split Viterbi
    kernel's steady state must start and stop exactly where every real cell is inside 1 <= i <= xLen; runs of 1..4
    diagonals go through the thread-per-run kernel"""
    qp = pc.default_params()
    s_or = oracle.scores(qp)
    g = api.QuaffGPU(lib_path=emu_lib)
    try:
        for ref_len, read_len, seed in ((330, 300, 41), (260, 300, 42), (600, 120, 43)):
            x, reads = pc.make_workload(ref_len=ref_len, n_reads=2, read_len=read_len, seed=seed)
            g.set_refs(x); g.set_reads(reads); g.set_params(qp)
            xi, yi = pc.all_pairs(len(x), len(reads))
            for cfg in (api.dp_config(kmer_threshold=3, band_size=40), api.dp_config(kmer_threshold=2, band_size=2),
                        api.dp_config(kmer_threshold=4, band_size=90, local=False)):
                pc.check_viterbi(g, oracle, x, reads, s_or, cfg, xi, yi)
    finally:
        g.close()


def test_emu_pool_with_more_contexts_than_reads(emu_lib, oracle):
    """QuaffGPUPool: contexts whose share of the reads is empty contribute empty results"""
    x, reads = pc.make_workload(ref_len=2000, n_reads=1, read_len=200, seed=3)
    qp = pc.default_params()
    cfg = api.dp_config(kmer_threshold=6)
    g = api.QuaffGPU(lib_path=emu_lib)
    g.set_refs(x); g.set_reads(reads); g.set_params(qp)
    one = g.align_reads(cfg, np.zeros(len(reads)))
    g.close()
    P = api.QuaffGPUPool(n_ctx=3, lib_path=emu_lib)
    try:
        P.set_refs(x); P.set_params(qp); P.set_reads(reads)
        a = P.align_reads(cfg, np.zeros(len(reads)))
        assert np.array_equal(a["score"], one["score"]) and np.array_equal(a["best_ref"], one["best_ref"])
        flat = [api._flatten(reads, True)]
        P.set_read_batches(flat)
        b = P.align_batches(cfg, [(0, np.zeros(len(reads)))])[0]
        c = P.align_stream(cfg, [flat[0] + (np.zeros(len(reads)),)])[0]
        assert np.array_equal(b["score"], one["score"]) and np.array_equal(c["score"], one["score"])
        assert np.array_equal(np.concatenate(one["paths"]) if isinstance(one["paths"], list) else one["paths"], b["paths"])
    finally:
        P.close()


def test_emu_pool_matches_single_context(emu, emu_lib, oracle):
    """qg_pool_*: chunks of reads over two contexts (host threads inside the library), results in read order; E-step count
    sum over contiguous read ranges -- both against the single-context entry points"""
    from quaff_b200.params import QuaffNullParams
    import os
    x, reads = pc.make_workload(ref_len=3000, n_reads=5, read_len=220, seed=13)
    qp = pc.default_params()
    nullp = QuaffNullParams.load(os.path.join(os.path.dirname(__file__), "golden", "testquaffnullparams.json"))
    cfg = api.dp_config(kmer_threshold=6)
    null_ll = np.array([api.null_loglike(nullp, r, emu.L) for r in reads])
    emu.set_refs(x); emu.set_reads(reads); emu.set_params(qp)
    one = emu.align_reads(cfg, null_ll, split_paths=False)
    P = api.QuaffPool(devices=[0], contexts_per_device=2, lib_path=emu_lib)
    try:
        assert P.size() == 2
        P.set_refs(x); P.set_params(qp)
        many = P.align_reads(cfg, reads, null_ll, chunk_reads=2)          # 3 chunks over 2 contexts
        for k in ("best_ref", "score", "x_start", "x_end", "paths", "path_offsets"):
            assert np.array_equal(np.asarray(one[k]), np.asarray(many[k])), k
        emu.set_fb_exact(True); P.set_fb_exact(True)
        e1 = emu.estep(cfg, True, null_ll)
        e2 = P.estep(cfg, True, reads, null_ll)
        assert e1["sort_order"] == e2["sort_order"] and np.array_equal(e1["y_loglike"], e2["y_loglike"])
        np.testing.assert_allclose(e2["counts"], e1["counts"], rtol=1e-12, atol=1e-300)
        assert abs(e1["loglike"] - e2["loglike"]) <= 1e-12 * abs(e1["loglike"])
    finally:
        P.close()


def test_emu_forward_backward_kernel_choice_per_pair(emu, oracle, workload):
    """a call whose pairs need different kernels (bands wider than 256 diagonals next to narrow runs) is split per pair:
    probability-space kernels for the pairs that fit one warp, log-space ones for the others -- each within its tolerance,
    whatever its batch-mates are"""
    x, reads, s_or = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
    cfg = api.dp_config(kmer_threshold=6, band_size=300)
    d, _ = emu.envelopes(cfg, xi, yi, cell_size=48)
    widths = [len(v) for v in d]
    assert min(widths) <= 256 < max(widths), widths                # both kinds present
    pc.check_forward(emu, oracle, x, reads, s_or, cfg, xi, yi, exact=False)
    pc.check_backward(emu, oracle, x, reads, s_or, cfg, xi, yi, exact=False)
