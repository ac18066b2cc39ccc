"""-m gpu: the CUDA library on a real B200, through the C ABI, against the oracle."""
import numpy as np
import pytest

import parity_cases as pc
from quaff_b200 import api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def workload(gpu, oracle):
    x, reads = pc.make_workload(ref_len=60000, n_reads=6, read_len=2500, seed=11)
    qp = pc.default_params()
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
