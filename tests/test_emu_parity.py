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


@pytest.fixture(scope="module")
def workload(emu, oracle):
    x, reads = pc.make_workload(ref_len=3000, n_reads=2, read_len=250, seed=3)
    qp = pc.default_params()
    emu.set_refs(x); emu.set_reads(reads); emu.set_params(qp)
    return x, reads, oracle.scores(qp)


def test_emu_envelopes(emu, oracle, workload):
    x, reads, _ = workload
    xi, yi = pc.all_pairs(len(x), len(reads))
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
