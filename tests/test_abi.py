"""not-gpu: the C-ABI library loads and exports every symbol include/quaffgpu.h declares; the product refuses
to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from quaff_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "quaffgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qg_[a-z_0-9]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    build.build()                       # nvcc cross-compiles for sm_100a without a GPU
    return ctypes.CDLL(api.DEFAULT_LIB)


def test_header_and_binding_agree():
    assert header_symbols() == sorted(api.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/quaffgpu.h but not exported"
    assert lib.qg_abi_version() == 1


def test_host_helpers_need_no_device(lib):
    lib.qg_counts_size.restype = ctypes.c_size_t
    assert lib.qg_counts_size(1, 0) == 4 * 4 * 94 + 4 * 94 + 4 + 4
    assert lib.qg_counts_size(3, 2) == 4 * 64 * 94 + 4 * 94 + 4 * 16 + 4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.QuaffGpuError) as ei:
        api.QuaffGPU(device=0)
    assert ei.value.code == 4 and "no CPU path" in str(ei.value)


def test_scores_helper_matches_oracle(lib, oracle):
    """qg_scores_from_params / qg_null_loglike are host arithmetic: checked here against the oracle, bit for bit"""
    import numpy as np
    import parity_cases as pc
    from oracle import pyoracle as po
    from quaff_b200.params import QuaffNullParams, random_params
    from quaff_b200.synth import random_ref, sample_reads
    L = api.load_library()
    for qp in (pc.default_params(), random_params(np.random.default_rng(1), 3, 2)):
        a = api.scores_from_params(qp, L); b = oracle.scores(qp)
        for f in ("match", "insert", "m2m", "m2i", "m2d", "m2e"):
            assert np.array_equal(getattr(a, f), getattr(b, f))
        assert (a.d2d, a.d2m, a.i2i, a.i2m) == (b.d2d, b.d2m, b.i2i, b.i2m)
    nullp = QuaffNullParams.load(os.path.join(ROOT, "tests", "golden", "testquaffnullparams.json"))
    rd, _, _ = sample_reads(random_ref(3000, 1), 2, 500, 2)
    for r in rd:
        assert api.null_loglike(nullp, r, L) == oracle.null_loglike(nullp, po.SeqBuf(r.tokens(), r.qual_scores()))
