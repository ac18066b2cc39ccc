"""Loader for the committed reference vectors of tests/golden/ (see make_golden.py) and engine-agnostic checks:
`engine` is either the oracle (CPU suite) or the CUDA library (GPU suite) behind a tiny adapter."""
from __future__ import annotations

import json
import os

import numpy as np

from quaff_b200.params import QuaffNullParams, QuaffParams, SymQualDist
from quaff_b200.seqs import FastSeq

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def params_from_payload(d) -> QuaffParams:
    qp = QuaffParams(match_k=d["match_k"], gap_k=d["gap_k"])
    qp.begin_insert = np.array(d["begin_insert"]); qp.begin_delete = np.array(d["begin_delete"])
    qp.extend_insert, qp.extend_delete = d["extend_insert"], d["extend_delete"]
    qp.ref_base = list(d["ref_base"])
    qp.insert = [SymQualDist(*v) for v in d["insert"]]
    qp.match = [[SymQualDist(*v) for v in row] for row in d["match"]]
    return qp


def null_from_payload(d) -> QuaffNullParams:
    return QuaffNullParams(null_emit=d["null_emit"], null=[SymQualDist(*v) for v in d["null"]])


def unrle(a) -> np.ndarray:
    a = np.asarray(a).reshape(-1, 2)
    return np.repeat(a[:, 0].astype(np.uint8), a[:, 1])


def load_case(name):
    with open(os.path.join(GOLD, name + ".json")) as fh:
        meta = json.load(fh)
    arrays = np.load(os.path.join(GOLD, name + ".npz")) if os.path.exists(os.path.join(GOLD, name + ".npz")) else {}
    return meta, arrays


def case_inputs(meta):
    x = [FastSeq(f"x{n}", s) for n, s in enumerate(meta["x"])]
    reads = [FastSeq(f"y{m}", s, q) for m, (s, q) in enumerate(meta["reads"])]
    return x, reads, params_from_payload(meta["params"])


SYNTH_CASES = ["synth_default", "synth_global_mem", "synth_order2", "synth_full"]
