"""SAM fields of an alignment as the reference writes them (Alignment::writeSam / cigarString, qmodel.cpp:611-653),
computed from the flat results of seam A (best reference strand, 1-based x interval, QG_OP_* state path, adjusted score).

Used by bench.py and the tests to compare a GPU batch with the text the reference CLI prints for the same reads.

Coordinates, as the reference composes them (SeqIntervalCoords::compose, fastseq.cpp:51-64):
  forward strand : POS = xStart, FLAG = 0, CIGAR = run-length of the path in order.
  reverse strand : the alignment is reverse-complemented before printing (qmodel.cpp:613-614).  The reference row's
                   interval on the original sequence is [L - xEnd + 1, L - xStart + 1]; FastSeq::revcomp of the GAPPED row
                   (fastseq.cpp:218-230) composes {1, columns, rev} with it, so POS = (L - xStart + 1) - columns + 1 where
                   columns counts gap columns too -- the reference's own arithmetic, reproduced as is.  FLAG = 16,
                   CIGAR = run-length of the reversed path.
CIGAR items are letter-then-count ("M12I1D2", qmodel.cpp:644-649); AS:i: is C round() of the null-adjusted score.
"""
from __future__ import annotations

import math
from typing import Dict, NamedTuple

import numpy as np

_OPS = np.array(list("MID"))


class SamFields(NamedTuple):
    flag: int
    rname: str
    pos: int
    cigar: str
    score: int


def c_round(v: float) -> int:
    """C99 round(): halves away from zero"""
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def cigar_string(path: np.ndarray) -> str:
    path = np.asarray(path)
    if len(path) == 0:
        return ""
    cut = np.flatnonzero(np.diff(path)) + 1
    starts = np.concatenate(([0], cut)); ends = np.concatenate((cut, [len(path)]))
    return "".join(f"{_OPS[path[s]]}{e - s}" for s, e in zip(starts, ends))


def sam_fields(ref_name: str, ref_len: int, reverse: bool, x_start: int, x_end: int, path: np.ndarray, score: float) -> SamFields:
    path = np.asarray(path)
    if not reverse:
        return SamFields(0, ref_name, int(x_start), cigar_string(path), c_round(score))
    pos = (ref_len - int(x_start) + 1) - len(path) + 1
    return SamFields(16, ref_name, pos, cigar_string(path[::-1]), c_round(score))


def parse_sam(text: str) -> Dict[str, SamFields]:
    out = {}
    for ln in text.splitlines():
        if not ln or ln.startswith("@"):
            continue
        f = ln.split("\t")
        tag = [t for t in f[11:] if t.startswith("AS:i:")]
        out[f[0]] = SamFields(int(f[1]), f[2], int(f[3]), f[5], int(tag[0][5:]))
    return out


def compare_batch(sam_text: str, reads, ref_name: str, ref_len: int, result: dict) -> int:
    """`result`: seam A output (quaff_b200.api align_reads / pool merge) whose first len(reads) entries belong to `reads`,
    references ordered [forward, revcomp].  Raises AssertionError on the first difference; returns the reads compared."""
    want = parse_sam(sam_text)
    off = np.asarray(result["path_offsets"], dtype=np.uint64)
    n = 0
    for m, r in enumerate(reads):
        best = int(result["best_ref"][m])
        if best == 0xFFFFFFFF:
            assert r.name not in want, f"{r.name}: the reference aligned it, the GPU path found no finite score"
            continue
        assert r.name in want, f"{r.name}: missing from the reference's SAM"
        got = sam_fields(ref_name, ref_len, best == 1, int(result["x_start"][m]), int(result["x_end"][m]),
                         result["paths"][int(off[m]):int(off[m + 1])], float(result["score"][m]))
        assert got == want[r.name], f"{r.name}: GPU {got[:3]} {got.cigar[:60]}.. AS {got.score}  vs reference {want[r.name][:3]} {want[r.name].cigar[:60]}.. AS {want[r.name].score}"
        n += 1
    return n
