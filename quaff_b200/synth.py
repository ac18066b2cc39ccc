"""Deterministic synthetic long-read workloads.

ref  : uniform-random ACGT.
reads: windows of the reference (random strand) passed through a per-base error channel --
       substitution 5 %, deletion 4 %, geometric insertion p = 4 % (about 12 % total error) --
       with Phred qualities clip(NegBin(r=20, p=0.62), 0, 93) + 33.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .seqs import DNA, FastSeq

_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)
_LET = np.frombuffer(DNA.encode(), dtype=np.uint8)


def random_ref(length: int, seed: int, name: str = "ref") -> FastSeq:
    rng = np.random.default_rng(seed)
    tok = rng.integers(0, 4, size=length, dtype=np.uint8)
    return FastSeq(name=name, seq=_LET[tok].tobytes().decode())


def mutate(tok: np.ndarray, rng: np.random.Generator, sub=0.05, dele=0.04, ins=0.04) -> np.ndarray:
    n = len(tok)
    keep = rng.random(n) >= dele
    subm = rng.random(n) < sub
    t = tok.copy()
    t[subm] = (t[subm] + rng.integers(1, 4, size=int(subm.sum()), dtype=np.uint8)) % 4
    nins = rng.geometric(1.0 - ins, size=n) - 1          # insertions after each base
    nins[~keep] = 0
    out_len = int(keep.sum() + nins.sum())
    out = np.empty(out_len, dtype=np.uint8)
    reps = keep.astype(np.int64) + nins
    idx = np.repeat(np.arange(n), reps)
    out[:] = t[idx]
    # positions that are inserted copies get random bases
    first = np.ones(out_len, dtype=bool)
    starts = np.cumsum(reps) - reps
    is_base = np.zeros(out_len, dtype=bool)
    is_base[starts[keep & (reps > 0)]] = True
    insm = ~is_base
    out[insm] = rng.integers(0, 4, size=int(insm.sum()), dtype=np.uint8)
    return out


def random_quals(n: int, rng: np.random.Generator) -> str:
    q = np.clip(rng.negative_binomial(20, 0.62, size=n), 0, 93).astype(np.uint8) + 33
    return q.tobytes().decode("latin-1")


def sample_reads(ref: FastSeq, n_reads: int, read_len: int, seed: int, both_strands: bool = True,
                 sub=0.05, dele=0.04, ins=0.04, name_prefix: str = "read") -> Tuple[List[FastSeq], np.ndarray, np.ndarray]:
    """Returns (reads, start positions, strand flags)."""
    rng = np.random.default_rng(seed)
    rtok = ref.tokens()
    L = len(rtok)
    reads, starts, strands = [], [], []
    for n in range(n_reads):
        wl = min(read_len, L)
        s = int(rng.integers(0, L - wl + 1))
        w = rtok[s:s + wl]
        rev = bool(both_strands and rng.random() < 0.5)
        if rev:
            w = _COMP[w[::-1]]
        t = mutate(w, rng, sub, dele, ins)
        reads.append(FastSeq(name=f"{name_prefix}{n}", seq=_LET[t].tobytes().decode(), qual=random_quals(len(t), rng)))
        starts.append(s)
        strands.append(rev)
    return reads, np.array(starts), np.array(strands)
