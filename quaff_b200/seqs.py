"""Sequence container and the three per-position encodings the DP consumes.

Host-side mirror of the reference's FastSeq (src/fastseq.h:41-74, src/fastseq.cpp:11-16, 71-109,
148-176, 206-238).  Only what the GPU path needs: ACGT tokens, Phred+33 quality scores clamped
to 0..93, reverse complement, and a FASTA/FASTQ(.gz) reader.  Unknown symbols are an error, as
in the reference (fastseq.cpp:76-79).
"""
from __future__ import annotations

import gzip
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

DNA = "ACGT"
QUAL_RANGE = 94          # FastSeq::qualScoreRange
_TOK = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(DNA):
    _TOK[ord(_c)] = _i
    _TOK[ord(_c.lower())] = _i


@dataclass
class FastSeq:
    name: str
    seq: str
    qual: str = ""
    comment: str = ""
    # provenance for reverse complements (SeqIntervalCoords, fastseq.h:29-39)
    source_name: str = ""
    source_rev: bool = False
    _tok: Optional[np.ndarray] = field(default=None, repr=False, compare=False)
    _q: Optional[np.ndarray] = field(default=None, repr=False, compare=False)

    def __len__(self) -> int:
        return len(self.seq)

    def has_qual(self) -> bool:
        return len(self.qual) == len(self.seq) and len(self.seq) > 0

    def tokens(self) -> np.ndarray:
        if self._tok is None:
            t = _TOK[np.frombuffer(self.seq.encode("ascii"), dtype=np.uint8)]
            if (t == 255).any():
                bad = self.seq[int(np.argmax(t == 255))]
                raise ValueError(f"Unknown symbol {bad} in sequence {self.name}")
            self._tok = np.ascontiguousarray(t)
        return self._tok

    def qual_scores(self) -> Optional[np.ndarray]:
        """clamp(c - '!', 0, 93) per base, or None when the read has no qualities."""
        if not self.has_qual():
            return None
        if self._q is None:
            q = np.frombuffer(self.qual.encode("latin-1"), dtype=np.uint8).astype(np.int32) - 33
            self._q = np.ascontiguousarray(np.clip(q, 0, QUAL_RANGE - 1).astype(np.uint8))
        return self._q

    def revcomp(self) -> "FastSeq":
        comp = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "T", "c": "G", "g": "C", "t": "A"}
        rc = "".join(comp.get(c, c) for c in reversed(self.seq))
        return FastSeq(name=f"revcomp({self.name})", seq=rc, qual=self.qual[::-1], comment=self.comment,
                       source_name=self.source_name or self.name, source_rev=not self.source_rev)


def add_revcomps(db: List[FastSeq]) -> List[FastSeq]:
    """addRevcomps (fastseq.cpp:232-238): originals first, then their reverse complements."""
    return list(db) + [s.revcomp() for s in db]


def read_fastseqs(path: str) -> List[FastSeq]:
    """FASTA / FASTQ reader (plain or gzip), multi-line FASTA sequences joined (kseq semantics)."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as fh:
        lines = [ln.rstrip("\n") for ln in fh]
    out: List[FastSeq] = []
    i = 0
    while i < len(lines):
        ln = lines[i]
        if not ln:
            i += 1
            continue
        if ln[0] == ">":
            head = ln[1:].split(None, 1)
            i += 1
            chunks = []
            while i < len(lines) and lines[i][:1] not in (">", "@"):
                chunks.append(lines[i].strip())
                i += 1
            out.append(FastSeq(name=head[0], seq="".join(chunks), comment=head[1] if len(head) > 1 else ""))
        elif ln[0] == "@":
            head = ln[1:].split(None, 1)
            seq = lines[i + 1].strip()
            qual = ""
            if i + 2 < len(lines) and lines[i + 2][:1] == "+":
                qual = lines[i + 3] if i + 3 < len(lines) else ""
                i += 4
            else:
                i += 2
            if len(qual) != len(seq):
                qual = ""              # initFastSeq keeps qualities only when lengths agree
            out.append(FastSeq(name=head[0], seq=seq, qual=qual, comment=head[1] if len(head) > 1 else ""))
        else:
            raise ValueError(f"{path}: unexpected line {i + 1}")
    return out
