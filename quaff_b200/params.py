"""QuaffParams / QuaffNullParams: the reference's JSON formats, host side.

Mirrors src/qmodel.cpp:165-273 (QuaffParams read/write), :1845-1907 (QuaffNullParams) and the
k-mer naming of src/qmodel.cpp:136-163.  As in the reference, `refBase` is written but NOT read
back (readJson never touches it, qmodel.cpp:210-268), so it stays at 0.25 unless set in-process.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

from .seqs import DNA


def gason_double(text: str) -> float:
    """A JSON number as the reference's parser reads it (gason's string2double, src/gason.cpp:73-117): digits are
    accumulated in double arithmetic -- result*10 + d, then d * 0.1^n for the fraction, then a squared-base power for the
    exponent -- which is NOT correctly rounded: the value can differ from strtod's by an ulp or two.  IEEE double
    operations are the same here and there, so this returns the reference's value bit for bit; the parameters the GPU
    path gets from a JSON file are then exactly the ones `quaff` computes with from the same file (score ties in the
    Viterbi traceback are decided in the last bit)."""
    s = text.strip()
    n = len(s)
    i = 0
    neg = i < n and s[i] == "-"
    if neg:
        i += 1
    result = 0.0
    while i < n and s[i].isdigit():
        result = (result * 10) + (ord(s[i]) - 48); i += 1
    if i < n and s[i] == ".":
        i += 1
        fraction = 1.0
        while i < n and s[i].isdigit():
            fraction *= 0.1
            result += (ord(s[i]) - 48) * fraction; i += 1
    if i < n and s[i] in "eE":
        i += 1
        base = 10.0
        if i < n and s[i] == "+":
            i += 1
        elif i < n and s[i] == "-":
            i += 1; base = 0.1
        exponent = 0
        while i < n and s[i].isdigit():
            exponent = exponent * 10 + (ord(s[i]) - 48); i += 1
        power = 1.0
        while exponent:
            if exponent & 1:
                power *= base
            exponent >>= 1; base *= base
        result *= power
    return -result if neg else result


def _loads(text: str):
    """json.loads with every number read the reference's way"""
    return json.loads(text, parse_float=gason_double, parse_int=gason_double)


def kmer_string(kmer: int, k: int) -> str:
    """kmerToString (fastseq.cpp:44-49): first base most significant."""
    s = []
    for _ in range(k):
        s.append(DNA[kmer % 4])
        kmer //= 4
    return "".join(reversed(s))


@dataclass
class SymQualDist:
    p: float = 0.25          # symProb
    q: float = 0.5           # qualTrialSuccessProb
    r: float = 47.0          # qualNumSuccessfulTrials (qualScoreRange / 2)

    def to_json(self) -> Dict[str, float]:
        return {"p": self.p, "q": self.q, "r": self.r}


@dataclass
class QuaffParams:
    match_k: int = 1         # matchContext.kmerLen (1 + -suborder)
    gap_k: int = 0           # indelContext.kmerLen (-gaporder)
    ref_base: List[float] = field(default_factory=lambda: [0.25] * 4)
    begin_insert: np.ndarray = None
    begin_delete: np.ndarray = None
    extend_insert: float = 0.5
    extend_delete: float = 0.5
    insert: List[SymQualDist] = None
    match: List[List[SymQualDist]] = None      # [ref base i][read context k-mer j]

    def __post_init__(self):
        nG, nK = 4 ** self.gap_k, 4 ** self.match_k
        if self.begin_insert is None:
            self.begin_insert = np.full(nG, 0.5)
        if self.begin_delete is None:
            self.begin_delete = np.full(nG, 0.5)
        if self.insert is None:
            self.insert = [SymQualDist() for _ in range(4)]
        if self.match is None:
            self.match = [[SymQualDist() for _ in range(nK)] for _ in range(4)]

    @property
    def n_match_kmers(self) -> int:
        return 4 ** self.match_k

    @property
    def n_gap_kmers(self) -> int:
        return 4 ** self.gap_k

    # ---- JSON -------------------------------------------------------------------------------
    @staticmethod
    def from_json(text: str) -> "QuaffParams":
        jm = _loads(text)
        mk = int(jm.get("matchOrder", 1))
        gk = int(jm.get("gapOrder", 0))
        qp = QuaffParams(match_k=mk, gap_k=gk)
        for g in range(qp.n_gap_kmers):
            key = kmer_string(g, gk)
            qp.begin_insert[g] = float(jm["beginInsert"][key])
            qp.begin_delete[g] = float(jm["beginDelete"][key])
        qp.extend_insert = float(jm["extendInsert"])
        qp.extend_delete = float(jm["extendDelete"])
        for i in range(4):
            d = jm["insert"][DNA[i]]
            qp.insert[i] = SymQualDist(float(d["p"]), float(d["q"]), float(d["r"]))
        for jp in range(0, qp.n_match_kmers, 4):
            prefix = kmer_string(jp, mk)[: mk - 1]
            for i in range(4):
                for js in range(4):
                    d = jm["match"][prefix][DNA[i]][DNA[js]]
                    qp.match[i][jp + js] = SymQualDist(float(d["p"]), float(d["q"]), float(d["r"]))
        return qp

    @staticmethod
    def load(path: str) -> "QuaffParams":
        with open(path) as fh:
            return QuaffParams.from_json(fh.read())

    def to_json(self) -> str:
        jm: Dict[str, object] = {}
        if self.match_k != 1:
            jm["matchOrder"] = self.match_k
        if self.gap_k != 0:
            jm["gapOrder"] = self.gap_k
        jm["refBase"] = {DNA[i]: self.ref_base[i] for i in range(4)}
        jm["beginInsert"] = {kmer_string(g, self.gap_k): float(self.begin_insert[g]) for g in range(self.n_gap_kmers)}
        jm["beginDelete"] = {kmer_string(g, self.gap_k): float(self.begin_delete[g]) for g in range(self.n_gap_kmers)}
        jm["extendInsert"] = self.extend_insert
        jm["extendDelete"] = self.extend_delete
        jm["insert"] = {DNA[i]: self.insert[i].to_json() for i in range(4)}
        m: Dict[str, object] = {}
        for jp in range(0, self.n_match_kmers, 4):
            prefix = kmer_string(jp, self.match_k)[: self.match_k - 1]
            m[prefix] = {DNA[i]: {DNA[js]: self.match[i][jp + js].to_json() for js in range(4)} for i in range(4)}
        jm["match"] = m
        return json.dumps(jm, indent=1)

    # ---- flat views for the C ABI -----------------------------------------------------------
    def match_pqr(self) -> np.ndarray:
        """[4][4^K][3] doubles (p, q, r)."""
        a = np.empty((4, self.n_match_kmers, 3))
        for i in range(4):
            for j in range(self.n_match_kmers):
                d = self.match[i][j]
                a[i, j] = (d.p, d.q, d.r)
        return a

    def insert_pqr(self) -> np.ndarray:
        return np.array([(d.p, d.q, d.r) for d in self.insert])


@dataclass
class QuaffNullParams:
    null_emit: float = 0.5
    null: List[SymQualDist] = field(default_factory=lambda: [SymQualDist() for _ in range(4)])

    @staticmethod
    def from_json(text: str) -> "QuaffNullParams":
        jm = _loads(text)
        np_ = QuaffNullParams(null_emit=float(jm["nullEmit"]))
        for i in range(4):
            d = jm["null"][DNA[i]]
            np_.null[i] = SymQualDist(float(d["p"]), float(d["q"]), float(d["r"]))
        return np_

    @staticmethod
    def load(path: str) -> "QuaffNullParams":
        with open(path) as fh:
            return QuaffNullParams.from_json(fh.read())

    def to_json(self) -> str:
        return json.dumps({"nullEmit": self.null_emit, "null": {DNA[i]: self.null[i].to_json() for i in range(4)}})

    def pqr(self) -> np.ndarray:
        return np.array([(d.p, d.q, d.r) for d in self.null])


def random_params(rng: np.random.Generator, match_k: int = 1, gap_k: int = 0) -> QuaffParams:
    """Plausible random parameters of a given order, for tests and synthetic benchmarks."""
    qp = QuaffParams(match_k=match_k, gap_k=gap_k)
    qp.begin_insert = rng.uniform(0.01, 0.08, qp.n_gap_kmers)
    qp.begin_delete = rng.uniform(0.01, 0.08, qp.n_gap_kmers)
    qp.extend_insert = float(rng.uniform(0.3, 0.7))
    qp.extend_delete = float(rng.uniform(0.3, 0.7))
    pins = rng.dirichlet(np.full(4, 20.0))
    for i in range(4):
        qp.insert[i] = SymQualDist(float(pins[i]), float(rng.uniform(0.6, 0.92)), float(rng.uniform(15, 95)))
    for j in range(qp.n_match_kmers):
        suffix = j % 4
        for i in range(4):
            # P(read base = suffix | ref base i) must sum to 1 over suffix for each (i, prefix)
            pass
    for jp in range(0, qp.n_match_kmers, 4):
        for i in range(4):
            alpha = np.full(4, 1.0)
            alpha[i] = 30.0
            pr = rng.dirichlet(alpha)
            for js in range(4):
                qp.match[i][jp + js] = SymQualDist(float(pr[js]), float(rng.uniform(0.6, 0.92)), float(rng.uniform(15, 95)))
    return qp
