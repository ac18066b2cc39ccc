"""Quick device probe: cfg4-shaped workload (8 kb reads vs a 5 Mb reference, both strands) through
seeding + Viterbi (+ Forward), printing the library's per-stage CUDA-event timings."""
import argparse
import json
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from quaff_b200 import api
from quaff_b200.params import QuaffParams
from quaff_b200.seqs import add_revcomps
from quaff_b200.synth import random_ref, sample_reads

ap = argparse.ArgumentParser()
ap.add_argument("--ref-len", type=int, default=5_000_000)
ap.add_argument("--reads", type=int, default=64)
ap.add_argument("--read-len", type=int, default=8000)
ap.add_argument("--forward", action="store_true")
ap.add_argument("--backward", action="store_true")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--lib", default=None)
a = ap.parse_args()

t0 = time.time()
ref = random_ref(a.ref_len, 1)
reads, _, _ = sample_reads(ref, a.reads, a.read_len, 2)
x = add_revcomps([ref])
print(f"synth {time.time() - t0:.1f}s", flush=True)
qp = QuaffParams.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "defaultparams.json"))
G = api.QuaffGPU(lib_path=a.lib) if a.lib else api.QuaffGPU()
G.set_refs(x); G.set_reads(reads); G.set_params(qp)
cfg = api.dp_config(kmer_threshold=20)
xi = np.tile(np.arange(2, dtype=np.uint32), len(reads)); yi = np.repeat(np.arange(len(reads), dtype=np.uint32), 2)
for rep in range(a.reps):
    G.stats(reset=True)
    t = time.time()
    v = G.viterbi(cfg, xi, yi)
    wall = time.time() - t
    st = G.stats()
    cu = st["cell_updates"]
    print(json.dumps(dict(rep=rep, what="viterbi", wall_s=round(wall, 3), pairs=len(xi), cu=cu, hits=st["kmer_hits"],
                          gcups_fill=round(cu / 1e9 / (st["ms_viterbi"] / 1e3), 2) if st["ms_viterbi"] else None,
                          ghits_s=round(st["kmer_hits"] / 1e9 / (st["ms_seed"] / 1e3), 2) if st["ms_seed"] else None,
                          **{k: round(val, 2) if isinstance(val, float) else val for k, val in st.items()})), flush=True)
    if a.forward:
        G.stats(reset=True)
        t = time.time(); f = G.forward(cfg, xi, yi); wall = time.time() - t
        st = G.stats()
        print(json.dumps(dict(rep=rep, what="forward", wall_s=round(wall, 3), gcups=round(st["cell_updates"] / 1e9 / (st["ms_forward"] / 1e3), 2),
                              ms_forward=round(st["ms_forward"], 2), ms_seed=round(st["ms_seed"], 2))), flush=True)
    if a.backward:
        G.stats(reset=True)
        t = time.time(); b = G.backward_counts(cfg, xi, yi); wall = time.time() - t
        st = G.stats()
        print(json.dumps(dict(rep=rep, what="backward", wall_s=round(wall, 3), gcups=round(st["cell_updates"] / 1e9 / (st["ms_backward"] / 1e3), 2),
                              ms_backward=round(st["ms_backward"], 2), ms_forward=round(st["ms_forward"], 2))), flush=True)
print("scores", v["score"][:6], "mean path len", np.mean([len(p) for p in v["paths"]]))
