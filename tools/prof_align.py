"""One small pass of the align hot path on cuda:0 for ncu captures (profiles/): N reads of the bench workload through one
context.  usage: python tools/prof_align.py [n_reads] [passes]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                              # noqa: E402
from quaff_b200 import api                                # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x, batches = bench.make_workload(0, 1, n)
qp, nullp = bench.load_models()
G = api.QuaffGPU(device=0)
G.set_refs(x); G.set_params(qp); G.set_reads(batches[0])
cfg = api.dp_config(kmer_threshold=20, band_size=64, kmer_len=6)
null_ll = np.array([api.null_loglike(nullp, r, G.L) for r in batches[0]])
for _ in range(passes):
    G.stats(reset=True)
    r = G.align_reads(cfg, null_ll, split_paths=False)
    st = G.stats()
    print({k: st[k] for k in ("ms_seed", "ms_viterbi", "ms_traceback", "ms_prep", "ms_envelope", "ms_d2h", "kernel_launches", "kmer_hits", "cell_updates", "trace_bytes")})
G.close()
