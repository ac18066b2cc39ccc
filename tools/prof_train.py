"""One E-step of the bench's train workload on cuda:0 for ncu captures (profiles/): N reads through one context.
usage: python tools/prof_train.py [n_reads] [passes]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                              # noqa: E402
from quaff_b200 import api                                # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x, batches = bench.make_workload(0, 1, n)
qp, nullp = bench.load_models()
G = api.QuaffGPU(device=0)
G.set_refs(x); G.set_params(qp); G.set_reads(batches[0])
cfg = api.dp_config(kmer_threshold=20, band_size=64, kmer_len=6)
null_ll = np.array([api.null_loglike(nullp, r, G.L) for r in batches[0]])
for _ in range(passes):
    G.stats(reset=True)
    r = G.estep(cfg, True, null_ll)
    st = G.stats()
    print({k: st[k] for k in ("ms_seed", "ms_forward", "ms_backward", "ms_prep", "kernel_launches", "cell_updates", "fwd_store_bytes")}, r["loglike"])
G.close()
