import time, numpy as np, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import parity_cases as pc
from quaff_b200 import api
g = api.QuaffGPU()
x, reads = pc.make_workload(ref_len=50000, n_reads=24, read_len=10000, seed=5)
g.set_refs(x); g.set_reads(reads); g.set_params(pc.default_params())
xi, yi = pc.all_pairs(len(x), len(reads))
cfg = api.dp_config(sparse=False)
for it in range(2):
    t = time.time(); v = g.viterbi(cfg, xi, yi); dt = time.time() - t
    st = g.stats()
    cells = sum(len(x[a]) * len(reads[b]) for a, b in zip(xi, yi))
    print("viterbi wide: %d pairs %.3g cells  wall %.3f s  %.1f GCUPS (wall)" % (len(xi), cells, dt, cells / dt / 1e9), {k: st[k] for k in st if k.startswith('ms_')})
t = time.time(); f = g.forward(cfg, xi, yi); dt = time.time() - t
print("forward wide: wall %.3f s  %.1f GCUPS" % (dt, cells / dt / 1e9))
