"""Diagnostic: per-call wall times of the contexts of a QuaffGPUPool over many steps of the bench workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from quaff_b200 import api

n_ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
x, batches = bench.make_workload(0, 2, 1536)
qp, nullp = bench.load_models()
P = api.QuaffGPUPool(device=0, n_ctx=n_ctx, lib_path=os.environ.get("QG_LIB"))
P.set_refs(x); P.set_params(qp)
cfg = api.dp_config(kmer_threshold=20, band_size=64, kmer_len=6)
flat = [api._flatten(b, True) for b in batches]
null_ll = [np.array([api.null_loglike(nullp, r, P.L) for r in b]) for b in batches]
P.set_read_batches(flat)
for i in range(3):
    P.align_batch(cfg, i % 2, null_ll[i % 2])
log = [[] for _ in range(n_ctx)]

def run(k, g):
    for t in range(steps):
        b = t % 2
        first, count, lo, hi = P.batch_ranges[b][k]
        g.stats(reset=True)
        t0 = time.perf_counter()
        g.align_reads(cfg, null_ll[b][lo:hi], first=first, count=count, split_paths=False)
        t1 = time.perf_counter()
        st = g.stats()
        log[k].append((t0, t1, st["ms_seed"], st["ms_viterbi"], st["ms_traceback"], st["ms_prep"], st["ms_d2h"], st["ms_h2d"]))
T0 = time.perf_counter()
P._each(run)
T1 = time.perf_counter()
print("total %.1f ms for %d steps -> %.1f ms/step, %.0f reads/s" % ((T1 - T0) * 1e3, steps, (T1 - T0) * 1e3 / steps, steps * 1536 / (T1 - T0)))
for k in range(n_ctx):
    d = np.array([(b - a) * 1e3 for a, b, *_ in log[k]])
    print("ctx %d call ms: median %.1f  p90 %.1f  max %.1f" % (k, np.median(d), np.percentile(d, 90), d.max()))
    for t, row in enumerate(log[k]):
        dur = (row[1] - row[0]) * 1e3
        if dur > 1.4 * np.median(d):
            print("   slow call step %d at %.0f ms: %.1f ms  seed %.1f vit %.1f tb %.1f prep %.1f d2h %.1f h2d %.1f" % (t, (row[0] - T0) * 1e3, dur, *row[2:]))
P.close()
