"""not-gpu: the N > 1 host path with world_size 2 over gloo -- read sharding + the one all-reduce of the E-step.
Each rank runs its shard through qg_estep on the CPU-thread emulation of the kernels (test infrastructure); the
summed counts must equal the single-process E-step of the oracle over all reads."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, emu_lib, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import parity_cases as pc
    from quaff_b200 import api
    from quaff_b200 import dist as qd
    from quaff_b200.params import QuaffNullParams
    x, reads = pc.make_workload(ref_len=2500, n_reads=3, read_len=220, seed=13)
    qp = pc.default_params()
    nullp = QuaffNullParams.load(os.path.join(ROOT, "tests", "golden", "testquaffnullparams.json"))
    G = api.QuaffGPU(lib_path=emu_lib)
    G.set_refs(x); G.set_params(qp)
    G.set_fb_exact(True)          # log-space kernels: per-read log-likelihoods bit-identical to the oracle
    null_ll = np.array([api.null_loglike(nullp, r, G.L) for r in reads])
    r = qd.distributed_estep(G, api.dp_config(kmer_threshold=6), reads, True, null_ll)
    q.put((rank, r["counts"], r["loglike"], r["shard"], list(r["y_loglike"])))
    G.close()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from quaff_b200.dist import shard_bounds
    for n in (0, 1, 5, 16, 17):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_estep_allreduce_world2(emu_lib, oracle):
    import parity_cases as pc
    from oracle import pyoracle as po
    from quaff_b200.params import QuaffNullParams
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emu_lib, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # both ranks hold the same global sums
    np.testing.assert_array_equal(res[0][1], res[1][1]); assert res[0][2] == res[1][2]
    assert res[0][3] == (0, 2) and res[1][3] == (2, 3)
    x, reads = pc.make_workload(ref_len=2500, n_reads=3, read_len=220, seed=13)
    qp = pc.default_params()
    nullp = QuaffNullParams.load(os.path.join(ROOT, "tests", "golden", "testquaffnullparams.json"))
    xs, ys = pc.seqbufs(x, reads)
    o = oracle.estep(xs, ys, oracle.scores(qp), nullp, True, po.make_config(kmer_threshold=6))
    np.testing.assert_allclose(res[0][1], o["counts"], rtol=1e-9, atol=1e-12)
    assert abs(res[0][2] - o["loglike"].sum()) <= 1e-9 * abs(res[0][2])
    np.testing.assert_allclose(res[0][4] + res[1][4], o["loglike"], rtol=1e-12)
