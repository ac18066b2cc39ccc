"""Multi-GPU plumbing: one process per GPU, reads sharded contiguously, the reference and the model replicated.

align / overlap need no collective (results are per read / per pair and are gathered on the host).  train needs one
small exchange per EM iteration: the per-rank partial QuaffParamCounts and log-likelihood are summed with a single
all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests) -- <= 24 509 doubles at -order 2, latency-bound.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split; the first (n % world) ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(items: Sequence, rank: int, world: int):
    lo, hi = shard_bounds(len(items), rank, world)
    return items[lo:hi]


def allreduce_counts(counts: np.ndarray, loglike: float, device: Optional[str] = None) -> Tuple[np.ndarray, float]:
    """Sum of (QuaffParamCounts, log-likelihood) over ranks.  No-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts, loglike
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.empty(len(counts) + 1, dtype=torch.float64, device=device)
    buf[:-1] = torch.from_numpy(np.ascontiguousarray(counts)).to(device)
    buf[-1] = loglike
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    out = buf.cpu().numpy()
    return out[:-1].copy(), float(out[-1])


def distributed_estep(G, cfg, reads, use_null: bool, null_ll: np.ndarray, sort_order: Optional[List[List[int]]] = None):
    """One E-step of `train`: this rank's shard of the reads through qg_estep, then the all-reduce.
    `reads`, `null_ll`, `sort_order` are the GLOBAL lists; returns global counts / log-likelihood and this rank's
    slice of the per-read outputs."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(len(reads), rank, world)
    G.set_reads(reads[lo:hi])
    r = G.estep(cfg, use_null, np.asarray(null_ll)[lo:hi], None if sort_order is None else sort_order[lo:hi])
    counts, ll = allreduce_counts(r["counts"], r["loglike"])
    return dict(counts=counts, loglike=ll, y_loglike=r["y_loglike"], sort_order=r["sort_order"], shard=(lo, hi))
