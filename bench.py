#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: `quaff align -kmatchband 64` on synthetic nanopore-like
8 kb reads against a 5 Mb reference, both strands (config 4), sharded by read over N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one batch of READS_PER_STEP reads per GPU through the whole hot path: k-mer seeding -> diagonal
envelope -> banded Viterbi fill -> best strand -> traceback -> paths on the host.
  value : reads/s of the whole job with the step's reads already resident in HBM (timed between
          torch.cuda.synchronize + barrier, max over ranks)
  e2e   : the same through the reference-facing call with HOST buffers: upload of the step's reads and download of
          the scores/paths inside the timed region
`--impl reference` times the reference's own CPU implementation (oracle/_ref/quaff, the unmodified sources built -O3,
`-threads <all host cores>`) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REF_LEN = 5_000_000
READ_LEN = 8000
READS_PER_STEP = 6144            # per GPU: 1536 per context x 4 contexts (Viterbi and traceback launches of 1536 reads fill the SMs, 768 did not); ~22 GB of traceback pointers per step
CFG4_WORKLOAD = ("cfg4: quaff align, synthetic 8 kb nanopore-like reads (12% error) vs 5 Mb random reference, both strands, "
                 "-kmatch 6 -kmatchn 20 -kmatchband 64, default params, fixed null model")
POOL_BATCHES = 2                 # distinct read batches cycled through the steps (bounds host synthesis time)
# peak lane-instructions/s and the instructions one cell update needs in the minimal formulation
SM_COUNT, LANES, SM_MAX_MHZ = 148, 128, 1965.0
PEAK_LANE_INSTR = SM_COUNT * LANES * SM_MAX_MHZ * 1e6
INSTR_PER_CU = dict(viterbi=13, forward=9, backward=25, overlap=26)
PEAK_SMEM_ATOMIC = SM_COUNT * 32 * SM_MAX_MHZ * 1e6      # histogram increments/s upper bound
MEASURED_SMEM_ATOMIC = SM_COUNT * 16 * SM_MAX_MHZ * 1e6  # what random full-warp shared atomics reach on this part (tools/ubench/atoms.cu)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        if os.environ.get("QG_BENCH_NO_SAMPLER"):
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(rank: int, n_batches: int, reads_per_step: int, ref_len: int = REF_LEN, read_len: int = READ_LEN):
    from quaff_b200.seqs import add_revcomps
    from quaff_b200.synth import random_ref, sample_reads
    ref = random_ref(ref_len, 1)                                   # reference seed 1, read seed 2 (+ rank)
    batches = []
    for b in range(n_batches):
        reads, _, _ = sample_reads(ref, reads_per_step, read_len, 2 + 1000 * rank + b, name_prefix=f"r{rank}b{b}_")
        batches.append(reads)
    return add_revcomps([ref]), batches


PARAMS_JSON = os.path.join(ROOT, "tests", "golden", "defaultparams.json")
NULL_JSON = os.path.join(ROOT, "tests", "golden", "testquaffnullparams.json")


def load_models():
    """The model files both arms read.  quaff_b200.params parses JSON numbers the way the reference's parser does (gason), so the
    GPU arm computes with exactly the values `quaff -params ... -null ...` holds after reading the same two files."""
    from quaff_b200.params import QuaffNullParams, QuaffParams
    return QuaffParams.load(PARAMS_JSON), QuaffNullParams.load(NULL_JSON)


# ------------------------------------------------------------------------------------------------------------------
def reference_cpu_run(n_reads: int, threads: int, ref_len: int = REF_LEN, read_len: int = READ_LEN, seed_rank: int = 0, name_prefix: str = "r0b0_"):
    """`quaff align ref.fa reads.fq -kmatchband 64 -threads T` with the reference binary built from the unmodified
    sources (oracle/_ref/quaff).  Returns (reads/s, seconds, kind, SAM text of the run or None)."""
    from oracle import pyoracle as po
    from quaff_b200.synth import random_ref, sample_reads
    qp, nullp = load_models()
    ref = random_ref(ref_len, 1)
    reads, _, _ = sample_reads(ref, n_reads, read_len, 2 + 1000 * seed_rank, name_prefix=name_prefix)   # = the first reads of rank 0's batch 0
    if os.path.exists(po.REF_QUAFF):
        with tempfile.TemporaryDirectory() as td:
            fa, fq = os.path.join(td, "ref.fa"), os.path.join(td, "reads.fq")
            pj, nj = os.path.join(td, "params.json"), os.path.join(td, "null.json")
            with open(fa, "w") as fh:
                fh.write(f">{ref.name}\n{ref.seq}\n")
            with open(fq, "w") as fh:
                for r in reads:
                    fh.write(f"@{r.name}\n{r.seq}\n+\n{r.qual}\n")
            open(pj, "w").write(open(PARAMS_JSON).read()); open(nj, "w").write(open(NULL_JSON).read())      # the same text the GPU arm parsed
            cmd = [po.REF_QUAFF, "align", fa, fq, "-params", pj, "-null", nj, "-kmatchband", "64", "-format", "sam", "-threads", str(threads)]
            t0 = time.time()
            res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=td)   # the reference CLI leaves a tempdir in its working directory
            dt = time.time() - t0
            if res.returncode != 0:
                raise RuntimeError("reference quaff failed: " + res.stderr[-500:])
        return n_reads / dt, dt, "reference", res.stdout
    # no reference build here: time the C restatement (1 thread)
    O = po.Oracle()
    s = O.scores(qp)
    from quaff_b200.seqs import add_revcomps
    xs = [po.SeqBuf(v.tokens(), None) for v in add_revcomps([ref])]
    cfg = po.make_config(kmer_threshold=20)
    t0 = time.time()
    for r in reads:
        y = po.SeqBuf(r.tokens(), r.qual_scores())
        for x in xs:
            O.viterbi(x, y, s, cfg)
    dt = time.time() - t0
    return n_reads / dt, dt, "port", None


REFERENCE_ARM_BUDGET_S = 200.0     # the whole `--impl reference` run: one step costs ~13 s per read per core at this shape


def run_reference_arm(args):
    """The reference's own CPU implementation on this box's host cores.  A step = `quaff align -threads <cores>` over one
    read per core (fewer reads would leave threads idle and understate it).  One step takes 13-16 s, so the arm runs as many
    of the requested warm-up/timed steps as fit REFERENCE_ARM_BUDGET_S and reports the numbers it really ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = max(2, min(cores, 32))
    t_begin = time.time()
    v0, dt0, kind, _ = reference_cpu_run(sample, cores, args.ref_len, args.read_len)         # first step: also the cost estimate
    fit = max(1, int((REFERENCE_ARM_BUDGET_S - (time.time() - t_begin)) // max(dt0, 1e-3)))
    warm = 1 if (args.warmup > 0 and fit >= 2) else 0                                        # the first run is the warm-up when a second one fits
    steps = max(1, min(args.steps, fit if warm else fit + 1))
    vals = [] if warm else [(v0, dt0)]
    while len(vals) < steps:
        v, dt, kind, _ = reference_cpu_run(sample, cores, args.ref_len, args.read_len)
        vals.append((v, dt))
    dt = float(np.mean([b for _, b in vals])); v = sample / dt
    line = {
        "impl": "reference", "metric": "align_reads_per_sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": len(vals),
        "warmup": warm, "requested_steps": args.steps, "requested_warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": CFG4_WORKLOAD, "reads_per_step_per_gpu": sample, "ref_len": args.ref_len, "read_len": args.read_len,
                   "note": "reference CPU implementation (oracle/_ref/quaff align -threads <cores>, unmodified sources, -O3); a step is one read per "
                           "host core of the same workload; steps/warmup are what was run inside the arm's %d s budget" % int(REFERENCE_ARM_BUDGET_S)},
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} reads x {args.read_len} b vs {args.ref_len} b, both strands, -threads {cores}, {len(vals)} timed run(s)"},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


GPU_QUAFF = os.path.join(ROOT, "host", "_build", "quaff-gpu")


def run_cli(args):
    """the drop-in itself, end to end.  `host/_build/quaff-gpu align ref.fa reads.fq ... -format sam -gpu <devices>` --
    the reference's own CLI with the three seams routed to libquaffgpu -- timed by wall clock from exec to exit: FASTA/FASTQ
    parsing, tokenising, the pooled GPU contexts, row assembly and SAM formatting in the worker threads, the write.  Reported next
    to the CUDA initialisation + parse time of the same command on two reads (a fixed cost that a longer read set amortises).
    Single process: with --gpus N the CLI itself drives N devices (`-gpu 0,..,N-1`, qg_pool_*)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from quaff_b200.synth import random_ref, sample_reads
    if not os.path.exists(GPU_QUAFF):
        print(json.dumps({"metric": "align_reads_per_sec_cli", "unavailable": "host/_build/quaff-gpu not built (needs the reference sources at build time)"}), flush=True)
        return
    n_distinct = args.reads_per_step if args.reads_per_step != READS_PER_STEP else 12288
    copies = int(os.environ.get("QB_CLI_COPIES", "4"))          # the read file = `copies` renamed copies of the distinct reads (synthesis is the slow part of this bench)
    n_reads = n_distinct * copies
    ref = random_ref(args.ref_len, 1)
    reads, _, _ = sample_reads(ref, n_distinct, args.read_len, 2, name_prefix="r0b0_")
    devs = ",".join(str(d) for d in range(args.gpus))
    with tempfile.TemporaryDirectory() as td:
        fa, fq, fq2 = os.path.join(td, "ref.fa"), os.path.join(td, "reads.fq"), os.path.join(td, "two.fq")
        with open(fa, "w") as fh:
            fh.write(f">{ref.name}\n{ref.seq}\n")
        with open(fq, "w") as fh:
            for cp in range(copies):
                for r in reads:
                    fh.write(f"@{r.name}{'' if cp == 0 else '_copy%d' % cp}\n{r.seq}\n+\n{r.qual}\n")
        with open(fq2, "w") as fh:
            for r in reads[:2]:
                fh.write(f"@{r.name}\n{r.seq}\n+\n{r.qual}\n")
        base = ["-params", PARAMS_JSON, "-null", NULL_JSON, "-kmatchband", "64", "-format", "sam", "-gpu", devs]
        out_sam = os.path.join(td, "out.sam")

        phases = {}

        def run(fastq):
            t0 = time.time()
            with open(out_sam, "w") as fh:
                res = subprocess.run([GPU_QUAFF, "align", fa, fastq] + base, stdout=fh, stderr=subprocess.PIPE, text=True, cwd=td,
                                     env=dict(os.environ, QUAFF_GPU_TRACE="1"))
            dt = time.time() - t0
            if res.returncode != 0:
                raise RuntimeError("quaff-gpu failed: " + res.stderr[-800:])
            phases.clear()
            for ln in res.stderr.splitlines():                   # "[quaff-gpu] <phase>   <seconds> s": the seam's own host-phase stamps
                if ln.startswith("[quaff-gpu]"):
                    name, _, val = ln[len("[quaff-gpu]"):].strip().rpartition("  ")
                    phases[name.strip()] = float(val.replace("s", "").strip())
            phases["whole process"] = dt
            return dt
        fixed = min(run(fq2) for _ in range(2))
        for _ in range(max(1, args.warmup // 3)):
            run(fq)
        sampler = ClockSampler(0); sampler.start()
        times = [run(fq) for _ in range(max(1, min(args.steps, 5)))]
        clocks = sampler.stop()
        n_sam = sum(1 for ln in open(out_sam) if ln and not ln.startswith("@"))
        sam_bytes = os.path.getsize(out_sam); fq_bytes = os.path.getsize(fq)
    dt = float(np.median(times))
    line = {
        "metric": "align_reads_per_sec_cli", "value": n_reads / dt, "unit": "reads/s", "n_gpus": args.gpus, "steps": len(times), "warmup": max(1, args.warmup // 3),
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": CFG4_WORKLOAD + "; through the reference CLI with -gpu (host/_build/quaff-gpu), FASTQ in, SAM out",
                   "reads": n_reads, "distinct_reads": n_distinct, "ref_len": args.ref_len, "read_len": args.read_len, "devices": devs,
                   "fastq_bytes": fq_bytes, "sam_bytes": sam_bytes, "sam_records": n_sam},
        "e2e": {"value": n_reads / dt, "unit": "reads/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                "note": "wall clock of the whole process, exec to exit"},
        "fixed_cost_s": fixed, "reads_per_sec_excluding_fixed_cost": n_reads / max(dt - fixed, 1e-9),
        "host_phases_s": dict(phases, note="last timed run; what is not listed (FASTA/FASTQ parsing by the reference's own loader, process start, "
                                           "CUDA initialisation before the seam) is the remainder of 'whole process'"),
        "gpu_launches": None, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def _dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU path"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    return torch, dist, rank, world, local, barrier


def run_cfg5(args):
    """BASELINE config 5: `quaff align -kmatchoff` (full, unbanded DP) on synthetic 10 kb reads against a 50 kb reference,
    both strands -- the tiled wavefront of qg_tile.cuh.  A step = `--reads-per-step` reads (default 16) of the 2 000."""
    from quaff_b200 import api
    from quaff_b200.seqs import add_revcomps
    from quaff_b200.synth import random_ref, sample_reads
    torch, dist, rank, world, local, barrier = _dist_setup()
    B = args.reads_per_step if args.reads_per_step != READS_PER_STEP else 16
    ref = random_ref(50_000, 3)                                                     # reference seed 3, reads seed 4
    x = add_revcomps([ref])
    batches = [sample_reads(ref, B, 10_000, 4 + 1000 * rank + b, name_prefix=f"r{rank}b{b}_")[0] for b in range(POOL_BATCHES)]
    qp, nullp = load_models()
    G = api.QuaffGPU(device=local)
    G.set_refs(x); G.set_params(qp)
    cfg = api.dp_config(sparse=False)
    null_ll = [np.array([api.null_loglike(nullp, r, G.L) for r in b]) for b in batches]
    flat = [api._flatten(b, True) for b in batches]

    def step(i):
        tok, qual, off = flat[i % POOL_BATCHES]
        G.set_seqs_raw(api.QG_READS, tok, qual, off)
        return G.align_reads(cfg, null_ll[i % POOL_BATCHES], split_paths=False)
    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(local); sampler.start()
    G.stats(reset=True)
    barrier(); t0 = time.perf_counter()
    for i in range(args.steps):
        r = step(args.warmup + i)
    barrier(); t1 = time.perf_counter()
    clocks = sampler.stop()
    st = G.stats()
    times = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(st["cell_updates"]), st["ms_viterbi"], float(st["kernel_launches"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX); dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dt = float(times[0])
    if rank == 0:
        cu = float(sums[0]); ms_fill = float(sums[1]) / world
        gcups_fill = cu / world / (ms_fill / 1e3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pyoracle as po
            if os.path.exists(po.REF_QUAFF):
                with tempfile.TemporaryDirectory() as td:
                    fa, fq, pj, nj = (os.path.join(td, n) for n in ("ref.fa", "reads.fq", "params.json", "null.json"))
                    open(fa, "w").write(f">{ref.name}\n{ref.seq}\n")
                    r0 = batches[0][0]
                    open(fq, "w").write(f"@{r0.name}\n{r0.seq}\n+\n{r0.qual}\n")
                    open(pj, "w").write(open(PARAMS_JSON).read()); open(nj, "w").write(open(NULL_JSON).read())
                    tc = time.time()
                    res = subprocess.run([po.REF_QUAFF, "align", fa, fq, "-params", pj, "-null", nj, "-kmatchoff", "-format", "sam", "-threads", "2"],
                                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=td)
                    secs = time.time() - tc
                if res.returncode == 0:
                    from quaff_b200 import sam
                    G.set_reads([r0])
                    g1 = G.align_reads(cfg, null_ll[0][:1], split_paths=False)
                    checked = sam.compare_batch(res.stdout, [r0], ref.name, len(ref), g1)
                    cpu = {"value": 1.0 / secs, "unit": "reads/s", "cores": 2, "kind": "reference", "parity_checked": checked,
                           "sample": f"1 read x 10 kb vs 50 kb, both strands, quaff align -kmatchoff -threads 2 ({secs:.1f} s; 12 GB per strand)"}
        line = {
            "metric": "align_reads_per_sec", "value": B * args.steps * world / dt, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg5: quaff align -kmatchoff (full DP), synthetic 10 kb reads vs 50 kb reference, both strands, default params",
                       "reads_per_step_per_gpu": B, "ref_len": 50_000, "read_len": 10_000,
                       "l2": "every step writes and re-reads its own pointer tiles (4 bit per cell, ~%.1f GB per step), far larger than L2" % (cu / world / args.steps / 2 / 1e9)},
            "e2e": {"value": B * args.steps * world / dt, "unit": "reads/s", "h2d_bytes_per_step": int(sum(a.nbytes for a in flat[0] if a is not None)),
                    "d2h_bytes_per_step": int(r["paths"].nbytes + 8 * 4 * B), "note": "this workload is timed through the host-buffer call only"},
            "gpu_launches": int(float(sums[2]) / world), "clocks": clocks,
            "gcups": {"viterbi_fill_wall": cu / dt / 1e9, "viterbi_fill_kernels": gcups_fill, "cell_updates_per_read": cu / (B * args.steps * world)},
            "roofline": {"kernel": "qg_tile_kernel<0>", "bound": "lane_instr", "achieved": gcups_fill, "peak": PEAK_LANE_INSTR / INSTR_PER_CU["viterbi"] / 1e9,
                         "unit": "GCUPS", "frac": gcups_fill * 1e9 / (PEAK_LANE_INSTR / INSTR_PER_CU["viterbi"]), "traffic": None},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    G.close()
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """One EM E-step per step (seam C, QuaffTrainer::getCounts, qmodel.cpp:2005-2032) over this rank's shard of cfg4-shaped
    reads: Forward of every read x strand, Backward + counts of the pairs that pass the gate, then ONE all-reduce of the
    QuaffParamCounts (+ log-likelihood) over the ranks -- NCCL over NVLink.  Before timing, the collective is checked: every rank
    runs a few reads, the reduced counts must equal what rank 0 gets from all of those reads alone (1e-9 relative)."""
    from quaff_b200 import api
    from quaff_b200.dist import allreduce_counts
    torch, dist, rank, world, local, barrier = _dist_setup()
    B = args.reads_per_step if args.reads_per_step != READS_PER_STEP else 1024      # 256 reads left three quarters of the warp slots empty (3.2 k vs 6.4 k reads/s)
    x, batches = make_workload(rank, POOL_BATCHES, B, args.ref_len, args.read_len)
    qp, nullp = load_models()
    G = api.QuaffGPU(device=local)
    G.set_refs(x); G.set_params(qp)
    # the timed E-steps go through the native pool (qg_pool_estep): `--contexts` contexts on this rank's GPU, each with a
    # contiguous range of the step's reads, so that one context's host work overlaps the other's kernels
    n_ctx = args.contexts if args.contexts != 4 else 2
    P = api.QuaffPool(devices=[local], contexts_per_device=n_ctx)
    P.set_refs(x); P.set_params(qp)
    cfg = api.dp_config(kmer_threshold=20, band_size=64, kmer_len=6)
    null_ll = [np.array([api.null_loglike(nullp, r, G.L) for r in b]) for b in batches]

    # ---- the collective, checked on a small global read list every rank can build (seed 9) --------------------------------
    allreduce_checked = None
    if world > 1:
        from quaff_b200.synth import sample_reads
        per = 3
        greads, _, _ = sample_reads(x[0], per * world, args.read_len, 9, name_prefix="g")
        gnull = np.array([api.null_loglike(nullp, r, G.L) for r in greads])
        lo, hi = rank * per, (rank + 1) * per
        G.set_reads(greads[lo:hi])
        mine = G.estep(cfg, True, gnull[lo:hi])
        red_counts, red_ll = allreduce_counts(mine["counts"], mine["loglike"])
        if rank == 0:
            G.set_reads(greads)
            alone = G.estep(cfg, True, gnull)
            scale = max(1.0, float(np.abs(alone["counts"]).max()))
            err = float(np.max(np.abs(red_counts - alone["counts"]) / (np.abs(alone["counts"]) + 1e-9 * scale)))
            assert err <= 1e-9 and abs(red_ll - alone["loglike"]) <= 1e-9 * abs(alone["loglike"]), (err, red_ll, alone["loglike"])
            allreduce_checked = {"reads": per * world, "max_rel_err": err, "loglike_abs_err": abs(red_ll - alone["loglike"])}

    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ar_ms = []

    flat = [api._flatten(b, True) for b in batches]               # host buffers of every batch (tokens, qualities, offsets), built once

    def step(i, timed=False):
        b = i % POOL_BATCHES
        r = P.estep_raw(cfg, True, *flat[b], null_ll=null_ll[b])    # the step's reads go up from host memory every step
        ev0.record()
        counts, ll = allreduce_counts(r["counts"], r["loglike"])
        ev1.record(); ev1.synchronize()
        if timed:
            ar_ms.append(ev0.elapsed_time(ev1))
        return counts, ll
    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(local); sampler.start()
    P.stats(reset=True)
    barrier(); t0 = time.perf_counter()
    for i in range(args.steps):
        counts, ll = step(args.warmup + i, True)
    barrier(); t1 = time.perf_counter()
    clocks = sampler.stop()
    sts = P.stats()
    st = {k: (sum(s_[k] for s_ in sts) if not k.startswith("ms_") else sum(s_[k] for s_ in sts) / len(sts)) for k in sts[0]}   # stage times: mean over the contexts (they overlap on the GPU)
    times = torch.tensor([t1 - t0, float(np.mean(ar_ms))], dtype=torch.float64, device="cuda")
    # Forward visits every envelope cell once; the Backward pass visits the gated pairs' cells again (cell_updates counts both)
    sums = torch.tensor([float(st["cell_updates"]), st["ms_forward"], st["ms_backward"], st["ms_seed"], float(st["kernel_launches"]), float(st["fwd_store_bytes"])],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX); dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dt = float(times[0])
    if rank == 0:
        cu = float(sums[0]); ms_f = float(sums[1]) / world; ms_b = float(sums[2]) / world; ms_s = float(sums[3]) / world
        h2d = int(sum(a.nbytes for a in flat[0] if a is not None))
        line = {
            "metric": "train_estep_reads_per_sec", "value": B * args.steps * world / dt, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "train E-step over cfg4's reads: 8 kb reads vs 5 Mb reference, both strands, -kmatch 6 -kmatchn 20 -kmatchband 64, default params, "
                                   "fixed null model; Forward of every pair, Backward + counts of the gated pairs, one counts all-reduce per step",
                       "reads_per_step_per_gpu": B, "ref_len": args.ref_len, "read_len": args.read_len, "contexts_per_gpu": n_ctx,
                       "collective": "all-reduce (sum) of %d doubles per step, %s" % (len(counts) + 1, "NCCL" if world > 1 else "single rank: no-op"),
                       "l2": "every step streams its own Forward checkpoints / row parameters, larger than L2; 2 read batches alternate"},
            "e2e": {"value": B * args.steps * world / dt, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(8 * (len(counts) + 1 + B)),
                    "note": "every step uploads its reads from host buffers and returns counts + per-read log-likelihoods to the host"},
            "gpu_launches": int(float(sums[4]) / world), "clocks": clocks,
            "gcups": {"forward_backward_wall": cu / dt / 1e9, "cells_per_step_per_gpu": cu / world / args.steps,
                      "forward_backward_kernels": cu / world / ((ms_f + ms_b) / 1e3) / 1e9,
                      "ms_forward_per_step": ms_f / args.steps, "ms_backward_per_step": ms_b / args.steps, "ms_seed_per_step": ms_s / args.steps},
            "allreduce": {"ms_per_step_max_over_ranks": float(times[1]), "doubles": len(counts) + 1, "checked": allreduce_checked},
            "roofline": {"kernel": "qg_forward/backward", "bound": "lane_instr", "achieved": cu / world / ((ms_f + ms_b) / 1e3) / 1e9,
                         "peak": PEAK_LANE_INSTR / ((INSTR_PER_CU["forward"] + INSTR_PER_CU["backward"]) / 2) / 1e9, "unit": "GCUPS",
                         "frac": cu / world / ((ms_f + ms_b) / 1e3) / (PEAK_LANE_INSTR / ((INSTR_PER_CU["forward"] + INSTR_PER_CU["backward"]) / 2)),
                         "traffic": None, "fwd_store_bytes_per_step": float(sums[5]) / world / args.steps},
            "loglike": ll,
        }
        print(json.dumps(line), flush=True)
    P.close(); G.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48, help="timed steps (default: ~4 s per arm at ~80 ms/step)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="quaff_b200")
    ap.add_argument("--reads-per-step", type=int, default=READS_PER_STEP)
    ap.add_argument("--ref-len", type=int, default=REF_LEN)
    ap.add_argument("--read-len", type=int, default=READ_LEN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the Forward/Backward side measurement")
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg5", "train", "cli"],
                    help="cfg4: the headline align benchmark (default, the contract line); cfg5: -kmatchoff full DP, 10 kb reads vs 50 kb; "
                         "train: one E-step (Forward + Backward + counts) per step over cfg4's reads with the counts all-reduce")
    ap.add_argument("--contexts", type=int, default=4, help="contexts (host thread + stream each) per GPU")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = max(args.warmup, 3) if os.environ.get("QB_ALLOW_SHORT_WARMUP") is None else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.workload == "cfg5":
        return run_cfg5(args)
    if args.workload == "cli":
        return run_cli(args)
    if args.workload == "train":
        return run_train(args)

    import torch
    import torch.distributed as dist
    from quaff_b200 import api

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU path"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.reads_per_step
    x, batches = make_workload(rank, POOL_BATCHES, B, args.ref_len, args.read_len)
    qp, nullp = load_models()
    P = api.QuaffGPUPool(device=local, n_ctx=args.contexts)
    G = P.ctxs[0]
    P.set_refs(x); P.set_params(qp)
    cfg = api.dp_config(kmer_threshold=20, band_size=64, kmer_len=6)
    flat = [api._flatten(b, True) for b in batches]
    null_ll = [np.array([api.null_loglike(nullp, r, G.L) for r in b]) for b in batches]
    # pinned host staging for the e2e arm
    pinned = []
    for tok, qual, off in flat:
        pt = torch.from_numpy(tok).pin_memory(); pq = torch.from_numpy(qual).pin_memory()
        pinned.append((pt.numpy(), pq.numpy(), off))

    # ---- resident-input arm: all pool batches live on the device as one READS set --------------------------------
    P.set_read_batches(flat)

    def resident_step(i):
        b = i % POOL_BATCHES
        return P.align_batch(cfg, b, null_ll[b])

    def pool_stats(reset=False):
        sts = P.stats(reset)
        out = {k: (sum(s[k] for s in sts) if not k.startswith("ms_") else max(s[k] for s in sts)) for k in sts[0]}
        out["ms_sum"] = {k: sum(s[k] for s in sts) for k in sts[0] if k.startswith("ms_")}
        return out

    for i in range(args.warmup):
        resident_step(i)
    sampler = ClockSampler(local); sampler.start()
    pool_stats(reset=True)
    # the K timed steps are handed to the pool at once: every context works through its share of each step without a
    # join between steps (each step's results still arrive on the host, merged, as one entry of `outs`)
    barrier(); t0 = time.perf_counter()
    outs = P.align_batches(cfg, [((args.warmup + i) % POOL_BATCHES, null_ll[(args.warmup + i) % POOL_BATCHES]) for i in range(args.steps)])
    barrier(); t1 = time.perf_counter()
    assert len(outs) == args.steps
    # a batch-0 result of the timed region (or one more untimed pass when no timed step used batch 0): diffed against the
    # reference CLI's SAM for its first reads further down
    b0 = [i for i in range(args.steps) if (args.warmup + i) % POOL_BATCHES == 0]
    parity_result = outs[b0[-1]] if b0 else P.align_batch(cfg, 0, null_ll[0])
    parity_from = "timed region" if b0 else "extra untimed pass"
    st_timed = pool_stats()
    dt = t1 - t0
    # ---- e2e arm: host buffers in, host results out, every step ---------------------------------------------------
    def e2e_step(i):
        b = i % POOL_BATCHES
        tok, qual, off = pinned[b]
        P.set_reads_raw(tok, qual, off)
        return P.align_reads(cfg, null_ll[b])

    for i in range(args.warmup):
        e2e_step(i)
    pool_stats(reset=True)
    barrier(); t2 = time.perf_counter()
    rs = P.align_stream(cfg, [pinned[(args.warmup + i) % POOL_BATCHES] + (null_ll[(args.warmup + i) % POOL_BATCHES],) for i in range(args.steps)])
    barrier(); t3 = time.perf_counter()
    r = rs[-1]
    assert len(rs) == args.steps
    st_e = pool_stats()
    clocks = sampler.stop()
    dt_e = t3 - t2
    h2d = int(pinned[0][0].nbytes + pinned[0][1].nbytes + pinned[0][2].nbytes)
    d2h = int(r["paths"].nbytes + r["score"].nbytes + r["best_ref"].nbytes + r["x_start"].nbytes + r["x_end"].nbytes + r["path_offsets"].nbytes)

    # ---- side measurement: Forward / Backward (train) GCUPS on a slice of the same reads ---------------------------
    train = None
    if not args.no_train:
        nt = min(B, 768)
        G.set_reads(batches[0][:nt])
        xi = np.tile(np.arange(2, dtype=np.uint32), nt); yi = np.repeat(np.arange(nt, dtype=np.uint32), 2)
        G.forward(cfg, xi, yi)
        G.stats(reset=True)
        G.forward(cfg, xi, yi)
        s1 = G.stats(reset=True)
        sel = xi == xi                                               # Backward for every pair of the slice
        G.backward_counts(cfg, xi[sel], yi[sel])                     # warm-up: the first call allocates the Forward store
        G.stats(reset=True)
        G.backward_counts(cfg, xi[sel], yi[sel])
        s2 = G.stats(reset=True)
        cu_f = s1["cell_updates"]; cu_b = s2["cell_updates"] / 2
        train = {"pairs": int(len(xi)), "forward_gcups": cu_f / 1e9 / (s1["ms_forward"] / 1e3),
                 "backward_gcups": cu_b / 1e9 / (s2["ms_backward"] / 1e3),
                 "fwd_bwd_gcups": 2 * cu_b / 1e9 / ((s2["ms_forward"] + s2["ms_backward"]) / 1e3),
                 "formulation": "FP64 probability space with the reference's log-sum-exp cut-off (default); the bit-exact log-space kernels are selectable (QG_OPT_FB_EXACT)"}

    # ---- kernel figures for the roofline: ONE context alone (the timed regions interleave several contexts on the GPU, so
    # their per-stream event times include interference), one step's share of the reads, CUDA events on the launching stream
    P.set_read_batches(flat)
    first, count, lo, hi = P.batch_ranges[0][0]
    G.align_reads(cfg, null_ll[0][lo:hi], first=first, count=count, split_paths=False)
    G.stats(reset=True)
    n_iso = 2
    for _ in range(n_iso):
        G.align_reads(cfg, null_ll[0][lo:hi], first=first, count=count, split_paths=False)
    st = G.stats()
    iso_reads = count * n_iso

    # ---- reduce over ranks: the slowest rank defines the step -----------------------------------------------------
    times = torch.tensor([dt, dt_e], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(st["cell_updates"]), float(st["kmer_hits"]), st["ms_seed"], st["ms_viterbi"], st["ms_traceback"],
                         float(st_timed["kernel_launches"]), float(st["trace_bytes"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        stage_max = sums.clone(); dist.all_reduce(stage_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    else:
        stage_max = sums
    dt, dt_e = float(times[0]), float(times[1])
    total_reads = B * args.steps * world
    value = total_reads / dt
    e2e_value = total_reads / dt_e

    if rank == 0:
        cu_total, hits_total = float(sums[0]), float(sums[1])
        ms_seed, ms_vit, ms_tb = float(stage_max[2]), float(stage_max[3]), float(stage_max[4])
        launches = int(float(sums[5]) / world)
        hbm_peak, peak_src = measured_peaks()
        cu_rank = cu_total / world; hits_rank = hits_total / world        # of the isolated pass (iso_reads reads)
        vit_cups = cu_rank / (ms_vit / 1e3); seed_hps = hits_rank / (ms_seed / 1e3)
        # dominant kernel by device time
        seed_dom = ms_seed >= ms_vit
        # algorithmic HBM bytes: seeding streams the 2 B/position k-mer codes of the reference once per (read, strand,
        # chunk overlap) ; the Viterbi fill writes 4 B of pointers per lane and macro-step (= trace_bytes)
        seed_bytes = 2.0 * (args.ref_len * 2) * iso_reads
        vit_bytes = float(sums[6]) / world
        # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of this same workload
        traffic = None; traffic_file = None
        try:
            traffic_file = "r02_traffic.json" if os.path.exists(os.path.join(ROOT, "profiles", "r02_traffic.json")) else "r01g_traffic.json"
            tj = json.load(open(os.path.join(ROOT, "profiles", traffic_file)))
            w = tj["workload"]
            if (w["ref_len"], w["read_len"]) == (args.ref_len, args.read_len):
                k = tj["qg_seed_kernel" if seed_dom else "qg_vit_kernel<3>"]
                # the capture's launch covered w["reads_per_launch"] reads; this run's launches cover iso_reads // n_iso
                traffic = (k["dram_bytes_read"] + k["dram_bytes_write"]) * (iso_reads // n_iso) / w["reads_per_launch"]
        except Exception:
            traffic = None
        seed_gbps = seed_bytes / (ms_seed / 1e3) / 1e9
        vit_gbps = vit_bytes / (ms_vit / 1e3) / 1e9
        seed_view = {
            "kernel": "qg_seed_kernel", "bound": "smem_atomic", "achieved": seed_hps / 1e9, "peak": PEAK_SMEM_ATOMIC / 1e9, "unit": "Ghit/s",
            "frac": seed_hps / PEAK_SMEM_ATOMIC,
            "measured_peak": MEASURED_SMEM_ATOMIC / 1e9, "frac_of_measured_peak": seed_hps / MEASURED_SMEM_ATOMIC,
            "definition": "SURVEY.md 8d / DESIGN.md 4.1: one shared-memory atomic per k-mer hit; peak = 148 SM x 32 lanes x 1965 MHz.  measured_peak = the "
                          "rate tools/ubench/atoms.cu reaches with all 32 lanes active on random counters (16 per cycle per SM; bank conflicts of a "
                          "random scatter are part of the problem)",
        }
        vit_view = {
            "kernel": "qg_vit_kernel<R>", "bound": "lane_instr", "achieved": vit_cups / 1e9, "peak": PEAK_LANE_INSTR / INSTR_PER_CU["viterbi"] / 1e9,
            "unit": "GCUPS", "frac": vit_cups / (PEAK_LANE_INSTR / INSTR_PER_CU["viterbi"]),
            "definition": "SURVEY.md 8d: 148 SM x 128 lanes x 1965 MHz lane-instructions/s over 13 instructions per cell update (FP32-minimal form; the "
                          "kernel computes in FP64 for bit-exact paths, whose issue rate is half: frac_fp64 = 2 x frac)",
        }
        roofline = dict(seed_view if seed_dom else vit_view)
        roofline.update({
            "traffic": traffic,
            "traffic_source": "profiles/%s (ncu --set full capture of this workload, scaled to this run's reads per launch)" % traffic_file,
            "algorithmic_bytes": seed_bytes / n_iso if seed_dom else vit_bytes / n_iso,
            "algorithmic_bytes_note": "per launch of the isolated pass.  seeding: 2 B k-mer code per reference position per read-strand (DESIGN.md 4.1), served "
                                      "by L2 -- DRAM traffic is ~0.2% of it, the kernel is not HBM-bound; Viterbi: the traceback pointers it writes",
            "hbm": {"seed_GBps": seed_gbps, "viterbi_trace_GBps": vit_gbps, "peak_GBps": hbm_peak, "peak_source": peak_src,
                    "seed_frac": seed_gbps / hbm_peak, "viterbi_frac": vit_gbps / hbm_peak},
            "all": {"seed": seed_view, "viterbi": vit_view,
                    "isolated_pass": {"reads": iso_reads, "ms_seed": ms_seed, "ms_viterbi_fill": ms_vit, "ms_traceback": ms_tb,
                                      "note": "one context alone after the timed regions; kernel times by CUDA events on its stream"}},
        })
        if train:
            roofline["all"]["forward_gcups"] = train["forward_gcups"]; roofline["all"]["backward_gcups"] = train["backward_gcups"]
        cpu = None
        parity_checked = 0; parity_note = "cpu_baseline leg not run"
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = max(2, min(cores, 32))
            sam_text = None
            try:
                v, secs, kind, sam_text = reference_cpu_run(sample, cores, args.ref_len, args.read_len)
                cpu = {"value": v, "unit": "reads/s", "cores": cores, "kind": kind,
                       "sample": f"{sample} reads x {args.read_len} b vs {args.ref_len} b, both strands, quaff align -threads {cores} ({secs:.1f} s)"}
            except Exception as e:                                    # the baseline is reported, never required
                cpu = {"value": None, "unit": "reads/s", "cores": cores, "kind": "unavailable", "sample": str(e)[:200]}
            if sam_text is not None:
                # the same reads went through the GPU batch: best strand, POS, CIGAR and AS:i: must be the reference's
                # (Alignment::writeSam, qmodel.cpp:611-622).  A difference is fatal: a fast wrong answer is not a result.
                from quaff_b200 import sam
                parity_checked = sam.compare_batch(sam_text, batches[0][:sample], x[0].name, len(x[0]), parity_result)
                parity_note = f"first {sample} reads of batch 0 ({parity_from}) vs the SAM of oracle/_ref/quaff: strand, POS, CIGAR, AS:i identical"
            else:
                parity_note = "no reference build on this box (oracle port timed): SAM comparison skipped"
        line = {
            "metric": "align_reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": CFG4_WORKLOAD,
                       "reads_per_step_per_gpu": B, "ref_len": args.ref_len, "read_len": args.read_len, "contexts_per_gpu": args.contexts,
                       "sharding": "reads over ranks, reference replicated, no collective on the align path",
                       "pipelining": "each context works through its share of the K timed steps without a per-step join; every step's results reach the host",
                       "l2": "inputs larger than L2: every step writes and re-reads its own ~%.1f GB of traceback pointers and alternates between %d read batches"
                             % (float(sums[6]) / world / iso_reads * B / 1e9, POOL_BATCHES)},
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": dt_e / args.steps * 1e3},
            "gpu_launches": launches,
            "clocks": clocks,
            "gcups": {"viterbi_fill": vit_cups / 1e9, "cell_updates_per_read": cu_rank / iso_reads, "kmer_hits_per_read": hits_rank / iso_reads},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "timed_region": {"wall_ms": dt * 1e3, "stage_ms_summed_over_contexts": st_timed["ms_sum"],
                             "note": "per-context CUDA-event stage times inside the timed region; their sum exceeds the wall time when the contexts' kernels overlap on the GPU"},
            "parity_checked": parity_checked, "parity_note": parity_note,
            "train": train,
        }
        print(json.dumps(line), flush=True)
    P.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
