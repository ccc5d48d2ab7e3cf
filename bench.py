#!/usr/bin/env python
"""bench.py -- motif queries/s of the folddisco hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      (CPU restatement of the reference)

Workload (BASELINE.json configs[2]): the five shipped motifs (serine peptidase, zinc finger, knottin, enolase,
aminopeptidase) replicated to a batch of 1024 queries against a human-proteome-scale synthetic database
(23 400 structures per GPU, seeded generator of folddisco_b200/synth.py), reference default flags
(-d 0.5 -a 5 --ca-distance 1.0) with --top 100.  One step = one batch through
make_query_map -> count_query (posting scan + vote) -> filter/sort/top -> candidate re-hash -> Kabsch RMSD.
With N > 1 GPUs the database grows with N (weak scaling): the index is hash-range sharded, every rank scans
its shard for the whole batch and the per-structure vote vectors are merged with one NCCL allreduce.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "motif queries/sec (batch of shipped motifs vs synthetic index; posting-list GB/s vs HBM peak in roofline)"
MOTIFS = [("query/4CHA.pdb", "B57,B102,C195"), ("query/1G2F.pdb", "F207,F212,F225,F229"),
          ("query/2N6N.pdb", "3,10,15,16,21,23,28,30"), ("query/2MNR.pdb", "164:H,195,221,247:ND,297:H"),
          ("query/1LAP.pdb", "250,255,273,332,334")]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--structs-per-gpu", type=int, default=23400)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--top", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=40, help="queries in the bounded CPU-baseline sample")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_motif_atoms():
    import fixtures as F
    atoms = F.config1_atoms()
    return [(atoms[p], q) for p, q in MOTIFS]


def database(rank, world, structs_per_gpu):
    """The same global database on every rank (weak scaling: world * structs_per_gpu structures)."""
    from folddisco_b200 import synth
    return synth.generate(structs_per_gpu * world, synth.SEED_BASE + 2)


# --------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the oracle (C++ restatement of the reference algorithm; the Rust crate cannot be built here),
    all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    import oracle_lib as O
    cores = os.cpu_count() or 1
    db = database(0, args.gpus, args.structs_per_gpu)
    from folddisco_b200 import synth
    parts = synth.split(db)
    t0 = time.time()
    comps = [O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"]) for p in parts]
    index = O.Index.build(comps, threads=cores)
    build_s = time.time() - t0
    nres = np.array([len(p["aa"]) for p in parts], np.uint64)
    plddt = np.zeros(len(parts), np.float32)
    qms, qcs = [], []
    for atoms, q in load_motif_atoms():
        s = O.Structure.from_atoms(atoms)
        ch, se, subs = O.parse_query_string(q, s.first_chain)
        qc = s.compact()
        qms.append(O.QueryMap(qc, ch, se, subs, index=index, total_structures=len(parts)))
        qcs.append(qc)
    sample = max(len(qms), args.cpu_sample)
    import ctypes as C
    maps = (O.VP * sample)(*[qms[k % len(qms)].h for k in range(sample)])
    qarr = (O.VP * sample)(*[qcs[k % len(qms)].h for k in range(sample)])
    store = (O.VP * len(comps))(*[c.h for c in comps])
    p = O.CountParams.defaults(top_n=args.top)

    def step():
        return O.lib().fdo_query_batch(maps, qarr, sample, index.h, store, len(comps), nres, plddt, C.byref(p), 0, 0,
                                       20.0, 1.0, 0, cores, None, None, None)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    qps = sample / dt
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 hashes / u8 postings / f32 idf / f64 Kabsch",
            "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": "%d queries (the five motifs cycled) per step, C++ restatement of the reference "
                                       "algorithm (oracle/), query-parallel over %d threads; index build %.1f s"
                                       % (sample, cores, build_s)},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": "configs[2]: batch of the 5 shipped motifs (replicated to %d queries) vs human-proteome-scale "
                        "synthetic index, %d structures per GPU" % (args.batch, args.structs_per_gpu),
            "structures": args.structs_per_gpu * world, "batch": args.batch, "top_n": args.top,
            "flags": "-d 0.5 -a 5 --ca-distance 1.0 --top %d, hash PDBTrRosetta 16/4 bins, cutoff 20 A" % args.top,
            "l2": "256 MiB buffer written between timed iterations (L2 flush)",
            "parallelism": "hash-range index shards x%d + NCCL vote allreduce" % world if world > 1 else "single GPU"}


# --------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import folddisco_b200 as fd
    from folddisco_b200 import host

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = fd.Context(local_rank)
    db = database(rank, world, args.structs_per_gpu)
    store = host.Store()
    store.add_soa(db)
    t0 = time.perf_counter()
    if world == 1:
        index = host.FolddiscoIndex.build(ctx, store)
        index.attach(ctx)
        sharded = None
    else:
        from folddisco_b200 import sharded as sh
        sharded = sh.ShardedIndex.build(ctx, store, rank, world)
        index = sharded.index
    build_s = time.perf_counter() - t0
    store.attach(ctx)
    hash_ms, post_ms = ctx.stage_ms("hash"), ctx.stage_ms("postings")
    motif_structs = [(host.CompactStructure.from_atoms(a), q) for a, q in load_motif_atoms()]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sp = host.SearchParams(top_n=args.top)
    stages = ("lookup", "scan", "select", "verify", "verify_edges", "verify_components", "verify_kabsch", "edges", "kabsch")

    def make_batch():
        qb = host.QueryBatch(index.params)
        qb.add_many([motif_structs[k % len(motif_structs)][0] for k in range(args.batch)],
                    [motif_structs[k % len(motif_structs)][1] for k in range(args.batch)])
        if sharded is None:
            qb.finalize(ctx)
        else:
            sharded.finalize(ctx, qb, dist)
        return qb

    def search(qb):
        if sharded is None:
            return host.search(ctx, qb, sp, labels=None)
        return sharded.search(ctx, qb, sp, dist)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- e2e: host structures -> results, every step (H2D of query descriptors, D2H of hits/edges/RMSD inside) ----
    for _ in range(args.warmup):
        search(make_batch())
    e2e_t, res = [], None
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        res = search(make_batch())
        barrier()
        e2e_t.append(time.perf_counter() - t0)
    # ---- value: query batch prepared (inputs resident), timed region = the search itself ----
    qb = make_batch()
    for _ in range(args.warmup):
        search(qb)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = ctx.kernel_launches
    st0 = {s: (ctx.stage_ms(s), ctx.stage_launches(s)) for s in stages}
    bytes_scanned = 0
    val_t = []
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        res = search(qb)
        barrier()
        val_t.append(time.perf_counter() - t0)
        bytes_scanned += ctx.last_posting_bytes
    clocks = sampler.finish()
    launches = ctx.kernel_launches - launches0
    st1 = {s: (ctx.stage_ms(s) - st0[s][0], ctx.stage_launches(s) - st0[s][1]) for s in stages}

    def reduce_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    val_s = reduce_max(sum(val_t)) / args.steps
    e2e_s = reduce_max(sum(e2e_t)) / args.steps
    if rank != 0:
        return
    hbm, peak_src = peaks()
    scan_ms = st1["scan"][0] / max(1, args.steps)          # one k3_scan launch per step
    bytes_per_launch = bytes_scanned / max(1, args.steps)
    survivors = int(res.struct_offsets[-1])
    algo_bytes = bytes_per_launch + 16 * survivors          # posting bytes + 16 B per survivor
    achieved = algo_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    n_struct_rows, n_match_rows = int(res.struct_offsets[-1]), int(res.match_offsets[-1])
    h2d, d2h = int(res.h2d_bytes), int(res.d2h_bytes)   # tallied by fdh_search from the buffers it copies
    line = {
        "metric": METRIC, "value": args.batch / val_s,
        "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": val_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 hashes / u8 postings / f32 idf / f64 Kabsch", "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": {"value": args.batch / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k3_scan (posting-list scan + vote)", "achieved": achieved, "peak": hbm,
                     "unit": "GB/s", "frac": achieved / hbm, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": scan_ms},
        "clocks": clocks,
        "stages_ms_per_step": {s: st1[s][0] / max(1, args.steps) for s in stages},
        "host_ms_per_step": res.host_ms, "search_wall_ms": res.wall_ms,
        "results_per_step": {"structure_rows": n_struct_rows, "match_rows": n_match_rows},
        "index": {"structures": len(store), "residues": int(store.num_residues), "build_s": build_s,
                  "k1_hash_ms": hash_ms, "k2_postings_ms": post_ms},
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(args, db, index)
    print(json.dumps(line))


def cpu_baseline(args, db, index):
    """The oracle (kind "port") on this box's host cores, bounded sample of the same workload."""
    import ctypes as C
    import oracle_lib as O
    from folddisco_b200 import synth
    cores = os.cpu_count() or 1
    parts = synth.split(db)
    comps = [O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"]) for p in parts]
    b = index.buffers()
    oix = O.Index.from_buffers(b.hashes, b.offsets, b.values)  # byte-identical to the oracle's own (tests)
    nres = np.array([len(p["aa"]) for p in parts], np.uint64)
    plddt = np.zeros(len(parts), np.float32)
    qms, qcs = [], []
    for atoms, q in load_motif_atoms():
        s = O.Structure.from_atoms(atoms)
        ch, se, subs = O.parse_query_string(q, s.first_chain)
        qc = s.compact()
        qms.append(O.QueryMap(qc, ch, se, subs, index=oix, total_structures=len(parts)))
        qcs.append(qc)
    sample = max(len(qms), args.cpu_sample)
    maps = (O.VP * sample)(*[qms[k % len(qms)].h for k in range(sample)])
    qarr = (O.VP * sample)(*[qcs[k % len(qms)].h for k in range(sample)])
    store = (O.VP * len(comps))(*[c.h for c in comps])
    p = O.CountParams.defaults(top_n=args.top)
    best = None
    for it in range(3):
        t0 = time.perf_counter()
        O.lib().fdo_query_batch(maps, qarr, sample, oix.h, store, len(comps), nres, plddt, C.byref(p), 0, 0, 20.0, 1.0, 0,
                                cores, None, None, None)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return {"value": sample / best, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": "%d queries (the five motifs cycled), best of 3, C++ restatement of the reference algorithm "
                      "(oracle/), query-parallel over %d threads" % (sample, cores)}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if args.gpus != 1 or world != 1:
            sys.stderr.write("bench.py: --gpus %d but WORLD_SIZE=%d; launch with torch.distributed.run\n" % (args.gpus, world))
            sys.exit(2)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
