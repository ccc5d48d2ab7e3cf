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
With N > 1 GPUs the problem grows with N (weak scaling): N x 23 400 structures, N x 1024 queries.  The index is
hash-range sharded; every rank builds the query maps of its 1024 queries, scans its shard for the WHOLE batch,
the non-empty vote cells are exchanged with one all-to-all over NVLink, and every rank finishes its own 1024 queries
(folddisco_b200/sharded.py).
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
    ap.add_argument("--batch", type=int, default=1024, help="queries per GPU")
    ap.add_argument("--top", type=int, default=100)
    ap.add_argument("--sub-batch", type=int, default=0, help="> 0: e2e through host.search_stream with this many queries per sub-batch (1 GPU)")
    ap.add_argument("--cpu-sample", type=int, default=40, help="queries in the bounded CPU-baseline sample")
    ap.add_argument("--sweep", default="2,4", help="index-size multipliers of the posting-scan sweep (N=1 only; '' = off)")
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


def scan_traffic():
    """DRAM bytes (read + write) of one k3_scan launch of this workload, from the committed ncu --set full capture"""
    p = os.path.join(ROOT, "profiles", "k3_scan_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get("dram_bytes_per_launch")
    return None


def load_motif_atoms():
    import fixtures as F
    atoms = F.config1_atoms()
    return [(atoms[p], q) for p, q in MOTIFS]


def tile_db(db, k):
    """the database repeated k times (structure ids shifted): k-fold longer posting lists, same per-hash frequencies"""
    R = int(db["row_offsets"][-1])
    ro = np.concatenate([db["row_offsets"][:-1].astype(np.uint64) + np.uint64(j * R) for j in range(k)] +
                        [np.array([k * R], np.uint64)])
    return dict(row_offsets=ro, n_xyz=np.tile(db["n_xyz"], (k, 1)), ca_xyz=np.tile(db["ca_xyz"], (k, 1)),
                cb_xyz=np.tile(db["cb_xyz"], (k, 1)), aa=np.tile(db["aa"], k))


def distinct_motifs(db, n, first, seed=0x5EED):
    """n DISTINCT motif queries sampled from the database itself: query number first + k takes structure
    (first + k) * 7919 mod S, a seeded centre residue and 3-8 residues whose C-alpha lies within 12 A of it
    (chain A, residue numbers = position + 1: the labels CompactStructure.from_soa gives).
    -> list of (structure number, residue positions, query string)"""
    ro = db["row_offsets"].astype(np.int64)
    S = len(ro) - 1
    out = []
    for k in range(first, first + n):
        rng = np.random.Generator(np.random.PCG64([seed, k]))
        s = (k * 7919) % S
        ca = db["ca_xyz"][ro[s]:ro[s + 1]]
        while True:
            c = int(rng.integers(0, len(ca)))
            near = np.flatnonzero(np.linalg.norm(ca - ca[c], axis=1) <= 12.0)
            near = near[db["aa"][ro[s]:ro[s + 1]][near] < 20]
            if len(near) >= 3:
                break
        m = int(min(len(near), rng.integers(3, 9)))
        pick = np.sort(rng.choice(near, m, replace=False))
        out.append((s, pick, ",".join("A%d" % (i + 1) for i in pick)))
    return out


def make_query_batch(ctx, index, db, n, first, sharded=None, dist=None, timing=None):
    """host.QueryBatch of n queries starting at global query number `first`: distinct motifs sampled from db, or (db is
    None) the five shipped motifs cycled"""
    from folddisco_b200 import host
    t0 = time.perf_counter()
    qb = host.QueryBatch(index.params)
    if db is None:
        motif_structs = [(host.CompactStructure.from_atoms(a), q) for a, q in load_motif_atoms()]
        which = np.arange(first, first + n, dtype=np.uint32) % len(motif_structs)
        qb.add_many_indexed([m[0] for m in motif_structs], [m[1] for m in motif_structs], which, which)
    else:
        ro = db["row_offsets"].astype(np.int64)
        motifs = distinct_motifs(db, n, first)
        comps = [host.CompactStructure.from_soa(db["n_xyz"][ro[s]:ro[s + 1]], db["ca_xyz"][ro[s]:ro[s + 1]],
                                                db["cb_xyz"][ro[s]:ro[s + 1]], db["aa"][ro[s]:ro[s + 1]])
                 for s, _, _ in motifs]
        idx = np.arange(n, dtype=np.uint32)
        qb.add_many_indexed(comps, [m[2] for m in motifs], idx, idx)
    t1 = time.perf_counter()
    if sharded is None:
        qb.finalize(ctx)
    else:
        sharded.prepare(ctx, qb, dist)
    if timing is not None:
        timing["query_maps"] += (t1 - t0) * 1e3
        timing["finalize"] += (time.perf_counter() - t1) * 1e3
        timing["calls"] += 1
    return qb


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
    # CPU index build is ~30 s per 23 400 structures: beyond 4 GPUs' worth the reference arm searches a 4x database
    # (a smaller index only makes the CPU arm faster)
    ref_gpus = min(args.gpus, 4)
    db = database(0, ref_gpus, args.structs_per_gpu)
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
                                       "algorithm (oracle/), query-parallel over %d threads; index of %d structures, build %.1f s"
                                       % (sample, cores, len(parts), build_s)},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "configs[2]: batch of the 5 shipped motifs (replicated to %d queries per GPU) vs "
                        "human-proteome-scale synthetic index, %d structures per GPU" % (args.batch, args.structs_per_gpu),
            "structures": args.structs_per_gpu * world, "batch": args.batch * world, "top_n": args.top,
            "flags": "-d 0.5 -a 5 --ca-distance 1.0 --top %d, hash PDBTrRosetta 16/4 bins, cutoff 20 A" % args.top,
            "l2": "256 MiB buffer written between timed iterations (L2 flush)",
            "parallelism": ("hash-range index shards x%d, every rank scans its shard for the whole batch, sparse vote "
                            "all-to-all (NCCL), every rank finishes %d queries" % (world, args.batch))
            if world > 1 else "single GPU"}


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
    hv = ("hv_flatten", "hv_upload", "hv_candidates", "hv_issue", "hv_wait_chunks", "hv_copy_tail")  # host wall clock inside the verification call

    prep_ms = {"query_maps": 0.0, "finalize": 0.0, "calls": 0}  # host wall clock of the e2e-only part of a step

    def make_batch():
        """this rank's args.batch queries of the global batch (query number q uses motif q mod 5)"""
        t0 = time.perf_counter()
        qb = host.QueryBatch(index.params)
        which = np.arange(rank * args.batch, (rank + 1) * args.batch, dtype=np.uint32) % len(motif_structs)
        # every query map is built on its own; the indexed call only spares the marshalling of 2 x batch Python objects
        qb.add_many_indexed([m[0] for m in motif_structs], [m[1] for m in motif_structs], which, which)
        t1 = time.perf_counter()
        if sharded is None:
            qb.finalize(ctx)
        else:
            sharded.prepare(ctx, qb, dist)
        t2 = time.perf_counter()
        prep_ms["query_maps"] += (t1 - t0) * 1e3
        prep_ms["finalize"] += (t2 - t1) * 1e3
        prep_ms["calls"] += 1
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
    def timed(fn):
        """one step bracketed by barrier + synchronize and by CUDA events on the current stream (the library call is
        synchronous, so the event interval covers host orchestration, copies and kernels of the step)"""
        flush.fill_(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        return out, e0.elapsed_time(e1) * 1e-3, wall

    def e2e_step():
        """host structures -> result rows through the public API: build the query batch, search it.  --sub-batch N > 0
        (one GPU) uses host.search_stream instead (sub-batches, the next one prepared on a second host thread while
        the current one is searched); measured slower at this batch size (per-call fixed costs: 13.1 ms at 512,
        14.8 ms at 256 vs 12.3 ms), so it is off by default"""
        if sharded is not None or args.sub_batch <= 0:
            return [search(make_batch())]
        ks = range(rank * args.batch, (rank + 1) * args.batch)
        return host.search_stream(ctx, [motif_structs[k % len(motif_structs)][0] for k in ks],
                                  [motif_structs[k % len(motif_structs)][1] for k in ks], sp, index.params,
                                  sub_batch=args.sub_batch)

    for _ in range(args.warmup):
        e2e_step()
    e2e_t, e2e_res = [], None
    for _ in range(args.steps):
        e2e_res, dt, _ = timed(e2e_step)
        e2e_t.append(dt)
    e2e_h2d, e2e_d2h = sum(int(r.h2d_bytes) for r in e2e_res), sum(int(r.d2h_bytes) for r in e2e_res)
    e2e_rows = (sum(int(r.struct_offsets[-1]) for r in e2e_res), sum(int(r.match_offsets[-1]) for r in e2e_res))
    del e2e_res
    # ---- value: query batch prepared (inputs resident), timed region = the search itself ----
    qb = make_batch()
    for _ in range(args.warmup):
        search(qb)
    if sharded is not None:
        sharded.merge_ms, sharded.merge_bytes = 0.0, 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = ctx.kernel_launches
    st0 = {s: (ctx.stage_ms(s), ctx.stage_launches(s)) for s in stages + hv}
    bytes_scanned = 0
    val_t = []
    wall_t = []
    for _ in range(args.steps):
        res, dt, wall = timed(lambda: search(qb))
        val_t.append(dt)
        wall_t.append(wall)
        bytes_scanned += ctx.last_posting_bytes
    clocks = sampler.finish()
    launches = ctx.kernel_launches - launches0
    st1 = {s: (ctx.stage_ms(s) - st0[s][0], ctx.stage_launches(s) - st0[s][1]) for s in stages + hv}

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
    if world > 1:  # every rank scanned its own shard: the launch's algorithmic bytes are this rank's
        survivors = 0
    n_struct_rows, n_match_rows = int(res.struct_offsets[-1]), int(res.match_offsets[-1])
    # tallied by fdh_search from the buffers it copies (rank 0's slice; every rank moves the same amount)
    h2d, d2h = e2e_h2d * world, e2e_d2h * world
    assert e2e_rows == (n_struct_rows, n_match_rows), (e2e_rows, n_struct_rows, n_match_rows)  # same rows either way
    line = {
        "metric": METRIC, "value": args.batch * world / val_s,
        "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": val_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 hashes / u8 postings / f32 idf / f64 Kabsch", "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": {"value": args.batch * world / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k3_scan (posting-list scan + vote)", "achieved": achieved, "peak": hbm,
                     "unit": "GB/s", "frac": achieved / hbm, "traffic": scan_traffic() if world == 1 else None,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": scan_ms,
                     "launches_per_step": "one per edge-word class of the batch (2 for the five shipped motifs); "
                                          "bytes and ms are per step",
                     "note": "at this scale the scan is bound by shared-memory vote traffic, CTA latency and issue "
                             "slots, not by HBM (a query's vote state is ~20x its posting bytes): DESIGN.md section 4"},
        "clocks": clocks,
        "stages_ms_per_step": {s: st1[s][0] / max(1, args.steps) for s in stages},
        "host_ms_per_step": res.host_ms, "search_wall_ms": res.wall_ms,
        "verify_host_wall_ms": {s[3:]: st1[s][0] / max(1, args.steps) for s in hv},
        "e2e_prepare_host_ms": {"query_maps (make_query_map x batch)": prep_ms["query_maps"] / max(1, prep_ms["calls"]),
                                "finalize (posting counts -> idf, verification tables -> device)":
                                    prep_ms["finalize"] / max(1, prep_ms["calls"])},
        "timing": "CUDA events around each step (max over ranks); wall-clock cross-check %.3f ms/step" % (
            1e3 * sum(wall_t) / max(1, args.steps)),
        "results_per_step": {"structure_rows": n_struct_rows, "match_rows": n_match_rows},
        "index": {"structures": len(store), "residues": int(store.num_residues), "build_s": build_s,
                  "k1_hash_ms": hash_ms, "k2_postings_ms": post_ms},
    }
    if sharded is not None:
        n_search = max(1, args.steps)
        line["vote_merge"] = {"collective": "NCCL all_to_all of the non-empty vote cells (+ all_gather of the counts)",
                              "ms_per_step": sharded.merge_ms / n_search,
                              "bytes_sent_per_rank_per_step": sharded.merge_bytes // n_search,
                              "results": "each rank finishes its own %d queries; rows stay on the rank that produced them" % args.batch}
    if world == 1:
        if args.sweep:
            line["scan_vs_index_size"] = scan_sweep(args, ctx, db, qb, sp, hbm, bytes_per_launch + 16 * survivors, scan_ms)
            index.attach(ctx)
        line["cpu_baseline"] = cpu_baseline(args, db, index)
    print(json.dumps(line), flush=True)


def scan_sweep(args, ctx, db, qb, sp, hbm, base_bytes, base_ms):
    """Posting-list GB/s of k3_scan vs index size: the database tiled k times (ids shifted), index rebuilt on the GPU,
    count_query only (skip_match), same batch.  Larger indexes have longer lists: more bytes per launch."""
    import torch
    from folddisco_b200 import host
    out = [{"structures": args.structs_per_gpu, "algorithmic_bytes_per_launch": base_bytes, "scan_ms": base_ms,
            "GBps": base_bytes / (base_ms * 1e-3) / 1e9, "frac": base_bytes / (base_ms * 1e-3) / 1e9 / hbm}]
    skip = host.SearchParams(top_n=args.top, skip_match=True)
    for k in [int(x) for x in args.sweep.split(",") if x]:
        S = len(db["row_offsets"]) - 1
        R = int(db["row_offsets"][-1])
        ro = np.concatenate([db["row_offsets"][:-1].astype(np.uint64) + np.uint64(j * R) for j in range(k)] +
                            [np.array([k * R], np.uint64)])
        tiled = dict(row_offsets=ro, n_xyz=np.tile(db["n_xyz"], (k, 1)), ca_xyz=np.tile(db["ca_xyz"], (k, 1)),
                     cb_xyz=np.tile(db["cb_xyz"], (k, 1)), aa=np.tile(db["aa"], k))
        store = host.Store()
        store.add_soa(tiled)
        del tiled
        ix = host.FolddiscoIndex.build(ctx, store)
        ix.attach(ctx)
        for _ in range(2):
            host.search(ctx, qb, skip)
        ms0, n = ctx.stage_ms("scan"), 3
        nbytes = 0
        for _ in range(n):
            r = host.search(ctx, qb, skip)
            nbytes += ctx.last_posting_bytes + 16 * int(r.struct_offsets[-1])
        ms = (ctx.stage_ms("scan") - ms0) / n
        gbps = nbytes / n / (ms * 1e-3) / 1e9
        out.append({"structures": S * k, "algorithmic_bytes_per_launch": nbytes / n, "scan_ms": ms, "GBps": gbps,
                    "frac": gbps / hbm})
        del ix, store
    return out


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
    # the contract is ONE JSON line on stdout: libraries (NCCL prints its version banner) get stderr instead
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:  # the ranks share the box's cores: the library's host-side steps get cores / ranks threads each
        os.environ.setdefault("FD_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // world)))
    if world != args.gpus:
        if args.gpus != 1 or world != 1:
            sys.stderr.write("bench.py: --gpus %d but WORLD_SIZE=%d; launch with torch.distributed.run\n" % (args.gpus, world))
            sys.exit(2)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
