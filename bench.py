#!/usr/bin/env python
"""bench.py -- motif queries/s of the folddisco hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      (CPU restatement of the reference)

Workload (BASELINE.json configs[2] scale): a batch of 1024 DISTINCT motif queries per GPU (3-8 residues within 12 A,
sampled with a fixed seed from the database's own structures; `distinct_motifs`) against a human-proteome-scale
synthetic database (23 400 structures per GPU, seeded generator of folddisco_b200/synth.py), reference default flags
(-d 0.5 -a 5 --ca-distance 1.0) with --top 100.  One step = one batch through
make_query_map -> count_query (posting scan + vote) -> filter/sort/top -> candidate re-hash -> Kabsch RMSD -> rows.
The batch of the five shipped motifs replicated 205 x (round 1's headline) is kept as the secondary `shipped_motifs`.

With N > 1 GPUs the problem grows with N (weak scaling): N x 23 400 structures, N x 1024 queries.  The structures are
split into contiguous ID RANGES, one per rank; every rank builds the query maps of its own 1024 queries, scans its
local index for the WHOLE batch (global list lengths all-reduced, so the idf weights are the unsharded ones), one
NCCL all-to-all of fixed-size per-query top-n blocks goes to the query's owner, which merges and verifies its own
queries.  All collectives run inside libfolddisco_b200.so (csrc/fd_comm.cu); torch.distributed only carries the
barrier / max-over-ranks of the timing and the 128-byte NCCL id.

Every line carries `parity_check`: the first queries of rank 0's slice, answered by the CPU oracle on the same
database, are compared row for row with the GPU's rows of the timed configuration.
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

METRIC = "motif queries/sec (batch of distinct motifs vs synthetic index; posting-list GB/s vs HBM peak in roofline)"
MOTIFS = [("query/4CHA.pdb", "B57,B102,C195"), ("query/1G2F.pdb", "F207,F212,F225,F229"),
          ("query/2N6N.pdb", "3,10,15,16,21,23,28,30"), ("query/2MNR.pdb", "164:H,195,221,247:ND,297:H"),
          ("query/1LAP.pdb", "250,255,273,332,334")]
DTYPE = "u32 hashes / u8 postings / f32 idf / f64 Kabsch"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--structs-per-gpu", type=int, default=23400)
    ap.add_argument("--batch", type=int, default=1024, help="queries per GPU")
    ap.add_argument("--top", type=int, default=100)
    ap.add_argument("--parity-queries", type=int, default=40, help="queries of rank 0's slice diffed against the oracle")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries per step of the CPU arm (0 = 16 x cores)")
    ap.add_argument("--sweep", default="4", help="index-size multipliers of the posting-scan sweep (N=1 only; '' = off; "
                                                 "23 = Swiss-Prot scale, 538 200 structures, about two more minutes)")
    ap.add_argument("--pair-table", type=int, default=1, help="build the structure store's pair table (verification by "
                                                             "hash lookup instead of re-hashing candidates)")
    ap.add_argument("--extras", type=int, default=1, help="secondary block: K1 per encoding, metrics / partial-fit step cost")
    ap.add_argument("--pipeline", type=int, default=1, help="e2e as a serving loop: query maps of the next batch built on a "
                                                            "second host thread while the current batch is searched (0 = "
                                                            "report only the one-batch-at-a-time step)")
    ap.add_argument("--shipped", type=int, default=1, help="also time the five shipped motifs x 205 (secondary line)")
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
    """DRAM bytes (read + write) of one scan launch of this workload, from the committed ncu --set full capture"""
    p = os.path.join(ROOT, "profiles", "k3_scan_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get("dram_bytes_per_launch")
    return None


def scan_traffic_source():
    p = os.path.join(ROOT, "profiles", "k3_scan_traffic.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return "ncu --set full capture of this command, profiles/%s (%s launches)" % (d.get("source"), d.get("launches_captured"))
    return None


def verify_block(st1, steps):
    """the kernels the step's time goes to (the posting scan is ~6 % of it): live stage times per step beside the
    issue-slot figures of the committed ncu capture of this command (profiles/k6_ncu_summary.json)"""
    out = {"ms_per_step": {"k6a (verify_edges)": st1["verify_edges"][0] / steps,
                           "k6b (verify_components)": st1["verify_components"][0] / steps,
                           "k6c (verify_kabsch)": st1["verify_kabsch"][0] / steps,
                           "k6d (rows)": st1["rows"][0] / steps, "pipeline wall (verify)": st1["verify"][0] / steps},
           "bound": "latency / issue slots (graph components, residue mapping and rescue are warp-serial integer work; "
                    "the pair-table lookups of k6a are dependent gathers)"}
    p = os.path.join(ROOT, "profiles", "k6_ncu_summary.json")
    if os.path.exists(p):
        out["ncu"] = json.load(open(p))
    return out


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


def query_inputs(db, n, first):
    """the host-side INPUTS of a batch, as a user holds them: (CompactStructure objects, query strings) of n queries
    starting at global query number `first` -- distinct motifs sampled from db, or (db is None) the five shipped
    motifs cycled"""
    from folddisco_b200 import host
    if db is None:
        motif_structs = [(host.CompactStructure.from_atoms(a), q) for a, q in load_motif_atoms()]
        which = np.arange(first, first + n, dtype=np.uint32) % len(motif_structs)
        return [m[0] for m in motif_structs], [m[1] for m in motif_structs], which, which
    ro = db["row_offsets"].astype(np.int64)
    motifs = distinct_motifs(db, n, first)
    comps = [host.CompactStructure.from_soa(db["n_xyz"][ro[s]:ro[s + 1]], db["ca_xyz"][ro[s]:ro[s + 1]],
                                            db["cb_xyz"][ro[s]:ro[s + 1]], db["aa"][ro[s]:ro[s + 1]])
             for s, _, _ in motifs]
    idx = np.arange(n, dtype=np.uint32)
    return comps, [m[2] for m in motifs], idx, idx


def make_query_batch(ctx, index, db, n, first, shards=None, timing=None, inputs=None):
    """host.QueryBatch from the batch's inputs: make_query_map of every query (add_many_indexed), then the idf lookup
    and the upload of the verification tables (finalize; with id-range shards the collective form)"""
    from folddisco_b200 import host
    if inputs is None:
        inputs = query_inputs(db, n, first)
    t0 = time.perf_counter()
    qb = host.QueryBatch(index.params)
    if isinstance(inputs, host.QueryInputs):
        qb.add_prepared(inputs)  # the inputs are already C arrays of structure handles / query strings
    else:
        qb.add_many_indexed(*inputs)
    t1 = time.perf_counter()
    if shards is None:
        qb.finalize(ctx)
    else:
        shards.prepare(ctx, qb)
    if timing is not None:
        timing["query_maps"] += (t1 - t0) * 1e3
        timing["finalize"] += (time.perf_counter() - t1) * 1e3
        timing["calls"] += 1
    return qb


def database(world, structs_per_gpu):
    """The same global database on every rank (weak scaling: world * structs_per_gpu structures)."""
    from folddisco_b200 import synth
    return synth.generate(structs_per_gpu * world, synth.SEED_BASE + 2)


def workload_config(args, world):
    return {"workload": "configs[2] scale: batch of %d DISTINCT motifs per GPU (3-8 residues within 12 A, sampled from "
                        "the database) vs human-proteome-scale synthetic index, %d structures per GPU"
                        % (args.batch, args.structs_per_gpu),
            "structures": args.structs_per_gpu * world, "batch": args.batch * world, "top_n": args.top,
            "flags": "-d 0.5 -a 5 --ca-distance 1.0 --top %d, hash PDBTrRosetta 16/4 bins, cutoff 20 A" % args.top,
            "l2": "256 MiB buffer written between timed iterations (L2 flush)",
            "parallelism": ("id-range index shards x%d: every rank scans its local index for the whole batch, one NCCL "
                            "all-to-all of per-query top-%d blocks, every rank merges and verifies its own %d queries"
                            % (world, args.top, args.batch)) if world > 1 else "single GPU"}


# --------------------------------------------------------------------------------------------------
class Oracle:
    """The CPU oracle over the bench database (checker / CPU arm; never on the measured GPU path)."""

    def __init__(self, db, index_buffers=None, threads=1):
        import oracle_lib as O
        from folddisco_b200 import synth
        self.O, self.db = O, db
        self.parts = synth.split(db)
        self.S = len(self.parts)
        self.nres = np.array([len(p["aa"]) for p in self.parts], np.uint64)
        self.plddt = np.zeros(self.S, np.float32)
        self._comps = {}
        t0 = time.perf_counter()
        if index_buffers is not None:  # byte-identical to the oracle's own build (tests/test_gpu_parity.py)
            self.index = O.Index.from_buffers(index_buffers.hashes, index_buffers.offsets, index_buffers.values)
        else:
            self.index = O.Index.build([self.comp(s) for s in range(self.S)], threads=threads)
        self.build_s = time.perf_counter() - t0

    def comp(self, s):
        c = self._comps.get(s)
        if c is None:
            p = self.parts[s]
            c = self.O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"],
                                        serial=np.arange(1, len(p["aa"]) + 1, dtype=np.uint64))
            self._comps[s] = c
        return c

    def query_maps(self, first, n):
        """oracle query maps of the distinct motifs first .. first + n"""
        O = self.O
        out = []
        for s, pick, qstr in distinct_motifs(self.db, n, first):
            ch, se, subs = O.parse_query_string(qstr, ord("A"))
            out.append(O.QueryMap(self.comp(s), ch, se, subs, index=self.index, total_structures=self.S))
        return out

    def parity(self, res, first, n, top_n):
        """diff of the GPU rows (host.Results, query k = global query first + k) against the oracle's"""
        import parity
        outer = self

        class Lazy(dict):
            def __missing__(self, k):
                return outer.comp(k)

        comps = Lazy()
        bad = []
        rows_checked = 0
        for k, om in enumerate(self.query_maps(first, n)):
            hits, rows = parity.oracle_query(om, self.index, comps, self.nres, self.plddt, top_n=top_n)
            bad += parity.diff_query(res, k, len(om.indices()), hits, rows, top_n=top_n)
            rows_checked += len(hits["nid"]) + len(rows)
        return {"queries": n, "rows_checked": rows_checked, "mismatches": len(bad), "first_mismatches": bad[:5],
                "checker": "CPU oracle (oracle/) on the same database; integer fields and residues exact, idf / RMSD 1e-4"}

    def batch_qps(self, first, n, top_n, threads, repeats=3):
        """count_query + retrieval of n distinct queries on `threads` host threads: best-of-repeats queries/s"""
        import ctypes as C
        O = self.O
        qms = self.query_maps(first, n)
        maps = (O.VP * n)(*[q.h for q in qms])
        qarr = (O.VP * n)(*[q.query.h for q in qms])
        store = (O.VP * self.S)(*[self.comp(s).h for s in range(self.S)])
        p = O.CountParams.defaults(top_n=top_n)
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.lib().fdo_query_batch(maps, qarr, n, self.index.h, store, self.S, self.nres, self.plddt, C.byref(p), 0, 0,
                                    20.0, 1.0, 0, threads, None, None, None)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        return n / best, best


def run_reference(args, rank, world):
    """CPU arm: the oracle (C++ restatement of the reference algorithm; the Rust crate cannot be built here), all host
    threads, same database size and the same distinct-motif workload as the GPU arm; each step a bounded sample."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    db = database(args.gpus, args.structs_per_gpu)
    orc = Oracle(db, threads=cores)
    sample = args.cpu_sample or 16 * cores
    for _ in range(max(0, min(args.warmup, 2) - 1)):
        orc.batch_qps(0, sample, args.top, cores, repeats=1)
    times = []
    for _ in range(max(1, min(args.steps, 5))):
        times.append(orc.batch_qps(0, sample, args.top, cores, repeats=1)[1])
    dt = float(np.mean(times))
    qps = sample / dt
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": "%d distinct queries per step (the first of the GPU arm's batch; %d steps timed), C++ "
                                       "restatement of the reference algorithm (oracle/, -O3), query-parallel over %d "
                                       "threads; index of %d structures built on the CPU in %.1f s"
                                       % (sample, len(times), cores, orc.S, orc.build_s)},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import folddisco_b200 as fd
    from folddisco_b200 import host, sharded

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = fd.Context(local_rank)
    db = database(world, args.structs_per_gpu)
    t0 = time.perf_counter()
    if world == 1:
        store = host.Store()
        store.add_soa(db)
        index = host.FolddiscoIndex.build(ctx, store)
        index.attach(ctx)
        build_s = time.perf_counter() - t0
        t1 = time.perf_counter()
        table_bytes = store.attach(ctx, pair_table=args.pair_table != 0, hash_params=index.params)
        table_s = time.perf_counter() - t1
        shards = None
    else:
        sharded.comm_init(ctx, rank, world, dist)
        shards, store = sharded.IdRangeShards.build(ctx, db, rank, world, pair_table=args.pair_table != 0)
        index = shards.index
        build_s = time.perf_counter() - t0
        table_bytes, table_s = shards.table_bytes, shards.table_s
    hash_ms, post_ms = ctx.stage_ms("hash"), ctx.stage_ms("postings")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sp = host.SearchParams(top_n=args.top)
    stages = ("lookup", "scan", "select", "exchange", "merge", "verify", "verify_edges", "verify_components",
              "verify_kabsch", "rows", "edges", "kabsch")
    hv = ("hv_flatten", "hv_upload", "hv_candidates", "hv_issue", "hv_wait_chunks", "hv_copy_tail", "cq_host_prepare",
          "cq_host_rest")
    prep_ms = {"query_maps": 0.0, "finalize": 0.0, "calls": 0}  # host wall clock of the e2e-only part of a step

    # host structures + query strings: the step's inputs, held the way a C / Rust host holds them (arrays of structure
    # handles and C strings; built once -- converting Python lists into those arrays costs 1.2 ms per batch and is the test
    # harness's language, not the library's work)
    inputs = host.QueryInputs(*query_inputs(db, args.batch, rank * args.batch))

    def make_batch(timing=prep_ms):
        return make_query_batch(ctx, index, db, args.batch, rank * args.batch, shards, timing, inputs)

    def search(qb):
        if shards is None:
            return host.search(ctx, qb, sp, labels=store)
        return shards.search(ctx, qb, sp, labels=store)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

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

    def reduce_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- e2e: host structures -> result rows through the public API, every step (H2D / D2H inside) ----
    e2e_res = None
    for _ in range(args.warmup):  # the previous result stays alive during a step, exactly as in the timed loop: the
        e2e_res = search(make_batch())  # second set of page-locked result blocks is allocated here, not under the clock
    e2e_t = []
    e2e_names = ("hv_flatten", "hv_upload", "lookup", "cq_host_prepare", "cq_host_rest", "fs_flatten", "fs_allgather",
                 "fs_parse", "fs_counts", "fs_allreduce", "fs_tables")
    e2e_s0 = {k: ctx.stage_ms(k) for k in e2e_names}
    for _ in range(args.steps):
        e2e_res, dt, _ = timed(lambda: search(make_batch()))
        e2e_t.append(dt)
    e2e_stage = {k: (ctx.stage_ms(k) - e2e_s0[k]) / max(1, args.steps) for k in e2e_names}
    e2e_h2d, e2e_d2h = int(e2e_res.h2d_bytes), int(e2e_res.d2h_bytes)
    e2e_rows = (int(e2e_res.struct_offsets[-1]), int(e2e_res.match_offsets[-1]))
    # ---- e2e as a serving loop (host.QueryMapWorker): make_query_map of batch k+1 runs on a second host thread while
    # batch k is finalized and searched.  Every timed step still holds ALL the work of one batch -- one make_query_map x
    # batch (waited for before the clock stops), one finalize, one search with its H2D / D2H -- only overlapped.
    pipe_t, pipe_err = [], None
    if args.pipeline:
        def finalize(qb_k):
            if shards is None:
                qb_k.finalize(ctx)
            else:
                shards.prepare(ctx, qb_k)
        try:
            pipe_t, e2e_res = serving_loop(host, index.params, inputs, finalize, search, timed, args.warmup, args.steps,
                                           e2e_rows)
        except Exception as e:  # the serving-loop figure never costs the line: e2e falls back to one batch at a time
            pipe_t, pipe_err = [], "%s: %s" % (type(e).__name__, e)
    pipe_fail = 1.0 if pipe_err is not None else 0.0
    del e2e_res
    # ---- value: query batch prepared (inputs resident), timed region = the search itself ----
    qb = make_batch(timing=None)
    res = None
    for _ in range(args.warmup):
        res = search(qb)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = ctx.kernel_launches
    general0 = ctx.stage_launches("general_candidates")
    gnames = ("general_reason_query", "general_reason_edges", "general_reason_nodes", "general_reason_components",
              "general_reason_lists")
    greason0 = {k: ctx.stage_launches(k) for k in gnames}
    gwall0 = ctx.stage_ms("general_path_wall")
    st0 = {s: (ctx.stage_ms(s), ctx.stage_launches(s)) for s in stages + hv}
    bytes_scanned, exch_bytes = 0, 0
    val_t, wall_t = [], []
    for _ in range(args.steps):
        res, dt, wall = timed(lambda: search(qb))
        val_t.append(dt)
        wall_t.append(wall)
        bytes_scanned += ctx.last_posting_bytes
        exch_bytes += ctx.last_exchange_bytes if shards is not None else 0
    clocks = sampler.finish()
    launches = ctx.kernel_launches - launches0
    st1 = {s: (ctx.stage_ms(s) - st0[s][0], ctx.stage_launches(s) - st0[s][1]) for s in stages + hv}
    steps = max(1, args.steps)
    val_s = reduce_max(sum(val_t)) / steps
    e2e_s = reduce_max(sum(e2e_t)) / steps
    pipe_ok = args.pipeline and reduce_max(pipe_fail) == 0.0  # every rank's loop ran (collective: all ranks call it)
    pipe_s = reduce_max(sum(pipe_t) if pipe_t else 0.0) / steps if args.pipeline else 0.0
    scan_ms = st1["scan"][0] / steps
    scan_ms_max = reduce_max(scan_ms)
    bytes_per_launch = bytes_scanned / steps
    bytes_max = reduce_max(bytes_per_launch)
    exch_ms_max = reduce_max(st1["exchange"][0] / steps)
    n_struct_rows, n_match_rows = int(res.struct_offsets[-1]), int(res.match_offsets[-1])
    assert e2e_rows == (n_struct_rows, n_match_rows), (e2e_rows, n_struct_rows, n_match_rows)  # same rows either way
    if rank != 0:
        return
    hbm, peak_src = peaks()
    # posting bytes + 16 B per survivor.  With id-range shards every rank scans for the WHOLE batch and
    # keeps its top n per query: the launch's algorithmic bytes are this rank's (max over ranks reported beside)
    survivors = n_struct_rows if world == 1 else args.batch * world * args.top
    algo_bytes = bytes_per_launch + 16 * survivors
    achieved = algo_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": args.batch * world / val_s,
        "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": val_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": e2e_block(args, world, e2e_s, pipe_s if pipe_ok else None, pipe_err, e2e_h2d, e2e_d2h),
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k3_scan_v3 (posting-list scan + vote + tile-level top-n)",
                     "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": scan_traffic() if world == 1 else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": scan_ms,
                     "launch_ms_max_over_ranks": scan_ms_max,
                     "algorithmic_bytes_max_over_ranks": bytes_max + 16 * survivors,
                     "traffic_source": scan_traffic_source() if world == 1 else None,
                     "note": "one launch per step.  Not HBM-bound and cannot be: a 1.3-byte posting costs ~25 lane "
                             "instructions of LEB128 decode and two shared-memory atomics (measured ceiling 5.5 votes / "
                             "clock / SM = 0.16 of the copy peak), and a motif query votes into twice as many cells as it "
                             "has postings; traffic exceeds the algorithmic bytes at this scale because lists average "
                             "~270 bytes (64-byte granules + 16 bytes of look-ahead, 32-byte sectors): DESIGN.md section 4"},
        "verify_kernels": verify_block(st1, steps),
        "clocks": clocks,
        "stages_ms_per_step": {s: st1[s][0] / steps for s in stages},
        "host_ms_per_step": res.host_ms, "search_wall_ms": res.wall_ms,
        "verify_host_wall_ms": {s[3:]: st1[s][0] / steps for s in hv if s.startswith("hv_")},
        "count_query_host_wall_ms": {s[8:]: st1[s][0] / steps for s in hv if s.startswith("cq_host_")},
        "e2e_prepare_host_ms": {"query_maps (make_query_map x batch)": prep_ms["query_maps"] / max(1, prep_ms["calls"]),
                                "finalize (posting counts -> idf, verification tables -> device)":
                                    prep_ms["finalize"] / max(1, prep_ms["calls"])},
        "e2e_stage_ms_per_step": {k: v for k, v in e2e_stage.items() if v > 0},
        "timing": "CUDA events around each step (max over ranks); wall-clock cross-check %.3f ms/step" % (
            1e3 * sum(wall_t) / steps),
        "results_per_step": {"structure_rows": n_struct_rows, "match_rows": n_match_rows,
                             "candidates_on_the_general_verification_path":
                                 (ctx.stage_launches("general_candidates") - general0) // steps,
                             "general_path_reasons": {k[15:]: (ctx.stage_launches(k) - greason0[k]) // steps for k in gnames},
                             "general_path_wall_ms": (ctx.stage_ms("general_path_wall") - gwall0) / steps},
    }
    if shards is not None:
        fs = ("fs_flatten", "fs_allgather", "fs_parse", "fs_counts", "fs_allreduce", "fs_tables")
        line["e2e_prepare_host_ms"]["finalize_sharded breakdown (mean per call)"] = {
            k[3:]: ctx.stage_ms(k) / max(1, ctx.stage_launches(k)) for k in fs}
        line["exchange"] = {"collective": "ncclSend/ncclRecv group (all-to-all) of per-query top-%d blocks + counts, "
                                          "inside fd_count_query_sharded" % args.top,
                            "ms_per_step_max_over_ranks": exch_ms_max,
                            "bytes_sent_per_rank_per_step": exch_bytes // steps}
    # ---- parity: the timed configuration's rows against the CPU oracle (rank 0's first queries) ----
    t0 = time.perf_counter()
    if world == 1:
        full_buffers = index.buffers()
    else:  # the oracle needs the whole index: built once more on this GPU from the full store (not timed)
        full_buffers = host.FolddiscoIndex.build(ctx, store).buffers()
    orc = Oracle(db, index_buffers=full_buffers)
    line["parity_check"] = orc.parity(res, rank * args.batch, min(args.parity_queries, args.batch), args.top)
    line["parity_check"]["seconds"] = time.perf_counter() - t0
    # ---- index build (path (i)) ----
    line["index_build"] = index_build_block(ctx, host, store if world == 1 else None, db, orc, build_s, hash_ms, post_ms,
                                            hbm)
    line["pair_table"] = {"bytes": table_bytes, "build_s": table_s, "device_ms": ctx.stage_ms("pair_table"),
                          "what": "per structure the (hash, i, j) of every hashed residue pair, sorted by hash (8 B each): "
                                  "verification looks query hashes up instead of re-hashing candidates; built once at "
                                  "store attach, outside the timed region (like the index)"}
    if world == 1:
        if args.shipped:
            line["shipped_motifs"] = shipped_line(args, ctx, host, index, sp, timed)
        if args.sweep:
            line["scan_vs_index_size"] = scan_sweep(args, ctx, db, qb, hbm, algo_bytes, scan_ms)
            index.attach(ctx)
        if args.extras:
            try:
                line["widened_rows"] = widened_block(args, ctx, fd, host, db, store, index, timed)
            except Exception as e:  # secondary measurements never cost the headline line
                line["widened_rows"] = {"error": "%s: %s" % (type(e).__name__, e)}
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or 16 * cores
        qps, secs = orc.batch_qps(0, sample, args.top, cores)
        line["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                "sample": "%d distinct queries (the first of the batch), best of 3 (%.2f s each), C++ "
                                          "restatement of the reference algorithm (oracle/, -O3), query-parallel over %d "
                                          "threads" % (sample, secs, cores)}
    print(json.dumps(line), flush=True)


def serving_loop(host, hash_params, inputs, finalize, search, timed, warmup, steps, want_rows):
    """e2e as a serving loop: `warmup` untimed and `steps` timed steps, each = take the query maps of this batch (built
    during the previous step), start the next batch's maps on the worker thread, finalize + search this batch, wait for
    the next batch's maps.  -> (per-step seconds from `timed`, last Results)"""
    maps = host.QueryMapWorker(hash_params)
    try:
        maps.start(inputs)  # pipeline fill, outside the clock (the loop then builds exactly one batch per step)

        def step():
            qb_k = maps.take()
            maps.start(inputs)
            finalize(qb_k)
            r = search(qb_k)
            maps.wait()
            return r

        res, ts = None, []
        for _ in range(warmup):
            res = step()
        for _ in range(steps):
            res, dt, _ = timed(step)
            ts.append(dt)
        rows = (int(res.struct_offsets[-1]), int(res.match_offsets[-1]))
        if rows != want_rows:
            raise RuntimeError("serving loop returned %r rows, one batch at a time %r" % (rows, want_rows))
        maps.take()
        return ts, res
    finally:
        maps.close()


def e2e_block(args, world, serial_s, pipe_s, pipe_err, h2d, d2h):
    """the e2e object of the line: throughput of the serving loop (query maps of the next batch overlapped with the search
    of the current one) when it was measured, with the one-batch-at-a-time step beside it; else that step alone"""
    step = ("host CompactStructures + query strings (held as C arrays of handles / strings, host.QueryInputs) -> "
            "make_query_map x batch -> posting counts / idf -> verification tables to the device -> count_query "
            "-> verification -> result rows in host memory")
    one = {"value": args.batch * world / serial_s, "unit": "queries/s", "ms_per_step": serial_s * 1e3}
    blk = {"value": one["value"], "unit": "queries/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
           "ms_per_step": one["ms_per_step"], "step": step}
    loop_mode = ("serving loop (host.QueryMapWorker / host.search_batches): make_query_map of batch k+1 on a "
                 "second host thread while batch k is finalized and searched; each timed step contains one "
                 "make_query_map x batch (waited for before the clock stops), one finalize and one search, "
                 "overlapped; one_batch_at_a_time is the same work with nothing overlapped")
    if pipe_s and pipe_s < serial_s:
        blk.update(value=args.batch * world / pipe_s, ms_per_step=pipe_s * 1e3, one_batch_at_a_time=one, mode=loop_mode)
    elif pipe_s:  # the overlap did not pay on this host: the line keeps the one-batch figure and shows the loop's beside it
        blk.update(mode="one batch at a time (nothing overlapped); the serving loop was measured too and was not faster",
                   serving_loop={"value": args.batch * world / pipe_s, "unit": "queries/s", "ms_per_step": pipe_s * 1e3})
    else:
        blk["mode"] = "one batch at a time (nothing overlapped)"
        if pipe_err:
            blk["serving_loop_error"] = pipe_err
    return blk


def shipped_line(args, ctx, host, index, sp, timed):
    """round 1's headline batch (the five shipped motifs x 205), device-timed, for continuity"""
    qb = make_query_batch(ctx, index, None, args.batch, 0)
    for _ in range(2):
        host.search(ctx, qb, sp)
    s0 = ctx.stage_ms("scan")
    ts = []
    n = max(2, args.steps // 2)
    for _ in range(n):
        _, dt, _ = timed(lambda: host.search(ctx, qb, sp))
        ts.append(dt)
    del qb
    return {"queries_per_s": args.batch / (sum(ts) / n), "ms_per_step": 1e3 * sum(ts) / n,
            "scan_ms": (ctx.stage_ms("scan") - s0) / n, "posting_bytes_per_step": int(ctx.last_posting_bytes),
            "workload": "the five shipped motifs cycled to %d queries (BENCH_r01's batch)" % args.batch}


def widened_block(args, ctx, fd, host, db, store, index, timed):
    """the rows of SURVEY 8f built this round, measured on the bench's own data (secondary; N = 1 only):
    K1 with every encoding of `--type` on the first structures of the database, and the per-step cost of the similarity
    metrics (K7) and of the LMS-QCP partial fit on the timed query batch"""
    out = {}
    ro = db["row_offsets"].astype(np.int64)
    m = min(3000, len(ro) - 1)
    R = int(ro[m])
    batch = fd.StructBatch(db["row_offsets"][:m + 1].astype(np.uint64), db["n_xyz"][:R], db["ca_xyz"][:R], db["cb_xyz"][:R],
                           db["aa"][:R])
    n = np.diff(ro[:m + 1])
    pair_tests = int((n * (n - 1)).sum())
    enc = {}
    for name, t, mb in (("PDBTrRosetta (default, tuned route)", 0, ()), ("PDBMotif", 1, ()), ("PDBMotifSinCos", 2, ()),
                        ("TrRosetta", 3, ()), ("PointPairFeature", 5, ()), ("TertiaryInteraction", 6, ()), ("Hybrid", 7, ()),
                        ("FolddiscoAngle", 8, ()), ("FolddiscoDist", 9, ()),
                        ("PDBTrRosetta --multiple-bins 16-4,8-3", 0, ((16, 4), (8, 3)))):
        hp = fd.HashParams(0, 0, 20.0, t, multiple_bins=mb)
        ctx.build_index(batch, hp)  # warm
        h0 = ctx.stage_ms("hash")
        ix = ctx.build_index(batch, hp)
        k1 = ctx.stage_ms("hash") - h0
        enc[name] = {"k1_hash_ms": k1, "pair_tests_per_s": pair_tests / (k1 * 1e-3) if k1 > 0 else None,
                     "hashes": int(ix.count), "posting_bytes": int(ix.value_bytes)}
    out["k1_encodings"] = {"structures": m, "pair_tests": pair_tests, "per_encoding": enc}
    # similarity metrics and partial fit on a 128-query slice of the timed batch (one warm-up, one timed step each; the
    # general path keeps every candidate's edge list on the host, so the slice bounds its memory)
    nq = min(128, args.batch)
    qb = make_query_batch(ctx, index, db, nq, 0)
    out["queries"] = nq
    base = host.SearchParams(top_n=args.top)
    for key, sp in (("default", base), ("want_metrics", host.SearchParams(top_n=args.top, want_metrics=True)),
                    ("partial_fit", host.SearchParams(top_n=args.top, partial_fit=True))):
        if key == "default":
            os.environ["FD_DEVICE_ROWS"] = "0"  # the same host row assembly as the two variants, for a like-for-like delta
        host.search(ctx, qb, sp, labels=store)
        m0, k0 = ctx.stage_ms("metrics"), ctx.stage_ms("kabsch")
        res, dt, _ = timed(lambda: host.search(ctx, qb, sp, labels=store))
        out[key] = {"ms_per_step": 1e3 * dt, "match_rows": int(len(res.matches)),
                    "k7_metrics_device_ms": ctx.stage_ms("metrics") - m0, "k5_superpose_device_ms": ctx.stage_ms("kabsch") - k0}
        os.environ.pop("FD_DEVICE_ROWS", None)
        del res
    del qb
    out["note"] = ("default = a 128-query slice of the timed batch with host row assembly (FD_DEVICE_ROWS=0); want_metrics adds one "
                   "fd_metrics_store_batch call (K7, thread per match); partial_fit verifies through the general path "
                   "(K4 + host graph step) and superposes with LMS-QCP (k5_lmsqcp_store)")
    return out


def index_build_block(ctx, host, store, db, orc, build_s, hash_ms, post_ms, hbm):
    """path (i): the rank's index build as it ran at start-up (cold: first launches of the process) and, on one GPU, a
    second, warm build timed stage by stage, beside the CPU oracle's build of a bounded sample"""
    ro = db["row_offsets"].astype(np.int64)
    out = {"first_build_s": build_s, "first_build_k1_hash_ms": hash_ms, "first_build_k2_postings_ms": post_ms,
           "note": "the first build includes CUDA module load and cold allocator pools; k2_postings_ms includes the copy "
                   "of the finished index to host memory (the drop-in files are written from it)"}
    if store is None:
        return out
    S = len(ro) - 1
    n = np.diff(ro)
    pair_tests = int((n * (n - 1)).sum())
    h0, p0 = ctx.stage_ms("hash"), ctx.stage_ms("postings")
    t0 = time.perf_counter()
    ix = host.FolddiscoIndex.build(ctx, store)
    warm_s = time.perf_counter() - t0
    b = ix.buffers()
    k1 = ctx.stage_ms("hash") - h0
    k2 = ctx.stage_ms("postings") - p0
    postings = int(np.count_nonzero(b.values < 128))
    # CPU: the oracle's build of the first 2000 structures on all cores (two passes over the pairs, like the reference)
    import oracle_lib as O
    cores = os.cpu_count() or 1
    m = min(2000, S)
    t0 = time.perf_counter()
    O.Index.build([orc.comp(s) for s in range(m)], threads=cores)
    cpu_s = time.perf_counter() - t0
    cpu_pairs = int((n[:m] * (n[:m] - 1)).sum())
    out.update({"structures": S, "residues": int(ro[-1]), "pair_tests": pair_tests, "postings": postings,
                "posting_bytes": int(len(b.values)), "warm_build_s": warm_s, "k1_hash_ms": k1, "k2_postings_ms": k2,
                "pair_tests_per_s": pair_tests / (k1 * 1e-3) if k1 > 0 else None,
                "postings_per_s": postings / (k1 * 1e-3) if k1 > 0 else None,
                "k2_GBps": (8.0 * postings + len(b.values)) / (k2 * 1e-3) / 1e9 if k2 > 0 else None,
                "k2_frac_of_hbm": (8.0 * postings + len(b.values)) / (k2 * 1e-3) / 1e9 / hbm if k2 > 0 else None,
                "cpu_build": {"structures": m, "seconds": cpu_s, "cores": cores, "pair_tests_per_s": cpu_pairs / cpu_s,
                              "kind": "port"},
                "ratio_pair_tests_per_s_whole_build": (pair_tests / warm_s) / (cpu_pairs / cpu_s)})
    return out


def scan_sweep(args, ctx, db, qb, hbm, base_bytes, base_ms):
    """Posting-list GB/s of the scan vs index size: the database tiled k times (ids shifted), index rebuilt on the GPU,
    count_query only (skip_match), same batch of distinct motifs.  Larger indexes have longer lists."""
    from folddisco_b200 import host
    out = [{"structures": args.structs_per_gpu, "algorithmic_bytes_per_launch": base_bytes, "scan_ms": base_ms,
            "GBps": base_bytes / (base_ms * 1e-3) / 1e9, "frac": base_bytes / (base_ms * 1e-3) / 1e9 / hbm}]
    skip = host.SearchParams(top_n=args.top, skip_match=True)
    for k in [int(x) for x in args.sweep.split(",") if x]:
        S = len(db["row_offsets"]) - 1
        store = host.Store()
        store.add_soa(tile_db(db, k))
        t0 = time.perf_counter()
        ix = host.FolddiscoIndex.build(ctx, store)
        ix.attach(ctx)
        build_s = time.perf_counter() - t0
        qb.finalize(ctx)  # idf weights of the larger database
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
                    "frac": gbps / hbm, "index_build_s": build_s})
        del ix, store
    return out


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
