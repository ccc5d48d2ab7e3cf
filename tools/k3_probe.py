#!/usr/bin/env python
"""count_query (K3) alone on the bench database: stage times per step for a batch of motif queries.

    python tools/k3_probe.py [--structs 23400] [--batch 1024] [--top 100] [--steps 5] [--tile K] [--distinct]

--tile K: the database tiled K times (ids shifted) -- longer posting lists, the same per-hash frequencies.
Used under ncu for the k3_scan captures in profiles/.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--structs", type=int, default=23400)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--top", type=int, default=100)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--tile", type=int, default=1)
    ap.add_argument("--distinct", action="store_true")
    ap.add_argument("--variants", default="", help="comma list of extra runs, each ENV=VALUE[+ENV=VALUE] (e.g. FD_K3_V1=1)")
    args = ap.parse_args()
    import bench
    import folddisco_b200 as fd
    from folddisco_b200 import host, synth
    ctx = fd.Context(0)
    db = synth.generate(args.structs, synth.SEED_BASE + 2)
    if args.tile > 1:
        db = bench.tile_db(db, args.tile)
    store = host.Store()
    store.add_soa(db)
    t0 = time.perf_counter()
    index = host.FolddiscoIndex.build(ctx, store)
    index.attach(ctx)
    build_s = time.perf_counter() - t0
    qb = bench.make_query_batch(ctx, index, db if args.distinct else None, args.batch, 0)
    sp = host.SearchParams(top_n=args.top, skip_match=True)

    def measure(tag):
        for _ in range(3):
            host.search(ctx, qb, sp)
        st = ("lookup", "scan", "select")
        s0 = {k: ctx.stage_ms(k) for k in st}
        t0 = time.perf_counter()
        nbytes = 0
        for _ in range(args.steps):
            r = host.search(ctx, qb, sp)
            nbytes += ctx.last_posting_bytes
        wall = (time.perf_counter() - t0) / args.steps
        out = {k: (ctx.stage_ms(k) - s0[k]) / args.steps for k in st}
        out.update(variant=tag, structures=len(store), batch=args.batch, build_s=build_s, wall_ms=wall * 1e3,
                   posting_bytes_per_step=nbytes / args.steps, rows=int(r.struct_offsets[-1]),
                   scan_GBps=nbytes / args.steps / (out["scan"] * 1e-3) / 1e9 if out["scan"] > 0 else None)
        print(json.dumps(out), flush=True)

    measure("default")
    for v in [x for x in args.variants.split(",") if x]:
        kv = dict(p.split("=") for p in v.split("+"))
        os.environ.update(kv)
        try:
            measure(v)
        finally:
            for k in kv:
                del os.environ[k]


if __name__ == "__main__":
    main()
