#!/usr/bin/env python
"""index build (path (i)) repeated in one process, with searches in between: stage times of every build"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import folddisco_b200 as fd
from folddisco_b200 import host, synth
ctx = fd.Context(0)
db = synth.generate(int(sys.argv[1]) if len(sys.argv) > 1 else 23400, synth.SEED_BASE + 2)
store = host.Store(); store.add_soa(db)
names = ("hash", "hash_alloc", "postings")
def snap(): return {k: ctx.stage_ms(k) for k in names}
def build(tag):
    a = snap(); t0 = time.perf_counter()
    ix = host.FolddiscoIndex.build(ctx, store)
    dt = time.perf_counter() - t0; b = snap()
    print(tag, "wall %.3f s" % dt, {k: round(b[k] - a[k], 1) for k in names}, flush=True)
    return ix
ix = build("build 1"); ix = build("build 2")
ix.attach(ctx); store.attach(ctx)
qb = bench.make_query_batch(ctx, ix, None, 1024, 0)
for _ in range(3): host.search(ctx, qb, host.SearchParams(top_n=100))
ix = build("build 3 (after searches)"); ix = build("build 4")
