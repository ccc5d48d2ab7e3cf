"""Drives the kernels of the widened rows once each, for an ncu capture:
K1 typed (TrRosetta, Hybrid, multiple bins), K4 typed + K5 (a PDBMotifSinCos search), K7 (metrics) and k5_lmsqcp_store
(partial fit) on a 3 000-structure synthetic database and 128 distinct motif queries.

    ncu --set full --clock-control none --import-source on -k regex:'k7_metrics|k5_lmsqcp' -c 2 \
        -o gpurun_out/rNN_widened -f python tools/profile_widened.py --only metrics,partial
(restrict the kernel set: a full capture replays every launch ~39 times, and the K1 / K4 launches of this script are
tens of milliseconds each)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import folddisco_b200 as fd  # noqa: E402
from folddisco_b200 import host, synth  # noqa: E402

ctx = fd.Context(0)
db = synth.generate(3000, synth.SEED_BASE + 2)
store = host.Store()
store.add_soa(db)
ro = db["row_offsets"]
batch = fd.StructBatch(ro, db["n_xyz"], db["ca_xyz"], db["cb_xyz"], db["aa"])
only = set(sys.argv[sys.argv.index("--only") + 1].split(",")) if "--only" in sys.argv else {"k1", "typed", "metrics", "partial"}
if "k1" in only:
    for t, mb in ((3, ()), (7, ()), (0, ((16, 4), (8, 3)))):
        ctx.build_index(batch, fd.HashParams(0, 0, 20.0, t, multiple_bins=mb))
for what, hp, sp in (("typed", fd.HashParams(0, 0, 20.0, 2), host.SearchParams(top_n=100)),
                     ("metrics", fd.HashParams(), host.SearchParams(top_n=100, want_metrics=True)),
                     ("partial", fd.HashParams(), host.SearchParams(top_n=100, partial_fit=True))):
    if what not in only:
        continue
    ix = host.FolddiscoIndex.build(ctx, store, hp)
    ix.attach(ctx)
    store.attach(ctx)
    qb = bench.make_query_batch(ctx, ix, db, 128, 0)
    res = host.search(ctx, qb, sp, labels=store)
    print("type", hp.hash_type, "matches", len(res.matches), "metrics", None if res.metrics is None else res.metrics.shape)
ctx.close()
