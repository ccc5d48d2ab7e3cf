#!/usr/bin/env python
"""Index-build sweep for path (i) (BASELINE configs[4]: "index-build kernel + query throughput sweep"): K1 pair hashing +
K2 postings for S synthetic structures on one B200, through the same host call the bench uses
(fdh_index_build -> fd_build_index, chunked by hash range above 600 M keys).

    python tools/build_sweep.py 23400 100000 > gpurun_out/build_sweep.json

Per size: residues, ordered pair tests sum n(n-1), postings (unique (hash, structure) pairs), device ms of the hash and
postings stages, pair tests / s, postings / s, index bytes, wall seconds of the whole build (host round trip included).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    import folddisco_b200 as fd
    from folddisco_b200 import host, synth
    sizes = [int(x) for x in sys.argv[1:]] or [23400, 100000]
    ctx = fd.Context(0)
    warm = host.Store()
    warm.add_soa(synth.generate(500, synth.SEED_BASE + 9))
    host.FolddiscoIndex.build(ctx, warm)  # module load, pool growth
    out = []
    for S in sizes:
        db = synth.generate(S, synth.SEED_BASE + 5)
        store = host.Store()
        store.add_soa(db)
        h0, p0 = ctx.stage_ms("hash"), ctx.stage_ms("postings")
        t0 = time.perf_counter()
        ix = host.FolddiscoIndex.build(ctx, store)
        wall = time.perf_counter() - t0
        hash_ms, post_ms = ctx.stage_ms("hash") - h0, ctx.stage_ms("postings") - p0
        n = np.diff(np.asarray(db["row_offsets"], np.uint64).astype(np.float64))
        pairs = float((n * (n - 1)).sum())
        import ctypes as C
        v = host._IndexBuffers()
        host._lib().fdh_index_get(ix.h, C.byref(v))  # a view: no copy of the index
        count, value_bytes = int(v.count), int(v.value_bytes)
        rec = {"structures": S, "residues": int(n.sum()), "pair_tests": pairs, "distinct_hashes": count,
               "posting_bytes": value_bytes, "hash_stage_ms": hash_ms, "postings_stage_ms": post_ms,
               "pair_tests_per_s": pairs / (hash_ms * 1e-3) if hash_ms > 0 else None,
               "posting_bytes_per_s_encode_stage": value_bytes / (post_ms * 1e-3) if post_ms > 0 else None,
               "build_wall_s": wall}
        out.append(rec)
        sys.stderr.write(json.dumps(rec) + "\n")
        del ix, store, db
    print(json.dumps({"index_build_sweep": out}))


if __name__ == "__main__":
    main()
