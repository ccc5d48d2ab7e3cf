"""torchrun worker of tests/test_gpu_sharded.py::test_two_rank_nccl: one rank per GPU, NCCL.  Each rank builds its
hash-range shard, runs the sharded search and compares its slice of the batch with the unsharded search computed on the
same rank (second context, full index)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as dist
    import fixtures as F
    import folddisco_b200 as fd
    from folddisco_b200 import host, sharded, synth
    from test_gpu_sharded import _motif_batch, same_rows
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    atoms = F.config1_atoms()
    db = synth.generate(3000, 41, mean_len=150.0, max_len=500)
    store = host.Store()
    store.add_soa(db)
    full = fd.Context(local)
    ix = host.FolddiscoIndex.build(full, store)
    ix.attach(full)
    store.attach(full)
    sp = host.SearchParams(top_n=50)
    qb_full = _motif_batch(host, ix.params, atoms, reps=3)
    qb_full.finalize(full)
    want = host.search(full, qb_full, sp)

    ctx = fd.Context(local)
    sh = sharded.ShardedIndex.build(ctx, store, rank, world)
    store.attach(ctx)
    qb = _motif_batch(host, ix.params, atoms, reps=3)
    sh.finalize(ctx, qb, dist)
    got = sh.search_dense(ctx, qb, sp, dist)
    q0, q1 = sharded.query_slice(len(qb), rank, world)
    total = 0
    for q in range(q0, q1):
        total += same_rows(got, want, q - q0, q)
        n = len(qb.indices(q))
        kg = sorted((int(m["nid"]), int(m["node_count"]), got.residue_string(m, n)) for m in got.sorted_matches(q - q0))
        kw = sorted((int(m["nid"]), int(m["node_count"]), want.residue_string(m, n)) for m in want.sorted_matches(q))
        assert kg == kw, q
    assert total > 10
    dense_ms = sh.merge_ms
    # sparse protocol: this rank owns a slice of the batch
    from test_gpu_sharded import _slice_batches
    qb_mine = _slice_batches(host, ix.params, atoms, world, reps=3)[rank]
    sh.merge_ms, sh.merge_bytes = 0.0, 0
    sh.prepare(ctx, qb_mine, dist)
    assert int(sh.slice_begin[rank]) == q0 and int(sh.slice_begin[rank + 1]) == q1
    got2 = sh.search(ctx, qb_mine, sp, dist)
    total2 = 0
    for q in range(q0, q1):
        total2 += same_rows(got2, want, q - q0, q)
        n = len(qb.indices(q))
        kg = sorted((int(m["nid"]), int(m["node_count"]), got2.residue_string(m, n)) for m in got2.sorted_matches(q - q0))
        kw = sorted((int(m["nid"]), int(m["node_count"]), want.residue_string(m, n)) for m in want.sorted_matches(q))
        assert kg == kw, q
    assert total2 == total
    print("sparse ok rank %d: exchange %.3f ms for %.2f MB (dense all_reduce: %.3f ms)" %
          (rank, sh.merge_ms, sh.merge_bytes / 1e6, dense_ms), flush=True)
    dist.barrier()
    print("sharded ok rank %d: %d queries, %d matches, merge %.3f ms for %.1f MB" %
          (rank, q1 - q0, total, sh.merge_ms, sh.merge_bytes / 1e6), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
