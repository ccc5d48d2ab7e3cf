"""torchrun worker of tests/test_gpu_sharded.py::test_id_range_two_ranks (also run in-process with world = 1): one
rank per GPU; the library's own NCCL communicator (fd_comm_init) carries the collectives.  Each rank builds the index of
its id range, searches its own slice of the batch through fdh_search_sharded and compares every row with the unsharded
search of the same queries on the same rank (second context, full index): rows must be IDENTICAL (bit for bit)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def run(rank, world, local, dist):
    import fixtures as F
    import folddisco_b200 as fd
    from folddisco_b200 import host, sharded, synth
    from test_gpu_sharded import _slice_batches
    atoms = F.config1_atoms()
    db = synth.generate(3000, 41, mean_len=150.0, max_len=500)
    store = host.Store()
    store.add_soa(db)
    full = fd.Context(local)
    ix = host.FolddiscoIndex.build(full, store)
    ix.attach(full)
    store.attach(full)
    total_rows = 0
    ctx = fd.Context(local)
    sharded.comm_init(ctx, rank, world, dist)
    assert ctx.comm_rank == rank and ctx.comm_world == world
    sh, full_store = sharded.IdRangeShards.build(ctx, db, rank, world)
    for top_n in (50, 7):
        sp = host.SearchParams(top_n=top_n)
        qb_ref = _slice_batches(host, ix.params, atoms, world, reps=3)[rank]
        qb_ref.finalize(full)
        want = host.search(full, qb_ref, sp, labels=store)
        qb = _slice_batches(host, ix.params, atoms, world, reps=3)[rank]
        sh.prepare(ctx, qb)
        for k in range(len(qb)):  # per-edge idf from the all-reduced list lengths == the unsharded index's
            assert np.array_equal(qb.query_map(k)["idf"], qb_ref.query_map(k)["idf"])
        got = sh.search(ctx, qb, sp, labels=full_store)
        assert np.array_equal(got.struct_offsets, want.struct_offsets)
        assert np.array_equal(got.match_offsets, want.match_offsets)
        for k in range(len(qb)):
            a, b = got.structures(k), want.structures(k)
            for f in ("nid", "total_match_count", "node_count", "edge_count", "idf", "max_matching_node_count",
                      "min_rmsd_with_max_match"):
                assert np.array_equal(a[f], b[f]), (k, f)
            ma, mb = got.sorted_matches(k), want.sorted_matches(k)
            for f in ("nid", "node_count", "idf", "rmsd"):
                assert np.array_equal(ma[f], mb[f]), (k, f)
            n = len(qb.indices(k))
            assert [got.residue_string(m, n) for m in ma] == [want.residue_string(m, n) for m in mb]
            total_rows += len(ma)
        if world > 1:
            assert ctx.last_exchange_bytes > 0
    assert total_rows > 10
    ctx.close()
    full.close()
    print("id-range ok rank %d rows %d" % (rank, total_rows), flush=True)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    run(rank, world, local, dist)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
