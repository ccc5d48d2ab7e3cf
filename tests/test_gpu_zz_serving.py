"""GPU test of the serving loop (host.search_batches / host.QueryMapWorker): the query maps of the next batch are built
on a second host thread while the current batch is finalized and searched.  Same rows as one batch at a time.
(Named to run after the other GPU tests: it was written after the round's GPU minutes were spent.)"""
import numpy as np
import pytest

import fixtures as F

pytestmark = pytest.mark.gpu


def test_search_batches_equals_one_batch_at_a_time():
    import folddisco_b200 as fd
    from folddisco_b200 import host, synth
    ctx = fd.Context(0)
    db = synth.generate(1500, 23, mean_len=180.0, max_len=500)
    store = host.Store()
    store.add_soa(db)
    ix = host.FolddiscoIndex.build(ctx, store)
    ix.attach(ctx)
    store.attach(ctx)
    atoms = F.config1_atoms()
    motifs = [(host.CompactStructure.from_atoms(atoms[p]), q) for p, q, _ in F.MOTIFS]
    structs = [motifs[k % 5][0] for k in range(37)]
    strings = [motifs[k % 5][1] for k in range(37)]
    sp = host.SearchParams(top_n=20)
    bounds = [(0, 16), (16, 32), (32, 37)]
    batches = [host.QueryInputs(structs[a:b], strings[a:b]) for a, b in bounds]
    # one batch at a time
    serial = []
    for a, b in bounds:
        qb = host.QueryBatch(ix.params)
        qb.add_many(structs[a:b], strings[a:b])
        qb.finalize(ctx)
        serial.append(host.search(ctx, qb, sp, labels=store))
    # the serving loop, twice (the worker thread and its batches are released in between)
    for _ in range(2):
        looped = list(host.search_batches(ctx, batches, sp, ix.params, labels=store))
        assert [len(p.struct_offsets) - 1 for p in looped] == [16, 16, 5]
        for (a, b), want, got in zip(bounds, serial, looped):
            for k in range(b - a):
                x, y = want.structures(k), got.structures(k)
                for f in ("nid", "total_match_count", "node_count", "edge_count", "idf", "max_matching_node_count",
                          "min_rmsd_with_max_match"):
                    assert np.array_equal(x[f], y[f]), (a + k, f)
                mx, my = want.sorted_matches(k), got.sorted_matches(k)
                assert len(mx) == len(my)
                for f in ("nid", "node_count", "idf", "rmsd"):
                    assert np.array_equal(mx[f], my[f]), (a + k, f)
                nres = len(strings[a + k].split(","))
                assert [want.residue_string(m, nres) for m in mx] == [got.residue_string(m, nres) for m in my]
        del looped
    del serial, qb, batches
    ctx.close()
