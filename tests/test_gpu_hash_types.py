"""GPU parity of the other encodings (`--type`) and `--multiple-bins` through the C ABI: K1 typed
(fd_hash_structures / fd_build_index), count_query on the resulting index, and the general verification path with the
typed K4 (fd_candidate_edges_batch) against the oracle run in the same encoding
(reference: src/controller/feature.rs:11-231, src/geometry/*.rs, src/controller/query.rs:53-206,
src/controller/retrieve.rs:52-156, 563-602)."""
import numpy as np
import pytest

import fixtures as F
import oracle_lib as O
import parity

pytestmark = pytest.mark.gpu

# (name, FD_HASH_* value, reference HashType index, nbin_dist, nbin_angle, --multiple-bins list)
CASES = [
    ("PDBMotif", 1, 0, 0, 0, []),
    ("PDBMotifSinCos", 2, 1, 0, 0, []),
    ("TrRosetta", 3, 2, 0, 0, []),
    ("PointPairFeature", 5, 4, 12, 5, []),
    ("TertiaryInteraction", 6, 5, 0, 0, []),
    ("Hybrid", 7, 6, 0, 0, []),
    ("FolddiscoAngle", 8, 7, 0, 0, []),
    ("FolddiscoDist", 9, 8, 0, 0, []),
    ("PDBTrRosetta-multiple-bins", 0, 3, 0, 0, [(16, 4), (8, 3)]),
    ("PDBMotifSinCos-multiple-bins", 2, 1, 0, 0, [(8, 3), (16, 4), (4, 2)]),
]


@pytest.fixture(scope="module")
def env():
    import folddisco_b200 as fd
    from folddisco_b200 import host, synth
    ctx = fd.Context(0)
    b = synth.generate(300, 11, mean_len=110.0, max_len=300)
    parts = synth.split(b)
    comps = [O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"],
                                serial=np.arange(1, len(p["aa"]) + 1, dtype=np.uint64)) for p in parts]
    store = host.Store()
    store.add_soa(b)
    yield dict(ctx=ctx, fd=fd, host=host, db=b, parts=parts, comps=comps, store=store, atoms=F.config1_atoms())
    ctx.close()


@pytest.mark.parametrize("name,fd_type,ref_type,nbd,nba,mb", CASES, ids=[c[0] for c in CASES])
def test_typed_index_and_search_vs_oracle(env, name, fd_type, ref_type, nbd, nba, mb):
    fd, host, ctx, store, comps = env["fd"], env["host"], env["ctx"], env["store"], env["comps"]
    hp = fd.HashParams(nbd, nba, 20.0, fd_type, multiple_bins=mb)
    with O.hash_mode(ref_type, mb):
        # K1: per-structure sorted unique hashes of the first structures
        few = env["parts"][:12]
        batch = fd.StructBatch.from_list([dict(n_xyz=p["n_xyz"], ca_xyz=p["ca_xyz"], cb_xyz=p["cb_xyz"], aa=p["aa"],
                                               cb_valid=None) for p in few])
        hashes, ro = ctx.hash_structures(batch, hp)
        for s in range(len(few)):
            want = comps[s].hashes(nbd, nba, 20.0, sorted_unique=True)
            assert np.array_equal(hashes[int(ro[s]):int(ro[s + 1])], want), (name, s)
        # K1 + K2: the whole index, byte-identical to the oracle's builder
        ix = host.FolddiscoIndex.build(ctx, store, hp)
        bufs = ix.buffers()
        oix = O.Index.build(comps, nbd, nba)
        assert np.array_equal(bufs.hashes, oix.hashes) and np.array_equal(bufs.offsets, oix.offsets)
        assert np.array_equal(bufs.values, oix.values)
        assert ix.params.hash_type == fd_type and ix.params.n_multiple_bins == len(mb)
        ix.attach(ctx)
        store.attach(ctx)
        nres, plddt = ix.lookup()
        # query path: make_query_map -> count_query -> general verification (typed K4 + host graph step + K5)
        qb = host.QueryBatch(ix.params)
        oqms = []
        for path, q, _ in F.MOTIFS:
            a = env["atoms"][path]
            qb.add(host.CompactStructure.from_atoms(a), q)
            s = O.Structure.from_atoms(a)
            ch, se, subs = O.parse_query_string(q, s.first_chain)
            oqms.append(O.QueryMap(s.compact(), ch, se, subs, nbin_dist=nbd, nbin_angle=nba, index=oix,
                                   total_structures=len(comps)))
        qb.finalize(ctx)
        res = host.search(ctx, qb, host.SearchParams(), labels=store)
        n_rows = 0
        for k, om in enumerate(oqms):
            assert np.array_equal(qb.query_map(k)["hash"], om.entries()["hash"])
            assert np.allclose(qb.query_map(k)["idf"], om.entries()["idf"], rtol=1e-5, atol=1e-6)
            op = O.CountParams.defaults(om.residue_count)
            hits = O.count_query(om, oix, nres.astype(np.uint64), plddt, op)
            rows = []
            for nid in hits["nid"]:
                r = O.retrieve(om, comps[int(nid)], nbin_dist=nbd, nbin_angle=nba)
                for m in range(len(r["rmsd"])):
                    rows.append((int(nid), int(r["some"][m].sum()), float(r["idf"][m]), float(r["rmsd"][m]),
                                 O.residues_to_string(r["some"][m], r["chain"][m], r["serial"][m])))
            bad = parity.diff_query(res, k, len(om.indices()), hits, rows)
            assert not bad, (name, bad[:5])
            n_rows += len(rows)
        assert n_rows > 0


def test_unknown_encodings_are_refused(env):
    fd, ctx = env["fd"], env["ctx"]
    p = env["parts"][0]
    batch = fd.StructBatch.from_list([dict(n_xyz=p["n_xyz"], ca_xyz=p["ca_xyz"], cb_xyz=p["cb_xyz"], aa=p["aa"],
                                           cb_valid=None)])
    for t in (10, 1000):  # not a HashType
        with pytest.raises(fd.FdError):
            ctx.hash_structures(batch, fd.HashParams(0, 0, 20.0, t))
    with pytest.raises(fd.FdError):  # the pair table is PDBTrRosetta single-bin only
        env["store"].attach(ctx, pair_table=True, hash_params=fd.HashParams(0, 0, 20.0, 2))
