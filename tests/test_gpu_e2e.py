"""End-to-end GPU tests through the host mirror of the reference interface (folddisco_b200.host):
index build -> files -> load -> attach -> query batch -> per-structure / per-match rows, against the README
goldens and against the oracle's full pipeline."""
import os

import numpy as np
import pytest

import fixtures as F
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import folddisco_b200 as fd
    from folddisco_b200 import host
    ctx = fd.Context(0)
    atoms = F.config1_atoms()
    names = F.serine_names()
    store = host.Store()
    for n in names:
        store.add(host.CompactStructure.from_atoms(atoms[n]), n)
    yield dict(ctx=ctx, host=host, fd=fd, atoms=atoms, names=names, store=store)
    ctx.close()


@pytest.mark.parametrize("verify_mode", [0, 1])
def test_config1_readme_rows(env, tmp_path, verify_mode):
    """verify_mode 0 = fused verification kernel, 1 = general path (K4 + host graph step + K5).
    configs[0]: index data/serine_peptidases + query 4CHA.pdb B57,B102,C195 -> README.md:218-224, 237-241"""
    host, ctx, store, names = env["host"], env["ctx"], env["store"], env["names"]
    ix = host.FolddiscoIndex.build(ctx, store)
    prefix = str(tmp_path / "serine_folddisco")
    ix.save(store, prefix, foldcomp_db="data/serine_peptidases")
    assert os.path.getsize(prefix) == F.CONFIG1_VALUE_BYTES
    assert os.path.getsize(prefix + ".offset") == F.CONFIG1_OFFSET_FILE_BYTES
    loaded = host.load_folddisco_index(prefix)
    loaded.attach(ctx)
    store.attach(ctx)
    qb = host.QueryBatch(loaded.params)
    qb.add(host.CompactStructure.from_atoms(env["atoms"]["query/4CHA.pdb"]), "B57,B102,C195")
    qb.finalize(ctx)
    assert len(qb.query_map(0)["hash"]) == F.CONFIG1_NUM_QUERY_HASHES
    res = host.search(ctx, qb, host.SearchParams(verify_mode=verify_mode), labels=store)
    srows = {os.path.basename(names[int(r["nid"])]): ("%.4f" % r["idf"], int(r["total_match_count"]),
             int(r["node_count"]), int(r["edge_count"])) for r in res.structures(0)}
    assert srows == {t: ("%.4f" % v[0], v[1], v[2], v[3]) for t, v in F.README_STRUCT_ROWS.items() if t != "1azw.pdb"} | \
        {"1azw.pdb": ("0.1856", 2, 2, 2)}
    mrows = [(os.path.basename(names[int(m["nid"])]), int(m["node_count"]), "%.4f" % m["idf"], "%.4f" % m["rmsd"],
              res.residue_string(m, 3)) for m in res.sorted_matches(0)]
    want = [(t, n, "%.4f" % i, "%.4f" % r, s) for t, n, i, r, s in F.README_MATCH_ROWS_DEFAULT]
    assert sorted(mrows) == sorted(want)
    # default sort: idf desc, rmsd asc (sort.rs:218-222)
    keys = [(-float(m["idf"]), float(m["rmsd"])) for m in res.sorted_matches(0)]
    assert keys == sorted(keys)
    # max_node_cov / min_rmsd columns of README.md:237-241
    per = {os.path.basename(names[int(r["nid"])]): (int(r["max_matching_node_count"]), "%.4f" % r["min_rmsd_with_max_match"])
           for r in res.structures(0)}
    assert per["4cha.pdb"] == (3, "0.0000") and per["1pq5.pdb"] == (3, "0.2609") and per["1l7a.pdb"] == (2, "0.7883")
    # the 1azw row of the README needs --ca-distance 1.5 (golden 3)
    res15 = host.search(ctx, qb, host.SearchParams(ca_dist_cutoff=1.5, verify_mode=verify_mode), labels=store)
    t, n, i, r, s = F.README_MATCH_ROW_1AZW_CA15
    rows15 = [(os.path.basename(names[int(m["nid"])]), int(m["node_count"]), "%.4f" % m["idf"], "%.4f" % m["rmsd"],
               res15.residue_string(m, 3)) for m in res15.sorted_matches(0)]
    assert (t, n, "%.4f" % i, "%.4f" % r, s) in rows15


def _oracle_rows(qm, comps, hits, ca_cutoff=1.0, which=0):
    rows = []
    for nid in hits["nid"]:
        r = O.retrieve(qm, comps[int(nid)], ca_cutoff=ca_cutoff, which=which)
        for m in range(len(r["rmsd"])):
            rows.append((int(nid), int(r["some"][m].sum()), float(r["idf"][m]), float(r["rmsd"][m]),
                         O.residues_to_string(r["some"][m], r["chain"][m], r["serial"][m])))
    return rows


@pytest.mark.parametrize("n_structs,seed,top_n,verify_mode", [(600, 5, None, 0), (600, 5, None, 1), (5000, 6, 40, 0),
                                                            (5000, 6, 40, 1)])
def test_synthetic_pipeline_vs_oracle(env, n_structs, seed, top_n, verify_mode):
    """all five shipped motifs against a synthetic database: matches (residues, node_count) bit-exact,
    idf / RMSD within 1e-4, vs the oracle's count_query + retrieval_wrapper."""
    from folddisco_b200 import synth
    host, ctx, fd = env["host"], env["ctx"], env["fd"]
    b = synth.generate(n_structs, seed, mean_len=150.0, max_len=500)
    store = host.Store()
    store.add_soa(b)
    ix = host.FolddiscoIndex.build(ctx, store)
    ix.attach(ctx)
    store.attach(ctx)
    bufs = ix.buffers()
    oix = O.Index.from_buffers(bufs.hashes, bufs.offsets, bufs.values)
    comps = [O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"],
                                serial=np.arange(1, len(p["aa"]) + 1, dtype=np.uint64)) for p in synth.split(b)]
    nres, plddt = ix.lookup()
    qb = host.QueryBatch(ix.params)
    oqms = []
    for path, q, _ in F.MOTIFS:
        a = env["atoms"][path]
        qb.add(host.CompactStructure.from_atoms(a), q)
        s = O.Structure.from_atoms(a)
        ch, se, subs = O.parse_query_string(q, s.first_chain)
        oqms.append(O.QueryMap(s.compact(), ch, se, subs, index=oix, total_structures=n_structs))
    qb.finalize(ctx)
    for k, om in enumerate(oqms):  # per-edge idf of the query map (query.rs:288)
        assert np.allclose(qb.query_map(k)["idf"], om.entries()["idf"], rtol=1e-5, atol=1e-6)
    sp = host.SearchParams(verify_mode=verify_mode) if top_n is None else host.SearchParams(top_n=top_n, verify_mode=verify_mode)
    res = host.search(ctx, qb, sp, labels=store)
    total_matches = 0
    for k, om in enumerate(oqms):
        op = O.CountParams.defaults(len(om.indices()), top_n=top_n if top_n is not None else O.UINT64_MAX)
        ohits = O.count_query(om, oix, nres.astype(np.uint64), plddt, op)
        srows = res.structures(k)
        got_ids = set(int(x) for x in srows["nid"])
        want_ids = set(int(x) for x in ohits["nid"])
        if top_n is None:
            assert got_ids == want_ids
        else:
            assert len(got_ids) == len(want_ids)
        common = got_ids & want_ids
        want_rows = [r for r in _oracle_rows(om, comps, ohits) if r[0] in common]
        got_rows = [(int(m["nid"]), int(m["node_count"]), float(m["idf"]), float(m["rmsd"]),
                     res.residue_string(m, len(om.indices()))) for m in res.sorted_matches(k) if int(m["nid"]) in common]
        assert sorted((r[0], r[1], r[4]) for r in got_rows) == sorted((r[0], r[1], r[4]) for r in want_rows), k
        # two components of one candidate (an SCC and the weak component around it) can map to the same residues
        # with the same RMSD but different edge sets, hence different idf: idf is part of the pairing key
        gw = sorted(got_rows, key=lambda r: (r[0], r[4], round(r[3], 3), round(r[2], 2)))
        ww = sorted(want_rows, key=lambda r: (r[0], r[4], round(r[3], 3), round(r[2], 2)))
        for g, w in zip(gw, ww):
            assert abs(g[2] - w[2]) <= 1e-4 * max(1.0, abs(w[2])), (g, w)   # match idf
            assert abs(g[3] - w[3]) <= 1e-4 * max(1.0, abs(w[3])), (g, w)   # RMSD
        total_matches += len(got_rows)
    assert total_matches > 20


def test_skip_match_and_filters(env):
    host, ctx, store = env["host"], env["ctx"], env["store"]
    ix = host.FolddiscoIndex.build(ctx, store)
    ix.attach(ctx)
    store.attach(ctx)
    qb = host.QueryBatch(ix.params)
    qb.add(host.CompactStructure.from_atoms(env["atoms"]["query/4CHA.pdb"]), "B57,B102,C195")
    qb.finalize(ctx)
    r = host.search(ctx, qb, host.SearchParams(skip_match=True), labels=store)
    assert len(r.structures(0)) == 5 and len(r.sorted_matches(0)) == 0
    sp = host.SearchParams()
    sp.max_matching_node_count = 3   # --max-node 3 (filter.rs:103-116)
    sp.rmsd_cutoff = 0.3             # --rmsd 0.3 (filter.rs:216-218)
    r = host.search(ctx, qb, sp, labels=store)
    assert sorted(os.path.basename(env["names"][int(x)]) for x in r.structures(0)["nid"]) == ["1pq5.pdb", "4cha.pdb"]
    assert all(m["rmsd"] <= 0.3 for m in r.sorted_matches(0)) and len(r.sorted_matches(0)) == 3


def test_full_size_properties(env):
    """BASELINE configs[2] at its full size (23 400 synthetic structures, the five motifs): size-independent
    properties instead of an oracle run (which would take minutes on the CPU).
      * index invariants: hashes strictly ascending, offsets strictly increasing, lists decode to strictly
        increasing ids < N, posting-count table == decoded lengths, sum of counts == number of varint terminators;
      * checksum of checksums: with no filter and no top-N, the match counts of all hits of a query add up to the
        posting counts of its hashes (every posting is exactly one vote);
      * edge_count / node_count bounds, hits ordered by (idf desc, nid asc);
      * idempotence: the same batch twice gives identical rows; top-N is a prefix of the full ranking;
      * a sample of the verified matches re-superposed on the CPU (numpy Kabsch) reproduces the RMSD within 1e-4."""
    from folddisco_b200 import synth
    host, ctx, fd = env["host"], env["ctx"], env["fd"]
    S = 23400
    db = synth.generate(S, synth.SEED_BASE + 2)
    store = host.Store()
    store.add_soa(db)
    ix = host.FolddiscoIndex.build(ctx, store)
    b = ix.buffers()
    assert np.all(np.diff(b.hashes.astype(np.int64)) > 0)
    assert b.offsets[0] == 0 and np.all(np.diff(b.offsets.astype(np.int64)) > 0) and b.offsets[-1] == len(b.values)
    ix.attach(ctx)
    store.attach(ctx)
    n_term = int(np.count_nonzero(b.values < 128))
    rng = np.random.default_rng(11)
    lens = np.diff(b.offsets.astype(np.int64))
    pick = np.concatenate([np.argsort(lens)[-20:], rng.integers(0, len(b.hashes), 300)])
    counts = ctx.posting_counts(b.hashes[pick])
    for k, li in enumerate(pick[:60]):
        ids = ctx.get_entries(int(b.hashes[li]))
        assert len(ids) == counts[k] and np.all(np.diff(ids.astype(np.int64)) > 0) and ids[-1] < S
    # all counts in chunks (the count table must account for every varint of the value file)
    total = 0
    for c0 in range(0, len(b.hashes), 1 << 22):
        total += int(ctx.posting_counts(b.hashes[c0:c0 + (1 << 22)]).astype(np.int64).sum())
    assert total == n_term
    qb = host.QueryBatch(ix.params)
    for path, q, _ in F.MOTIFS:
        qb.add(host.CompactStructure.from_atoms(env["atoms"][path]), q)
    qb.finalize(ctx)
    full = host.search(ctx, qb, host.SearchParams(skip_match=True))
    for k in range(len(F.MOTIFS)):
        rows = full.structures(k)
        qm = qb.query_map(k)
        assert int(rows["total_match_count"].astype(np.int64).sum()) == int(ctx.posting_counts(qm["hash"]).astype(np.int64).sum())
        n_edges = len(set(zip(qm["qi"].tolist(), qm["qj"].tolist())))
        n_nodes = len(set(qm["qi"].tolist()))
        assert rows["edge_count"].max() <= n_edges and rows["node_count"].max() <= n_nodes
        assert np.all(rows["edge_count"] <= rows["total_match_count"]) and np.all(rows["node_count"] <= rows["edge_count"])
        assert np.all(rows["idf"][:-1] >= rows["idf"][1:])
        assert len(np.unique(rows["nid"])) == len(rows) and rows["nid"].max() < S
    sp = host.SearchParams(top_n=100)
    r1 = host.search(ctx, qb, sp)
    r2 = host.search(ctx, qb, sp)
    def same(a, b):  # field by field: the padding bytes of the row structs are not defined
        return all(np.array_equal(a[f], b[f]) for f in a.dtype.names if not f.startswith("_"))
    assert same(r1.structs, r2.structs) and same(r1.matches, r2.matches) and same(r1.residues, r2.residues)
    soa = {k: env["host"].CompactStructure.from_atoms(env["atoms"][p]).soa() for k, (p, _, _) in enumerate(F.MOTIFS)}
    checked = 0
    for k in range(len(F.MOTIFS)):
        top = r1.structures(k)
        assert len(top) == 100
        # top-N is the head of the full ranking (up to ties at the cut)
        cut = float(full.structures(k)["idf"][99])
        assert np.all(top["idf"] >= cut - 1e-6)
        assert set(top["nid"][top["idf"] > cut + 1e-6].tolist()) <= set(full.structures(k)["nid"][:100].tolist())
        idx = qb.indices(k)
        for m in r1.sorted_matches(k)[:25]:
            res = r1.residues[int(m["res_begin"]):int(m["res_begin"]) + len(idx)]
            sel = [j for j in range(len(idx)) if res[j]["some"]]
            if len(sel) < 2:
                continue
            base = int(db["row_offsets"][int(m["nid"])])
            tr = [base + int(res[j]["serial"]) for j in sel]
            x = np.stack([np.stack([db["ca_xyz"][r], db["cb_xyz"][r]]) for r in tr]).reshape(-1, 3).astype(np.float64)
            y = np.stack([np.stack([soa[k]["ca_xyz"][idx[j]], soa[k]["cb_xyz"][idx[j]]]) for j in sel]).reshape(-1, 3).astype(np.float64)
            U, t = m["U"].reshape(3, 3).astype(np.float64), m["t"].astype(np.float64)
            rmsd = np.sqrt((((x @ U.T + t) - y) ** 2).sum(axis=1).mean())
            assert abs(rmsd - float(m["rmsd"])) <= 1e-4 * max(1.0, rmsd), (k, rmsd, float(m["rmsd"]))
            # and no rigid transform does better than the reported one (Kabsch optimality, numpy SVD)
            xc, yc = x - x.mean(0), y - y.mean(0)
            Uo, _, Vt = np.linalg.svd(xc.T @ yc)
            d = np.sign(np.linalg.det(Uo @ Vt))
            R = (Uo @ np.diag([1, 1, d]) @ Vt).T
            best = np.sqrt((((xc @ R.T) - yc) ** 2).sum(axis=1).mean())
            assert rmsd <= best + 1e-3
            checked += 1
    assert checked > 50


def test_search_stream_equals_one_batch(env):
    """host.search_stream (sub-batches, the next one prepared on a second host thread on a lane of the context) returns
    the rows of one host.search over the whole batch, query by query"""
    from folddisco_b200 import synth
    host, fd = env["host"], env["fd"]
    ctx = fd.Context(0)
    db = synth.generate(1500, 23, mean_len=180.0, max_len=500)
    store = host.Store()
    store.add_soa(db)
    ix = host.FolddiscoIndex.build(ctx, store)
    ix.attach(ctx)
    store.attach(ctx)
    motifs = [(host.CompactStructure.from_atoms(env["atoms"][p]), q) for p, q, _ in F.MOTIFS]
    structs = [motifs[k % 5][0] for k in range(37)]
    strings = [motifs[k % 5][1] for k in range(37)]
    sp = host.SearchParams(top_n=20)
    qb = host.QueryBatch(ix.params)
    qb.add_many(structs, strings)
    qb.finalize(ctx)
    whole = host.search(ctx, qb, sp, labels=store)
    parts = host.search_stream(ctx, structs, strings, sp, ix.params, sub_batch=8, labels=store)
    assert [len(p.struct_offsets) - 1 for p in parts] == [8, 8, 8, 8, 5]
    q = 0
    for p in parts:
        for k in range(len(p.struct_offsets) - 1):
            a, b = whole.structures(q), p.structures(k)
            for f in ("nid", "total_match_count", "node_count", "edge_count", "idf", "max_matching_node_count",
                      "min_rmsd_with_max_match"):
                assert np.array_equal(a[f], b[f]), (q, f)
            ma, mb = whole.sorted_matches(q), p.sorted_matches(k)
            assert len(ma) == len(mb)
            for f in ("nid", "node_count", "idf", "rmsd"):
                assert np.array_equal(ma[f], mb[f]), (q, f)
            nres = len(strings[q].split(","))
            assert [whole.residue_string(m, nres) for m in ma] == [p.residue_string(m, nres) for m in mb]
            q += 1
    assert q == 37
    del whole, parts, qb
    ctx.close()


def _soa_comps(db):
    from folddisco_b200 import synth
    return [O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"],
                               serial=np.arange(1, len(p["aa"]) + 1, dtype=np.uint64)) for p in synth.split(db)]


def test_full_size_rows_vs_oracle(env):
    """BASELINE configs[2] at its FULL size -- the bench's own database (23 400 synthetic structures, --top 100) --
    compared row for row with the oracle: the five shipped motifs and the first distinct motifs of the bench batch.
    Structure rows (ids, match / node / edge counts exact, idf 1e-4) and match rows (residues exact, idf / RMSD 1e-4)."""
    import bench
    import parity
    from folddisco_b200 import synth
    host, ctx = env["host"], env["ctx"]
    S = 23400
    db = synth.generate(S, synth.SEED_BASE + 2)
    store = host.Store()
    store.add_soa(db)
    ix = host.FolddiscoIndex.build(ctx, store)
    ix.attach(ctx)
    store.attach(ctx)
    b = ix.buffers()
    oix = O.Index.from_buffers(b.hashes, b.offsets, b.values)
    comps = _soa_comps(db)
    nres, plddt = ix.lookup()
    qb = host.QueryBatch(ix.params)
    oqms = []
    for path, q, _ in F.MOTIFS:
        a = env["atoms"][path]
        qb.add(host.CompactStructure.from_atoms(a), q)
        s = O.Structure.from_atoms(a)
        ch, se, subs = O.parse_query_string(q, s.first_chain)
        oqms.append(O.QueryMap(s.compact(), ch, se, subs, index=oix, total_structures=S))
    ro = db["row_offsets"].astype(np.int64)
    for s, pick, qstr in bench.distinct_motifs(db, 11, 0):
        qb.add(host.CompactStructure.from_soa(db["n_xyz"][ro[s]:ro[s + 1]], db["ca_xyz"][ro[s]:ro[s + 1]],
                                              db["cb_xyz"][ro[s]:ro[s + 1]], db["aa"][ro[s]:ro[s + 1]]), qstr)
        ch, se, subs = O.parse_query_string(qstr, ord("A"))
        oqms.append(O.QueryMap(comps[s], ch, se, subs, index=oix, total_structures=S))
    qb.finalize(ctx)
    res = host.search(ctx, qb, host.SearchParams(top_n=100), labels=store)
    bad, n_match = [], 0
    for k, om in enumerate(oqms):
        hits, rows = parity.oracle_query(om, oix, comps, nres, plddt, top_n=100)
        bad += parity.diff_query(res, k, len(om.indices()), hits, rows, top_n=100)
        n_match += len(rows)
    assert not bad, bad[:10]
    assert n_match > 1000


def test_missing_query_residue_counts_in_the_ratio(env):
    """A -q residue that the query structure does not have still counts in residue_count (query_pdb.rs:355-359), the
    denominator of --covered-node-ratio: with B57,B102,C195,C999 a two-node hit has ratio 2/4, not 2/3."""
    host, ctx, store, names = env["host"], env["ctx"], env["store"], env["names"]
    ix = host.FolddiscoIndex.build(ctx, store)
    ix.attach(ctx)
    store.attach(ctx)
    b = ix.buffers()
    oix = O.Index.from_buffers(b.hashes, b.offsets, b.values)
    nres, plddt = ix.lookup()
    q = "B57,B102,C195,C999"
    a = env["atoms"]["query/4CHA.pdb"]
    qb = host.QueryBatch(ix.params)
    qb.add(host.CompactStructure.from_atoms(a), q)
    qb.finalize(ctx)
    s = O.Structure.from_atoms(a)
    ch, se, subs = O.parse_query_string(q, s.first_chain)
    om = O.QueryMap(s.compact(), ch, se, subs, index=oix, total_structures=len(names))
    assert om.residue_count == 4 and len(om.indices()) == 3
    sp = host.SearchParams(covered_node_ratio=0.6)
    res = host.search(ctx, qb, sp, labels=store)
    op = O.CountParams.defaults(om.residue_count)
    op.covered_node_ratio = 0.6
    hits = O.count_query(om, oix, nres.astype(np.uint64), plddt, op)
    assert sorted(int(x) for x in res.structures(0)["nid"]) == sorted(int(x) for x in hits["nid"])
    got = {os.path.basename(names[int(x)]) for x in res.structures(0)["nid"]}
    assert got == {"4cha.pdb", "1pq5.pdb"}  # the three two-node structures pass 2/3 >= 0.6 but not 2/4


def test_long_structure_16769_residues(env):
    """data/long/6FF7.pdb uncropped (16 769 residues: the largest shipped input, 2.8e8 ordered pairs): K1 hashes ==
    the oracle's and the committed digest; the index of a database that contains it is byte-identical to the oracle's;
    a motif taken from it is found and verified in it with the oracle's rows (K3 + K6 on a 16 k-residue candidate)."""
    import hashlib
    import parity
    from folddisco_b200 import synth
    host, ctx, fd = env["host"], env["ctx"], env["fd"]
    z = np.load(os.path.join(F.GOLDEN, "long_6FF7.npz"))
    n = len(z["aa"])
    assert n == 16769
    big = dict(n_xyz=z["n_xyz"], ca_xyz=z["ca_xyz"], cb_xyz=z["cb_xyz"], aa=z["aa"], cb_valid=z["cb_valid"])
    batch = fd.StructBatch.from_list([big])
    hashes, ro = ctx.hash_structures(batch)
    assert len(hashes) == int(z["n_unique"])
    assert hashlib.sha256(hashes.astype("<u4").tobytes()).hexdigest() == str(z["sha256"])
    ocomp = O.Compact.from_soa(z["n_xyz"], z["ca_xyz"], z["cb_xyz"], z["aa"], cb_valid=z["cb_valid"], chain=z["chain"],
                               serial=z["serial"])
    assert np.array_equal(hashes, ocomp.hashes(sorted_unique=True))
    # a small database around it
    db = synth.generate(60, 99, mean_len=150.0, max_len=400)
    parts = synth.split(db)
    store = host.Store()
    comps = []

    def add_synth(p, name):
        store.add(host.CompactStructure.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"]), name)
        comps.append(O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"],
                                        serial=np.arange(1, len(p["aa"]) + 1, dtype=np.uint64)))

    for k, p in enumerate(parts[:30]):
        add_synth(p, "s%d" % k)
    store.add(host.CompactStructure.from_soa(z["n_xyz"], z["ca_xyz"], z["cb_xyz"], z["aa"], cb_valid=z["cb_valid"],
                                             chain=z["chain"], serial=z["serial"]), "6FF7")
    comps.append(ocomp)
    for k, p in enumerate(parts[30:]):
        add_synth(p, "t%d" % k)
    ix = host.FolddiscoIndex.build(ctx, store)
    want = O.Index.build(comps, threads=8)
    b = ix.buffers()
    assert np.array_equal(b.hashes, want.hashes) and np.array_equal(b.offsets, want.offsets)
    assert np.array_equal(b.values, want.values)
    ix.attach(ctx)
    store.attach(ctx)
    nres, plddt = ix.lookup()
    assert int(nres[30]) == n
    # a motif from the middle of the long structure: residues near residue 9000 with a usable amino acid
    ca = z["ca_xyz"]
    near = np.flatnonzero((np.linalg.norm(ca - ca[9000], axis=1) <= 11.0) & (z["aa"] < 20) & (z["cb_valid"] == 1))[:5]
    qstr = ",".join("%s%d" % (chr(int(z["chain"][i])), int(z["serial"][i])) for i in near)
    qb = host.QueryBatch(ix.params)
    qb.add(host.CompactStructure.from_soa(z["n_xyz"], z["ca_xyz"], z["cb_xyz"], z["aa"], cb_valid=z["cb_valid"],
                                          chain=z["chain"], serial=z["serial"]), qstr)
    qb.finalize(ctx)
    ch, se, subs = O.parse_query_string(qstr, int(z["chain"][0]))
    om = O.QueryMap(ocomp, ch, se, subs, index=want, total_structures=len(comps))
    res = host.search(ctx, qb, host.SearchParams(), labels=store)
    hits, rows = parity.oracle_query(om, want, comps, nres, plddt)
    bad = parity.diff_query(res, 0, len(om.indices()), hits, rows)
    assert not bad, bad[:10]
    assert 30 in set(int(x) for x in res.structures(0)["nid"])
    assert any(r[0] == 30 and r[1] == len(near) for r in rows)  # the motif itself, all residues matched


def test_pair_table_verification_equals_rehash(env):
    """fd_store_build_pair_table: verification by hash lookup in the store's pair table returns exactly the rows of
    the re-hash path (k6a_table vs k6a_edges), for motif-sized queries (amino-acid prefilter) and for queries of more
    than 200 hashes (every pair; retrieve.rs:24, 569), with and without a table that matches the hash parameters."""
    import bench
    from folddisco_b200 import synth
    host, ctx = env["host"], env["ctx"]
    db = synth.generate(2500, 77, mean_len=180.0, max_len=600)
    store = host.Store()
    store.add_soa(db)
    ix = host.FolddiscoIndex.build(ctx, store)
    ix.attach(ctx)
    qb = host.QueryBatch(ix.params)
    for path, q, _ in F.MOTIFS:
        qb.add(host.CompactStructure.from_atoms(env["atoms"][path]), q)
    inputs = bench.query_inputs(db, 40, 3)
    qb.add_many_indexed(*inputs)
    qb.finalize(ctx)
    assert max(len(qb.query_map(k)["hash"]) for k in range(len(qb))) > 200
    sp = host.SearchParams(top_n=60)

    def rows(res):
        out = []
        for k in range(len(qb)):
            s = res.structures(k)
            m = res.sorted_matches(k)
            n = len(qb.indices(k))
            srow = [tuple(x[f].item() for f in ("nid", "total_match_count", "node_count", "edge_count", "idf",
                                                "max_matching_node_count", "min_rmsd_with_max_match")) for x in s]
            out.append((srow, [(int(x["nid"]), int(x["node_count"]), float(x["idf"]), float(x["rmsd"]),
                                res.residue_string(x, n)) for x in m]))
        return out

    store.attach(ctx)
    base = rows(host.search(ctx, qb, sp, labels=store))
    nbytes = store.attach(ctx, pair_table=True, hash_params=ix.params)
    assert nbytes > 8 * 50 * int(db["row_offsets"][-1])
    t0 = ctx.stage_launches("pair_table")
    got = rows(host.search(ctx, qb, sp, labels=store))
    assert got == base
    assert sum(len(m) for _, m in got) > 500
    os.environ["FD_VERIFY_TABLE"] = "0"
    try:
        assert rows(host.search(ctx, qb, sp, labels=store)) == base
    finally:
        del os.environ["FD_VERIFY_TABLE"]
    # a table built with other hash parameters is ignored, a budget that is too small keeps the re-hash path
    assert store.attach(ctx, pair_table=True, hash_params=env["fd"].HashParams(8, 3, 20.0)) > 0
    assert rows(host.search(ctx, qb, sp, labels=store)) == base
    assert store.attach(ctx, pair_table=True, hash_params=ix.params, max_table_bytes=1000) is None
    assert rows(host.search(ctx, qb, sp, labels=store)) == base


def test_whole_structure_query_skip_match(env):
    """an empty query string (query.rs:226-233: every residue) through the host API with skip_match: the query map,
    the per-structure rows and their order equal the oracle's; mixed in one batch with a motif query; and, with
    matching, the verification of the whole-chain query through the general path."""
    from folddisco_b200 import synth
    host, ctx, fd = env["host"], env["ctx"], env["fd"]
    n_structs = 900
    b = synth.generate(n_structs, 9, mean_len=150.0, max_len=500)
    store = host.Store()
    store.add_soa(b)
    ix = host.FolddiscoIndex.build(ctx, store)
    ix.attach(ctx)
    store.attach(ctx)
    bufs = ix.buffers()
    oix = O.Index.from_buffers(bufs.hashes, bufs.offsets, bufs.values)
    nres, plddt = ix.lookup()
    a = env["atoms"]["query/1G2F.pdb"]
    qb = host.QueryBatch(ix.params)
    qb.add(host.CompactStructure.from_atoms(a), "")
    qb.add(host.CompactStructure.from_atoms(env["atoms"]["query/4CHA.pdb"]), "B57,B102,C195")
    qb.finalize(ctx)
    s = O.Structure.from_atoms(a)
    om = O.QueryMap(s.compact(), *O.parse_query_string("", s.first_chain), index=oix, total_structures=n_structs)
    gm, wm = qb.query_map(0), om.entries()
    assert len(gm["hash"]) == len(wm["hash"]) > 4095
    assert np.array_equal(gm["hash"], wm["hash"]) and np.array_equal(gm["qi"], wm["qi"]) and np.array_equal(gm["qj"], wm["qj"])
    res = host.search(ctx, qb, host.SearchParams(top_n=50, skip_match=True), labels=store)
    op = O.CountParams.defaults(len(om.indices()), top_n=50)
    want = O.count_query(om, oix, nres.astype(np.uint64), plddt, op)
    rows = res.structures(0)
    assert len(rows) == len(want["nid"]) == 50
    w = {int(n): (int(m), int(nc), int(ec), float(i)) for n, m, nc, ec, i in
         zip(want["nid"], want["match_count"], want["node_count"], want["edge_count"], want["idf"])}
    common = [r for r in rows if int(r["nid"]) in w]
    assert len(common) >= 45  # the cut may fall inside a group of near-equal idf
    for r in common:
        g = (int(r["total_match_count"]), int(r["node_count"]), int(r["edge_count"]))
        assert g == w[int(r["nid"])][:3]
        assert abs(float(r["idf"]) - w[int(r["nid"])][3]) <= 1e-4 * max(1.0, w[int(r["nid"])][3])
    assert len(res.structures(1)) > 0
    # with matching: the whole-chain query's candidates go through the general verification path (more than 16 query
    # residues, thousands of observed pairs); their matches equal the oracle's retrieval, structure by structure
    res5 = host.search(ctx, qb, host.SearchParams(top_n=5), labels=store)
    parts = synth.split(b)
    nq = len(om.indices())
    n_matches = 0
    for r in res5.structures(0):
        nid = int(r["nid"])
        p = parts[nid]
        comp = O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"],
                                  serial=np.arange(1, len(p["aa"]) + 1, dtype=np.uint64))
        m = O.retrieve(om, comp)
        want_rows = sorted((int(m["some"][k].sum()), O.residues_to_string(m["some"][k], m["chain"][k], m["serial"][k]))
                           for k in range(len(m["rmsd"])))
        got = [x for x in res5.sorted_matches(0) if int(x["nid"]) == nid]
        got_rows = sorted((int(x["node_count"]), res5.residue_string(x, nq)) for x in got)
        assert got_rows == want_rows, nid
        for x in got:
            key = res5.residue_string(x, nq)
            cands = [k for k in range(len(m["rmsd"])) if O.residues_to_string(m["some"][k], m["chain"][k], m["serial"][k]) == key]
            k = min(cands, key=lambda k: abs(float(m["rmsd"][k]) - float(x["rmsd"])))
            assert abs(float(m["rmsd"][k]) - float(x["rmsd"])) <= 1e-4 * max(1.0, float(m["rmsd"][k]))
            assert abs(float(m["idf"][k]) - float(x["idf"])) <= 1e-4 * max(1.0, float(m["idf"][k]))
        n_matches += len(got)
    assert n_matches > 0


def _rows(res, nq):
    """rows of every query as plain tuples (without the padding bytes of the row structs and the positions of a
    structure's matches in the batch-wide match array)"""
    out = []
    for q in range(nq):
        st = res.structures(q)
        out.append(([tuple(r[k].item() for k in st.dtype.names if k not in ("_pad", "match_begin", "match_end")) +
                     (int(r["match_end"] - r["match_begin"]),) for r in st],
                    [(int(m["nid"]), int(m["node_count"]), float(m["idf"]), float(m["rmsd"])) for m in res.sorted_matches(q)]))
    return out


def _same_rows(a, b):
    assert len(a) == len(b)
    for q, ((sa, ma), (sb, mb)) in enumerate(zip(a, b)):
        assert len(sa) == len(sb) and len(ma) == len(mb), (q, len(sa), len(sb), len(ma), len(mb))
        for k, (x, y) in enumerate(zip(sa, sb)):
            assert x == y, ("structure row", q, k, x, y)
        for k, (x, y) in enumerate(zip(ma, mb)):
            assert x == y, ("match row", q, k, x, y)


def test_repeated_batch_and_device_rows(env):
    """(1) A query batch searched again carries its batch id: count_query reuses the uploaded batch and its lookup
    results -- the rows must not change, and they must follow a changed filter, a changed batch and a newly attached
    index.  (2) Rows assembled on the device (fd_verify_rows, the default) equal the host assembly
    (FD_DEVICE_ROWS=0) row for row: summaries, both sort orders, residue labels."""
    from folddisco_b200 import synth
    host, ctx = env["host"], env["ctx"]
    b = synth.generate(800, 11, mean_len=150.0, max_len=500)
    store = host.Store()
    store.add_soa(b)
    ix = host.FolddiscoIndex.build(ctx, store)
    ix.attach(ctx)
    store.attach(ctx)
    qb = host.QueryBatch(ix.params)
    for path, q, _ in F.MOTIFS:
        qb.add(host.CompactStructure.from_atoms(env["atoms"][path]), q)
    qb.finalize(ctx)
    sp = host.SearchParams(top_n=60)
    first = host.search(ctx, qb, sp, labels=store)
    again = host.search(ctx, qb, sp, labels=store)  # cache hit
    nq = len(F.MOTIFS)
    _same_rows(_rows(first, nq), _rows(again, nq))
    for q in range(nq):  # residue labels of every match row
        for m in first.sorted_matches(q)[:20]:
            assert first.residue_string(m, len(qb.indices(q))) == again.residue_string(m, len(qb.indices(q)))
    os.environ["FD_DEVICE_ROWS"] = "0"
    try:
        on_host = host.search(ctx, qb, sp, labels=store)
    finally:
        del os.environ["FD_DEVICE_ROWS"]
    _same_rows(_rows(first, nq), _rows(on_host, nq))
    for q in range(nq):
        ms, mh = first.sorted_matches(q), on_host.sorted_matches(q)
        assert len(ms) == len(mh)
        for x, y in zip(ms, mh):
            assert first.residue_string(x, len(qb.indices(q))) == on_host.residue_string(y, len(qb.indices(q)))
            assert np.array_equal(x["U"], y["U"]) and np.array_equal(x["t"], y["t"])
    assert sum(len(first.sorted_matches(q)) for q in range(nq)) > 50
    # a changed filter is a different cache key; a changed batch has a new id
    sp2 = host.SearchParams(top_n=10)
    few = host.search(ctx, qb, sp2, labels=store)
    assert all(len(few.structures(q)) <= 10 for q in range(nq)) and any(len(first.structures(q)) > 10 for q in range(nq))
    qb.add(host.CompactStructure.from_atoms(env["atoms"]["query/4CHA.pdb"]), "B57,B102,C195")
    qb.finalize(ctx)
    more = host.search(ctx, qb, sp, labels=store)
    _same_rows(_rows(more, nq), _rows(first, nq))
    _same_rows(_rows(more, nq + 1)[nq:], _rows(first, 1))
    # a newly attached index invalidates the cached lookup results
    b2 = synth.generate(300, 12, mean_len=150.0, max_len=500)
    store2 = host.Store()
    store2.add_soa(b2)
    ix2 = host.FolddiscoIndex.build(ctx, store2)
    ix2.attach(ctx)
    store2.attach(ctx)
    qb.finalize(ctx)
    other = host.search(ctx, qb, sp, labels=store2)
    assert all(int(r["nid"]) < 300 for q in range(nq) for r in other.structures(q))
