"""GPU parity tests: the CUDA path, called through the C ABI (include/folddisco_b200.h), against the CPU oracle.

Bit-exact for hashes, index bytes, ids, counts; 1e-4 for idf and RMSD (tolerances written at each assert).
Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import os

import numpy as np
import pytest

import fixtures as F
import oracle_lib as O

pytestmark = pytest.mark.gpu

IDF_RTOL = 1e-4  # north_star: float within 1e-4
RMSD_ATOL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    import folddisco_b200 as fd
    c = fd.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def config1():
    atoms = F.config1_atoms()
    names = F.serine_names()
    comps = [O.Structure.from_atoms(atoms[n]).compact() for n in names]
    soas = [c.soa() for c in comps]
    return dict(names=names, comps=comps, soas=soas, atoms=atoms)


def soa_to_struct(d):
    """oracle SoA -> product convention: aa = 128 + code for non-canonical residue names"""
    aa = d["aa"].copy()
    canon = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE", "PRO",
             "SER", "THR", "TRP", "TYR", "VAL"]
    for i in range(len(aa)):
        if aa[i] != 255 and bytes(d["res_name"][i]).decode() != canon[aa[i]]:
            aa[i] = 128 + aa[i]
    return dict(n_xyz=d["n_xyz"], ca_xyz=d["ca_xyz"], cb_xyz=d["cb_xyz"], aa=aa, cb_valid=d["cb_valid"])


def make_batch(soas):
    import folddisco_b200 as fd
    return fd.StructBatch.from_list([soa_to_struct(d) for d in soas])


def synth_compacts(n, seed, **kw):
    from folddisco_b200 import synth
    kw.setdefault("mean_len", 120.0)
    kw.setdefault("max_len", 400)
    b = synth.generate(n, seed, **kw)
    parts = synth.split(b)
    comps = [O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"]) for p in parts]
    return b, parts, comps


# ------------------------------------------------------------------------------------------------
def test_device_math_bit_exact(ctx):
    """fd_math.cuh on the device == its host build == the oracle's exact math, bit for bit."""
    from folddisco_b200 import capi
    rng = np.random.default_rng(11)
    n = 200000
    a = np.concatenate([rng.uniform(-4, 4, n), [0.0, -0.0, np.pi, -np.pi, np.nan, np.inf, 1e-30, 3.1415927, 1.5707964]]).astype(np.float32)
    u = np.concatenate([rng.uniform(-1, 1, n), [1.0, -1.0, 1.0000001, -1.0000001, 0.0, np.nan, 0.5, -0.5, 0.99999994]]).astype(np.float32)
    y = np.concatenate([rng.normal(0, 1, n), [0.0, -0.0, 0.0, 1.0, -1.0, np.nan, 1e-20, 0.0, 5.0]]).astype(np.float32)
    x = np.concatenate([rng.normal(0, 1, n), [1.0, -1.0, 0.0, 0.0, 0.0, 1.0, -1e20, -0.0, np.inf]]).astype(np.float32)
    L = O.lib()
    for op, (p, q) in enumerate([(a, None), (a, None), (u, None), (y, x)]):
        dev = ctx.math_probe(op, p, q)
        host = capi.math_host(op, p, q)
        assert np.array_equal(dev.view(np.uint32), host.view(np.uint32)) or np.array_equal(
            dev[~np.isnan(dev)].view(np.uint32), host[~np.isnan(host)].view(np.uint32))
        assert np.array_equal(np.isnan(dev), np.isnan(host))
        # oracle spot check on a subsample
        fn = [L.fdo_math_sinf, L.fdo_math_cosf, L.fdo_math_acosf, None][op]
        for k in list(range(0, n, 997)) + list(range(n, len(p))):
            want = np.float32(fn(float(p[k]))) if fn else np.float32(L.fdo_math_atan2f(float(p[k]), float(q[k])))
            assert (np.isnan(want) and np.isnan(dev[k])) or want == dev[k], (op, p[k], dev[k], want)


def test_hash_structures_config1(ctx, config1):
    """K1 == get_geometric_hash_as_u32_from_structure + sort + dedup on every config-1 structure (bit-exact)."""
    batch = make_batch(config1["soas"])
    hashes, ro = ctx.hash_structures(batch)
    assert len(ro) == len(config1["comps"]) + 1
    for s, c in enumerate(config1["comps"]):
        want = c.hashes(sorted_unique=True)
        got = hashes[int(ro[s]):int(ro[s + 1])]
        assert np.array_equal(got, want), config1["names"][s]


def test_hash_structures_edge_cases(ctx, config1):
    import folddisco_b200 as fd
    d = {k: v[:60].copy() for k, v in config1["soas"][3].items()}
    d["aa"][5] = 255        # unknown residue: no features (feature.rs:21-23)
    d["cb_valid"][9] = 0    # residue without CB / virtual CB (core.rs:153-155)
    empty = {k: v[:0] for k, v in d.items()}
    one = {k: v[:1] for k, v in d.items()}
    parts = [d, empty, one, {k: v[:33] for k, v in config1["soas"][0].items()}]
    batch = make_batch(parts)
    hashes, ro = ctx.hash_structures(batch, fd.HashParams(16, 4, 20.0))
    for s, p in enumerate(parts):
        c = O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"], p["cb_valid"])
        assert np.array_equal(hashes[int(ro[s]):int(ro[s + 1])], c.hashes(16, 4, 20.0, sorted_unique=True)), s
    # non-default bins and cutoff (pdb_tr.rs:22-35 clamps)
    for nbd, nba, cut in ((8, 3, 12.0), (20, 9, 20.0), (16, 4, 5.0)):
        h2, r2 = ctx.hash_structures(make_batch([d]), fd.HashParams(nbd, nba, cut))
        c = O.Compact.from_soa(d["n_xyz"], d["ca_xyz"], d["cb_xyz"], d["aa"], d["cb_valid"])
        assert np.array_equal(h2, c.hashes(nbd, nba, cut, sorted_unique=True))
    # zero structures
    h0, r0 = ctx.hash_structures(fd.StructBatch(np.zeros(1, np.uint64), np.zeros((0, 3)), np.zeros((0, 3)),
                                                np.zeros((0, 3)), np.zeros(0, np.uint8)))
    assert len(h0) == 0 and r0.tolist() == [0]


def test_build_index_config1_byte_exact(ctx, config1):
    """K1+K2: the three index arrays are byte-identical to the oracle's (and to the reference sizes)."""
    ix = ctx.build_index(make_batch(config1["soas"]))
    want = O.Index.build(config1["comps"])
    assert ix.count == F.CONFIG1_NUM_HASHES and ix.value_bytes == F.CONFIG1_VALUE_BYTES
    assert len(ix.offset_file_bytes()) == F.CONFIG1_OFFSET_FILE_BYTES
    assert np.array_equal(ix.hashes, want.hashes)
    assert np.array_equal(ix.offsets, want.offsets)
    assert np.array_equal(ix.values, want.values)


def test_build_postings_varints(ctx):
    """fd_build_postings: ids >= 128 / >= 16384 / >= 2^21 need 2 / 3 / 4 LEB128 bytes; id 0 -> one 0x00 byte."""
    rng = np.random.default_rng(5)
    S = 3000
    rows = [np.unique(rng.integers(0, 50, rng.integers(0, 12)).astype(np.uint32)) for _ in range(S)]
    ro = np.zeros(S + 1, np.uint64)
    ro[1:] = np.cumsum([len(r) for r in rows])
    flat = np.concatenate(rows).astype(np.uint32)
    for first_id in (0, 5_000_000):
        got = ctx.build_postings(flat, ro, first_id=first_id)
        # oracle with the same ids: shift rows by first_id using empty leading rows is too slow; decode instead
        want = O.Index.from_csr(flat, ro)
        assert np.array_equal(got.hashes, want.hashes)
        dec = O.Index.from_buffers(got.hashes, got.offsets, got.values)
        for h in want.hashes[:50]:
            assert np.array_equal(dec.entries(int(h)), want.entries(int(h)) + np.uint64(first_id))
        if first_id == 0:
            assert np.array_equal(got.offsets, want.offsets) and np.array_equal(got.values, want.values)
    # empty input
    e = ctx.build_postings(np.zeros(0, np.uint32), np.zeros(1, np.uint64))
    assert e.count == 0 and e.value_bytes == 0 and e.offsets.tolist() == [0]


def test_build_index_synthetic_and_shards(ctx):
    """400 synthetic structures: fused build == oracle; two hash-range shards concatenate to the full index."""
    b, parts, comps = synth_compacts(400, 1234)
    import folddisco_b200 as fd
    batch = fd.StructBatch(b["row_offsets"], b["n_xyz"], b["ca_xyz"], b["cb_xyz"], b["aa"])
    ix = ctx.build_index(batch)
    want = O.Index.build(comps, threads=8)
    assert np.array_equal(ix.hashes, want.hashes) and np.array_equal(ix.offsets, want.offsets)
    assert np.array_equal(ix.values, want.values)
    cut = int(np.median(ix.hashes)) & ~((1 << 20) - 1)
    lo = ctx.build_index(batch, hash_lo=0, hash_hi=cut)
    hi = ctx.build_index(batch, hash_lo=cut, hash_hi=1 << 32)
    assert np.array_equal(np.concatenate([lo.hashes, hi.hashes]), ix.hashes)
    assert np.array_equal(np.concatenate([lo.values, hi.values]), ix.values)
    assert np.array_equal(np.concatenate([lo.offsets[:-1], hi.offsets + lo.offsets[-1]]), ix.offsets)


def _attach_synth(ctx, n, seed, **kw):
    b, parts, comps = synth_compacts(n, seed, **kw)
    import folddisco_b200 as fd
    batch = fd.StructBatch(b["row_offsets"], b["n_xyz"], b["ca_xyz"], b["cb_xyz"], b["aa"])
    ix = ctx.build_index(batch)
    nres = batch.nres
    rng = np.random.default_rng(seed)
    plddt = rng.uniform(30, 95, n).astype(np.float32)
    ctx.index_attach(ix, nres, plddt)
    oix = O.Index.from_buffers(ix.hashes, ix.offsets, ix.values)
    return dict(batch=batch, ix=ix, oix=oix, nres=nres.astype(np.uint64), plddt=plddt, comps=comps, parts=parts)


def test_attach_counts_and_decode(ctx):
    """directory + skip table + counts: posting counts and decoded lists == oracle get_entries (bit-exact)."""
    # three templates, tiny jitter: near-duplicate structures give posting lists that span several 256 B segments
    env = _attach_synth(ctx, 3000, 77, jitter=0.02, mutate=0.0, template_ids=[4, 5, 9])
    ix, oix = env["ix"], env["oix"]
    lens = np.diff(ix.offsets.astype(np.int64))
    longest = np.argsort(lens)[-40:]
    rng = np.random.default_rng(3)
    pick = np.concatenate([longest, rng.integers(0, ix.count, 200)])
    hs = ix.hashes[pick]
    absent = np.array([1, 2, 0x3fffffff, 0xffffffff], np.uint32)
    absent = absent[~np.isin(absent, ix.hashes)]
    counts = ctx.posting_counts(np.concatenate([hs, absent]))
    for k, h in enumerate(hs):
        want = oix.entries(int(h))
        assert counts[k] == len(want)
        if k < 60:
            assert np.array_equal(ctx.get_entries(int(h)), want)
    assert np.all(counts[len(hs):] == 0)
    assert lens.max() > 512, "test index too small to exercise multi-segment lists"


def _query_inputs(qm):
    """oracle query map -> fd_query dict (edges / nodes densified in first-appearance order)"""
    e = qm.entries()
    edges, nodes = {}, {}
    edge_of_hash = []
    for qi, qj in zip(e["qi"].tolist(), e["qj"].tolist()):
        if (qi, qj) not in edges:
            edges[(qi, qj)] = len(edges)
            nodes.setdefault(qi, len(nodes))
        edge_of_hash.append(edges[(qi, qj)])
    edge_node = [nodes[qi] for (qi, qj) in edges]
    return dict(hashes=e["hash"], edge_of_hash=np.array(edge_of_hash, np.uint16),
                edge_node=np.array(edge_node, np.uint16), n_nodes=len(nodes),
                expected_node_count=len(qm.indices()))


def _motif_qmaps(oix, n_structs, extra=()):
    atoms = F.config1_atoms()
    out = []
    for path, q, _ in list(F.MOTIFS) + list(extra):
        s = O.Structure.from_atoms(atoms[path])
        ch, se, subs = O.parse_query_string(q, s.first_chain)
        out.append(O.QueryMap(s.compact(), ch, se, subs, index=oix, total_structures=n_structs))
    return out


def _compare_hits(got, want, top_n=None):
    """got: HIT_DTYPE array (idf desc, nid asc); want: oracle dict in the same order.  Integer fields exact,
    idf within IDF_RTOL, order equal modulo idf ties / near-ties."""
    w = {int(n): (int(m), int(nc), int(ec), float(i)) for n, m, nc, ec, i in
         zip(want["nid"], want["match_count"], want["node_count"], want["edge_count"], want["idf"])}
    g = {int(r["nid"]): (int(r["match_count"]), int(r["node_count"]), int(r["edge_count"]), float(r["idf"])) for r in got}
    if top_n is None:
        assert set(g) == set(w)
    else:
        # the cut may fall inside a group of (near-)equal idf: everything strictly above the cut must agree
        assert len(g) == len(w)
        if len(w):
            cut = min(v[3] for v in w.values())
            strict = {n for n, v in w.items() if v[3] > cut * (1 + 2 * IDF_RTOL) + 1e-7}
            assert strict <= set(g)
    for n in set(g) & set(w):
        assert g[n][:3] == w[n][:3], (n, g[n], w[n])
        assert abs(g[n][3] - w[n][3]) <= IDF_RTOL * max(1.0, abs(w[n][3])), (n, g[n], w[n])
    idf = got["idf"]
    assert np.all(idf[:-1] >= idf[1:]), "hits must be ordered by idf descending"
    same = idf[:-1] == idf[1:]
    assert np.all(got["nid"][:-1][same] < got["nid"][1:][same]), "ties must be ordered by ascending nid"


def test_count_query_config1(ctx, config1):
    """K3 on the config-1 index reproduces README.md:237-241 and the oracle."""
    ix = ctx.build_index(make_batch(config1["soas"]))
    nres = np.array([c.nres for c in config1["comps"]], np.uint32)
    plddt = np.array([c.avg_plddt for c in config1["comps"]], np.float32)
    ctx.index_attach(ix, nres, plddt)
    oix = O.Index.from_buffers(ix.hashes, ix.offsets, ix.values)
    qms = _motif_qmaps(oix, len(nres))
    got = ctx.count_query_batch([_query_inputs(qm) for qm in qms])
    for qm, g in zip(qms, got):
        _compare_hits(g, O.count_query(qm, oix, nres.astype(np.uint64), plddt))
    rows = {config1["names"][int(r["nid"])].split("/")[-1]: ("%.4f" % r["idf"], int(r["match_count"]),
            int(r["node_count"]), int(r["edge_count"])) for r in got[0]}
    assert rows == {t: ("%.4f" % v[0], v[1], v[2], v[3]) for t, v in F.README_STRUCT_ROWS.items()}


@pytest.mark.parametrize("n_structs,seed,kw", [(700, 21, {}), (9000, 22, {}),
                                               (12000, 23, dict(jitter=0.05, mutate=0.02, template_ids=[4, 5, 9, 10]))])
def test_count_query_synthetic(ctx, n_structs, seed, kw):
    """single-tile (700) and multi-tile (9000, 12000 > tile capacity) vote kernels vs the oracle, with filters /
    top-N; the third case has long multi-segment posting lists that cross tile boundaries."""
    import folddisco_b200 as fd
    env = _attach_synth(ctx, n_structs, seed, **kw)
    oix, nres, plddt = env["oix"], env["nres"], env["plddt"]
    qms = _motif_qmaps(oix, n_structs, extra=[("query/4CHA.pdb", "B57:X,B102,C195:ST", None)])
    queries = [_query_inputs(qm) for qm in qms]
    got = ctx.count_query_batch(queries)
    total_bytes = 0
    for qm, g in zip(qms, got):
        _compare_hits(g, O.count_query(qm, oix, nres, plddt))
        for h in qm.entries()["hash"]:
            k = np.searchsorted(env["ix"].hashes, h)
            if k < env["ix"].count and env["ix"].hashes[k] == h:
                total_bytes += int(env["ix"].offsets[k + 1] - env["ix"].offsets[k])
    assert ctx.last_posting_bytes == total_bytes
    assert sum(len(g) for g in got) > 50
    # filters + top-N (filter.rs:76-100, query_pdb.rs:404-411)
    p = fd.PrefilterParams(top_n=25, length_penalty=0.3, total_match_count=2, covered_node_count=2,
                           covered_node_ratio=0.5, idf_score_cutoff=0.01, num_res_cutoff=250, plddt_cutoff=40.0)
    got = ctx.count_query_batch(queries, p)
    for qm, g in zip(qms, got):
        op = O.CountParams(-1.0, -1, -1.0, 0.3, 2, 2, 0.5, 0.01, 250, 40.0, len(qm.indices()), 25, 1)
        _compare_hits(g, O.count_query(qm, oix, nres, plddt, op), top_n=25)
    # the histogram pre-selection of the top-N path (normally only used when the sort would be large)
    os.environ["FD_K3_TOPSEL_RATIO"] = "0"
    try:
        got2 = ctx.count_query_batch(queries, p)
    finally:
        del os.environ["FD_K3_TOPSEL_RATIO"]
    for g, g2 in zip(got, got2):
        assert np.array_equal(g, g2)
    # freq filter and sampling (count_query.rs:124-128, 222-253)
    p = fd.PrefilterParams(freq_filter=0.02)
    got = ctx.count_query_batch(queries[:2], p)
    for qm, g in zip(qms[:2], got):
        op = O.CountParams.defaults(len(qm.indices()))
        op.freq_filter = 0.02
        _compare_hits(g, O.count_query(qm, oix, nres, plddt, op))
    # empty batch and a query whose hashes are all absent
    assert ctx.count_query_batch([]) == []
    none = dict(hashes=np.array([1, 2, 3], np.uint32), edge_of_hash=np.zeros(3, np.uint16),
                edge_node=np.zeros(1, np.uint16), n_nodes=1)
    assert len(ctx.count_query_batch([none])[0]) == 0


def test_count_query_whole_structure(ctx):
    """whole-structure queries (empty query string: every residue is a query residue, query.rs:226-233) take the
    global-memory count_query path: thousands of edges, ~10^4-10^5 hashes.  Rows == oracle for the whole chain of
    1G2F and a ~200-residue synthetic chain, alone, chunked by node group (tiny key budget), mixed with motif-sized
    queries in one batch, and with filters / top-n / sampling."""
    import folddisco_b200 as fd
    n_structs = 1500
    env = _attach_synth(ctx, n_structs, 31)
    oix, nres, plddt = env["oix"], env["nres"], env["plddt"]
    atoms = F.config1_atoms()
    s = O.Structure.from_atoms(atoms["query/1G2F.pdb"])
    big = max(range(n_structs), key=lambda i: env["comps"][i].nres if env["comps"][i].nres <= 220 else 0)
    whole = [O.QueryMap(s.compact(), *O.parse_query_string("", s.first_chain), index=oix, total_structures=n_structs),
             O.QueryMap(env["comps"][big], *O.parse_query_string("", ord("A")), index=oix, total_structures=n_structs)]
    for qm in whole:
        assert len(qm.entries()["hash"]) > 4095 and len(qm.indices()) > 100
    small = _motif_qmaps(oix, n_structs)[:2]
    queries = [_query_inputs(qm) for qm in whole]
    got = ctx.count_query_batch(queries)
    want = [O.count_query(qm, oix, nres, plddt) for qm in whole]
    for g, w in zip(got, want):
        assert len(g) > 100
        _compare_hits(g, w)
    os.environ["FD_K3W_CHUNK_KEYS"] = "20000"  # many chunks of a few node groups each
    try:
        got_chunked = ctx.count_query_batch(queries)
    finally:
        del os.environ["FD_K3W_CHUNK_KEYS"]
    for g, g2 in zip(got, got_chunked):
        assert np.array_equal(g, g2)
    # mixed batch: rows come back in the caller's order
    mixed = [_query_inputs(small[0]), queries[0], _query_inputs(small[1]), queries[1]]
    got_m = ctx.count_query_batch(mixed)
    assert np.array_equal(got_m[1], got[0]) and np.array_equal(got_m[3], got[1])
    for qm, g in zip(small, (got_m[0], got_m[2])):
        _compare_hits(g, O.count_query(qm, oix, nres, plddt))
    # filters + top-n, then sampling of the rarest hashes
    p = fd.PrefilterParams(top_n=30, length_penalty=0.3, total_match_count=3, covered_node_count=2,
                           covered_node_ratio=0.02, idf_score_cutoff=0.01, num_res_cutoff=300, plddt_cutoff=0.0)
    for qm, g in zip(whole, ctx.count_query_batch(queries, p)):
        op = O.CountParams(-1.0, -1, -1.0, 0.3, 3, 2, 0.02, 0.01, 300, 0.0, len(qm.indices()), 30, 1)
        _compare_hits(g, O.count_query(qm, oix, nres, plddt, op), top_n=30)
    p = fd.PrefilterParams(sampling_ratio=0.25)
    for qm, g in zip(whole[:1], ctx.count_query_batch(queries[:1], p)):
        op = O.CountParams.defaults(len(qm.indices()))
        op.sampling_ratio = 0.25
        _compare_hits(g, O.count_query(qm, oix, nres, plddt, op))


@pytest.mark.parametrize("env", [dict(FD_K3_CTAS="4", FD_K3_THREADS="128"), dict(FD_K3_CTAS="1", FD_K3_THREADS="512"),
                                 dict(FD_K3_CTAS="3", FD_K3_THREADS="256", FD_K3_LIMIT="0"), dict(FD_K3_V1="1")])
def test_scan_plan_variants(ctx, env):
    """k3_scan_v3 under other shared-memory plans (4 small tiles per SM: every query of the 12 000-structure case is
    split into several id tiles; one large tile; the tile-level top-n pre-selection off) returns exactly the rows of
    the default plan (so does the first-generation kernel, FD_K3_V1=1), and those equal the oracle."""
    import folddisco_b200 as fd
    e = _attach_synth(ctx, 12000, 23, jitter=0.05, mutate=0.02, template_ids=[4, 5, 9, 10])
    oix, nres, plddt = e["oix"], e["nres"], e["plddt"]
    qms = _motif_qmaps(oix, 12000, extra=[("query/4CHA.pdb", "B57:X,B102,C195:ST", None)])
    queries = [_query_inputs(qm) for qm in qms] * 7  # more work items than one wave of CTAs takes
    p_all = fd.PrefilterParams()
    p_top = fd.PrefilterParams(top_n=40)
    base_all = ctx.count_query_batch(queries, p_all)
    base_top = ctx.count_query_batch(queries, p_top)
    for qm, g in zip(qms, base_all):
        _compare_hits(g, O.count_query(qm, oix, nres, plddt))
    for k, v in env.items():
        os.environ[k] = v
    try:
        got_all = ctx.count_query_batch(queries, p_all)
        got_top = ctx.count_query_batch(queries, p_top)
    finally:
        for k in env:
            del os.environ[k]
    for a, b in zip(base_all, got_all):
        assert np.array_equal(a, b)
    for a, b in zip(base_top, got_top):
        assert np.array_equal(a, b)
        assert len(a) <= 40
    assert max(len(a) for a in base_top) == 40


def test_count_query_id_range_shards(ctx):
    """SURVEY 8e ablation / fd_count_query_batch_ex: the database cut into three id ranges, one index per range, every
    shard given the GLOBAL list lengths -> the merged rows are bit-identical to the unsharded search (same idf
    weights, same fixed-point scale, a structure's postings all live on its shard)."""
    import folddisco_b200 as fd
    n = 6000
    b, parts, comps = synth_compacts(n, 31)
    batch = fd.StructBatch(b["row_offsets"], b["n_xyz"], b["ca_xyz"], b["cb_xyz"], b["aa"])
    nres = batch.nres
    plddt = np.random.default_rng(5).uniform(30, 95, n).astype(np.float32)
    full = ctx.build_index(batch)
    ctx.index_attach(full, nres, plddt)
    oix = O.Index.from_buffers(full.hashes, full.offsets, full.values)
    qms = _motif_qmaps(oix, n)
    queries = [_query_inputs(qm) for qm in qms]
    flat = np.concatenate([q["hashes"] for q in queries])
    for p in (fd.PrefilterParams(), fd.PrefilterParams(top_n=30),
              fd.PrefilterParams(top_n=30, freq_filter=0.05, num_res_cutoff=200)):
        ctx.index_attach(full, nres, plddt)
        want = ctx.count_query_batch(queries, p)
        gcounts = ctx.posting_counts(flat)
        cuts = [0, 1500, 3100, n]
        ro = b["row_offsets"].astype(np.int64)
        merged = [[] for _ in queries]
        total = np.zeros(len(flat), np.uint64)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            sl = slice(ro[lo], ro[hi])
            sb = fd.StructBatch(b["row_offsets"][lo:hi + 1] - b["row_offsets"][lo], b["n_xyz"][sl], b["ca_xyz"][sl],
                                b["cb_xyz"][sl], b["aa"][sl])
            six = ctx.build_index(sb)
            ctx.index_attach(six, nres[lo:hi], plddt[lo:hi])
            total += ctx.posting_counts(flat)
            got = ctx.count_query_batch(queries, p, global_counts=gcounts, global_n_structs=n)
            for k, g in enumerate(got):
                g = g.copy()
                g["nid"] += lo
                merged[k].append(g)
        assert np.array_equal(total, gcounts.astype(np.uint64))  # a list is the concatenation of its shards' pieces
        for k, w in enumerate(want):
            m = np.concatenate(merged[k])
            order = np.lexsort((m["nid"], -m["idf"].astype(np.float64)))
            m = m[order][:len(w)] if p.top_n < 1 << 62 else m[order]
            assert np.array_equal(m, w), k


def test_kabsch_batch(ctx):
    """K5 vs the oracle's f64 Kabsch: rmsd / U / t within 1e-4 (north_star tolerance)."""
    rng = np.random.default_rng(9)
    movs, refs, offs = [], [], [0]
    for a in range(300):
        m = int(rng.integers(2, 12)) * 2
        ref = rng.normal(0, 8, (m, 3))
        Rm = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        if np.linalg.det(Rm) < 0:
            Rm[:, 0] *= -1
        mov = (ref - ref.mean(0)) @ Rm.T + rng.normal(0, 0.3 * (a % 5), (m, 3)) + rng.uniform(-30, 30, 3)
        movs.append(mov)
        refs.append(ref)
        offs.append(offs[-1] + m)
    # degenerate inputs: identical points, collinear points
    movs.append(np.ones((4, 3)));  refs.append(np.ones((4, 3)) * 2);  offs.append(offs[-1] + 4)
    movs.append(np.outer(np.arange(4), [1, 2, 3]));  refs.append(np.outer(np.arange(4), [3, 2, 1]));  offs.append(offs[-1] + 4)
    mov = np.concatenate(movs).astype(np.float32)
    ref = np.concatenate(refs).astype(np.float32)
    rmsd, U, t = ctx.kabsch_batch(mov, ref, np.array(offs, np.uint32))
    for a in range(len(offs) - 1):
        Uo, to, ro = O.kabsch(mov[offs[a]:offs[a + 1]], ref[offs[a]:offs[a + 1]])
        assert abs(rmsd[a] - ro) <= RMSD_ATOL * max(1.0, ro), (a, rmsd[a], ro)
        if a < 300:
            assert np.allclose(U[a], Uo, atol=1e-4) and np.allclose(t[a], to, atol=1e-3)
    # reference KAT (kabsch.rs:560-600)
    src = [[6.994, 8.354, 42.405], [9.429, 7.479, 48.266], [5.547, 0.158, 42.050]]
    t1 = [[-13.958, -1.741, -4.223], [-12.833, 3.134, -7.780], [-5.720, -2.218, -3.368]]
    r, _, _ = ctx.kabsch_batch(np.array(t1, np.float32), np.array(src, np.float32), np.array([0, 3], np.uint32))
    assert r[0] < 0.2


def _retrieval_inputs(qm):
    e = qm.entries()
    # observed_distance_map flattened in insertion order: recompute from the query structure like query.rs:271-280
    qc = qm.query
    idx = qm.indices()
    d = qc.soa()
    aa1, aa2, dist, qi = [], [], [], []
    for a in idx:
        for b in idx:
            if a == b:
                continue
            f = qc.pair_feature(int(a), int(b))
            if f is None:
                continue
            aa1.append(d["aa"][a]); aa2.append(d["aa"][b]); dist.append(f[2]); qi.append(a)
    return dict(hashes_sorted=np.sort(e["hash"]), aa1=np.array(aa1, np.uint8), aa2=np.array(aa2, np.uint8),
                ca_dist=np.array(dist, np.float32), q_index=np.array(qi, np.uint32))


def test_candidate_edges(ctx, config1):
    """K4 vs retrieve_with_prefilter of the oracle on config 1 (all five targets, three motifs incl. knottin
    which has more than 200 hashes... no: 172, so it still uses the prefilter) and on synthetic targets."""
    import folddisco_b200 as fd
    batch = make_batch(config1["soas"])
    ctx.store_attach(batch)
    ix = ctx.build_index(batch)
    oix = O.Index.from_buffers(ix.hashes, ix.offsets, ix.values)
    qms = _motif_qmaps(oix, 5)
    rqs = [_retrieval_inputs(qm) for qm in qms]
    cand_q, cand_t = [], []
    for q in range(len(qms)):
        for t in range(5):
            cand_q.append(q); cand_t.append(t)
    for ca_cut in (1.0, 1.5):
        edges, pairs = ctx.candidate_edges_batch(rqs, cand_q, cand_t, ca_dist_cutoff=ca_cut)
        n_edges = 0
        for c, (q, t) in enumerate(zip(cand_q, cand_t)):
            r = O.retrieve(qms[q], config1["comps"][t], ca_cutoff=ca_cut)
            ei, ej, eh = r["edges"]
            mine = edges[edges["cand"] == c]
            assert mine["i"].tolist() == ei.tolist() and mine["j"].tolist() == ej.tolist(), (q, t)
            assert mine["hash"].tolist() == eh.tolist(), (q, t)
            n_edges += len(ei)
        assert n_edges > 20
        assert np.all(np.diff(pairs["cand"].astype(np.int64)) >= 0)
    # all-pairs fallback: a query with > 200 hashes (retrieve.rs:569) -- widen the thresholds
    atoms = F.config1_atoms()
    s = O.Structure.from_atoms(atoms["query/2N6N.pdb"])
    ch, se, subs = O.parse_query_string("3,10,15,16,21,23,28,30", s.first_chain)
    big = O.QueryMap(s.compact(), ch, se, subs, dist_thr=(0.5, 1.0, 1.5), angle_thr=(5.0, 10.0, 15.0))
    assert len(big.entries()["hash"]) > 200
    edges, pairs = ctx.candidate_edges_batch([_retrieval_inputs(big)], [0] * 5, list(range(5)))
    for t in range(5):
        ei, ej, eh = O.retrieve(big, config1["comps"][t])["edges"]
        mine = edges[edges["cand"] == t]
        assert mine["i"].tolist() == ei.tolist() and mine["j"].tolist() == ej.tolist() and mine["hash"].tolist() == eh.tolist()
