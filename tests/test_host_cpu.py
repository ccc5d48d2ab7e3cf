"""CPU tests of the host side (folddisco_b200/csrc/host): parsing, CompactStructure quirks, query-map construction
and index files against the oracle.  No GPU needed: nothing here launches a kernel."""
import glob
import os

import numpy as np
import pytest

import fixtures as F
import oracle_lib as O
from conftest import REF, needs_reference


@pytest.fixture(scope="module")
def host():
    from folddisco_b200 import build
    build.build()
    from folddisco_b200 import host
    return host


def _same_compact(h, o):
    d, e = h.soa(), o.soa()
    assert h.num_residues == o.nres
    for k in ("n_xyz", "ca_xyz", "cb_xyz", "cb_valid", "chain", "serial", "b_factor"):
        assert np.array_equal(d[k], e[k]), k
    code = np.where(d["aa"] == 255, 255, d["aa"] & 0x7F)
    assert np.array_equal(code, e["aa"])
    canon = [b"ALA", b"ARG", b"ASN", b"ASP", b"CYS", b"GLN", b"GLU", b"GLY", b"HIS", b"ILE", b"LEU", b"LYS", b"MET",
             b"PHE", b"PRO", b"SER", b"THR", b"TRP", b"TYR", b"VAL"]
    for i in range(len(code)):
        if code[i] != 255:
            assert (d["aa"][i] >= 128) == (bytes(e["res_name"][i]) != canon[code[i]])


def test_compact_builder_matches_oracle(host):
    for name, a in F.config1_atoms().items():
        _same_compact(host.CompactStructure.from_atoms(a), O.Structure.from_atoms(a).compact())


def test_compact_builder_quirks(host):
    """SURVEY 8a Q1-Q5 on a hand-made atom list: last atom never consumed, chain/B taken from the flush atom, stale
    backbone C reused, residues without N/CA dropped, virtual CB for any residue lacking CB, modified residues."""
    rows = [  # name, res, chain, serial, xyz, b
        (" N  ", "ALA", "A", 1, (0.0, 0.0, 0.0), 10.0), (" CA ", "ALA", "A", 1, (1.4, 0.0, 0.0), 11.0),
        (" C  ", "ALA", "A", 1, (2.0, 1.4, 0.0), 12.0), (" CB ", "ALA", "A", 1, (1.9, -0.8, 1.2), 13.0),
        (" N  ", "GLY", "A", 2, (3.3, 1.5, 0.0), 20.0), (" CA ", "GLY", "A", 2, (4.0, 2.8, 0.0), 21.0),
        (" C  ", "GLY", "A", 2, (5.5, 2.6, 0.0), 22.0),
        (" CA ", "SER", "A", 3, (6.9, 3.9, 0.2), 30.0),                       # no N: dropped (Q4)
        (" N  ", "MSE", "B", 4, (8.0, 4.0, 1.0), 40.0), (" CA ", "MSE", "B", 4, (9.2, 4.8, 1.2), 41.0),  # no C, no CB
        (" N  ", "LYS", "B", 5, (10.0, 6.0, 2.0), 50.0), (" CA ", "LYS", "B", 5, (11.2, 6.5, 2.5), 51.0),
        (" CB ", "LYS", "B", 5, (11.0, 7.5, 3.6), 52.0), (" O  ", "LYS", "B", 5, (12.0, 5.0, 3.0), 53.0),
    ]
    a = dict(x=np.array([r[4][0] for r in rows], np.float32), y=np.array([r[4][1] for r in rows], np.float32),
             z=np.array([r[4][2] for r in rows], np.float32),
             atom_name=np.array([list(r[0].encode()) for r in rows], np.uint8),
             chain=np.array([ord(r[2]) for r in rows], np.uint8),
             res_name=np.array([list(r[1].encode()) for r in rows], np.uint8),
             res_serial=np.array([r[3] for r in rows], np.uint64), b_factor=np.array([r[5] for r in rows], np.float32))
    h = host.CompactStructure.from_atoms(a)
    _same_compact(h, O.Structure.from_atoms(a).compact())
    d = h.soa()
    assert d["serial"].tolist() == [1, 2, 4, 5]
    assert d["aa"].tolist() == [0, 7, 128 + 12, 11]
    assert d["cb_valid"].tolist() == [1, 1, 1, 1]          # MSE reuses the stale C of GLY 2 (Q3, Q5)
    assert d["chain"].tolist() == [65, 65, 66, 66]         # Q2: chain of the atom that triggered the flush
    assert d["b_factor"].tolist() == [20.0, 30.0, 50.0, 53.0]
    assert h.first_chain == ord("A")


def test_parse_query_string(host):
    """src/controller/query.rs:425-465"""
    for q, dc in (("A250,A232,A269", "A"), ("A250-252,B232:H,269:NDp", "C"), ("1-3:X", "1"), ("B57,B102,C195", "B"),
                  ("164:H,195,221,247:ND,297:H", "A"), (" A1 , A2 ", "A")):
        res, sub = host.parse_query_string(q, ord(dc))
        ch, se, osub = O.parse_query_string(q, ord(dc))
        assert res == list(zip(ch.tolist(), se.tolist())) and sub == osub, q
    assert host.parse_query_string("", ord("A")) == ([], [])
    with pytest.raises(ValueError):
        host.parse_query_string("A12,", ord("A"))
    with pytest.raises(ValueError):
        host.parse_query_string("Axx", ord("A"))


@pytest.mark.parametrize("dist_thr,angle_thr", [((0.5,), (5.0,)), ((0.5, 1.0), (5.0, 10.0)), ((), ())])
def test_query_map_matches_oracle(host, dist_thr, angle_thr):
    """make_query_map: same hashes, same edge per hash, same insertion order as the oracle (idf needs the GPU)."""
    atoms = F.config1_atoms()
    qb = host.QueryBatch(dist_thr=dist_thr, angle_thr=angle_thr)
    cases = list(F.MOTIFS) + [("query/4CHA.pdb", "B57:X,B102,C195:ST", None), ("query/4CHA.pdb", "B57,B57,Z9,C195", None),
                              ("query/1G2F.pdb", "", None)]  # empty query string = every residue (query.rs:226-233)
    for path, q, _ in cases:
        qb.add(host.CompactStructure.from_atoms(atoms[path]), q)
    for k, (path, q, want_n) in enumerate(cases):
        s = O.Structure.from_atoms(atoms[path])
        ch, se, subs = O.parse_query_string(q, s.first_chain)
        om = O.QueryMap(s.compact(), ch, se, subs, dist_thr=dist_thr, angle_thr=angle_thr)
        e, g = om.entries(), qb.query_map(k)
        for f in ("hash", "qi", "qj", "primary"):
            assert np.array_equal(e[f], g[f]), (path, f)
        assert np.array_equal(om.indices(), qb.indices(k))
        if want_n is not None and dist_thr == (0.5,):
            assert len(g["hash"]) == want_n


def test_index_files_byte_identical(host, tmp_path):
    """fdh_index_save writes PREFIX / .offset / .lookup / .type byte-identical to the oracle's writers."""
    import ctypes as C
    from folddisco_b200 import capi
    atoms = F.config1_atoms()
    names = F.serine_names()
    store = host.Store()
    comps = []
    for n in names:
        store.add(host.CompactStructure.from_atoms(atoms[n]), n)
        comps.append(O.Structure.from_atoms(atoms[n]).compact())
    oix = O.Index.build(comps)
    hashes, offsets, values = oix.hashes, oix.offsets, oix.values
    b = capi._IndexBuffers(len(hashes), hashes.ctypes.data_as(C.POINTER(C.c_uint32)),
                           offsets.ctypes.data_as(C.POINTER(C.c_uint64)), len(values),
                           values.ctypes.data_as(C.POINTER(C.c_uint8)))
    p = capi.HashParams(0, 0, 20.0)
    ix = host.FolddiscoIndex(host._lib().fdh_index_from_buffers(C.byref(b), store.h, C.byref(p)))
    mine, ref = str(tmp_path / "mine"), str(tmp_path / "ref")
    ix.save(store, mine, max_residue=50000, foldcomp_db="data/serine_peptidases")
    oix.save(ref)
    nres = np.array([c.nres for c in comps], np.uint64)
    plddt = np.array([c.avg_plddt for c in comps], np.float32)
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    O.lib().fdo_lookup_save((ref + ".lookup").encode(), len(names), arr, nres, plddt)
    O.lib().fdo_type_save((ref + ".type").encode(), 0, 0, 20.0, len(names), 50000, b"data/serine_peptidases")
    for ext in ("", ".offset", ".lookup", ".type"):
        assert open(mine + ext, "rb").read() == open(ref + ext, "rb").read(), ext
    assert os.path.getsize(mine) == F.CONFIG1_VALUE_BYTES
    assert os.path.getsize(mine + ".offset") == F.CONFIG1_OFFSET_FILE_BYTES
    first = open(mine + ".lookup").readline().rstrip("\n").split("\t")
    assert first[:3] == ["0", names[0], "626"] and "%.4f" % float(first[3]) == "34.2399" and first[4] == "0"
    # load back (mmap) and compare
    back = host.load_folddisco_index(mine)
    bb = back.buffers()
    assert np.array_equal(bb.hashes, hashes) and np.array_equal(bb.offsets, offsets) and np.array_equal(bb.values, values)
    nr, pl = back.lookup()
    assert nr.tolist() == nres.tolist() and np.allclose(pl, plddt)
    assert back.name(4) == names[4] and back.params.dist_cutoff == 20.0
    del back, bb
    with open(mine, "r+b") as fh:  # a value file that is shorter than the offsets say: refused at load
        fh.truncate(F.CONFIG1_VALUE_BYTES // 2)
    with pytest.raises(Exception) as e:
        host.load_folddisco_index(mine)
    assert "inconsistent" in str(e.value)
    # hostile / corrupt offset files: a count that would wrap the size computation, offsets that run backwards
    raw = bytearray(open(mine + ".offset", "rb").read())
    bad = bytearray(raw)
    bad[0:8] = np.uint64(1 << 62).tobytes()
    open(mine + ".offset", "wb").write(bad)
    with pytest.raises(Exception) as e:
        host.load_folddisco_index(mine)
    assert "corrupted" in str(e.value)
    bad = bytearray(raw)
    count = int(np.frombuffer(raw[:8], np.uint64)[0])
    pos = 8 + 4 * count + 8 * 100  # offsets[100] := a value beyond its successor
    bad[pos:pos + 8] = np.uint64(F.CONFIG1_VALUE_BYTES).tobytes()
    open(mine + ".offset", "wb").write(bad)
    with pytest.raises(Exception) as e:
        host.load_folddisco_index(mine)
    assert "not ascending" in str(e.value)


@needs_reference
def test_pdb_reader_on_reference_files(host):
    files = sorted(glob.glob(REF + "/data/serine_peptidases/*.pdb")) + sorted(glob.glob(REF + "/query/*.pdb")) + \
        sorted(glob.glob(REF + "/data/homeobox/*.pdb")) + [REF + "/data/AF-P17538-F1-model_v4.pdb"] + \
        sorted(glob.glob(REF + "/data/long/*.pdb"))[:2] + sorted(glob.glob(REF + "/data/io_test/*.pdb"))
    assert len(files) > 15
    for p in files:
        _same_compact(host.read_structure_from_path(p), O.Structure.read_pdb(p).compact())
    assert host.read_structure_from_path(REF + "/data/homeobox/1akha-.pdb").num_residues == 49  # pdb.rs:142
    with pytest.raises(Exception):
        host.read_structure_from_path("/nonexistent/x.pdb")


def test_store_file_round_trip(host, tmp_path):
    """PREFIX.store (fdh_store_save / fdh_store_load): the arrays a query run attaches come back bit for bit"""
    import ctypes as C
    atoms = F.config1_atoms()
    st = host.Store()
    names = F.serine_names()
    for n in names:
        st.add(host.CompactStructure.from_atoms(atoms[n]), n)
    p = str(tmp_path / "x.store")
    st.save(p)
    back = host.Store.load(p)
    assert len(back) == len(st) == 5 and back.num_residues == st.num_residues
    assert [back.name(i) for i in range(5)] == names
    for a, b in zip(st.lookup(), back.lookup()):
        assert np.array_equal(a, b)
    va, vb = st.batch_view(), back.batch_view()
    R = st.num_residues
    for f, n in (("n_xyz", 12 * R), ("ca_xyz", 12 * R), ("cb_xyz", 12 * R), ("aa", R), ("cb_valid", R)):
        assert C.string_at(getattr(va, f), n) == C.string_at(getattr(vb, f), n)
    with open(p, "r+b") as fh:  # a truncated or foreign file is refused
        fh.truncate(1000)
    with pytest.raises(Exception):
        host.Store.load(p)
    with open(p, "wb") as fh:  # a header that promises more than the file holds is refused before any allocation
        fh.write(b"FDB2STR1" + np.array([1 << 35, 1 << 36], np.uint64).tobytes() + b"\0" * 64)
    with pytest.raises(Exception):
        host.Store.load(p)


def test_cli_fails_loudly_without_a_gpu():
    """the command-line front end has no CPU path either"""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "folddisco_b200", "folddisco-b200")
    r = subprocess.run([cli, "query", "-p", "x.pdb", "-q", "A1,A2", "-i", "nowhere"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    assert subprocess.run([cli, "version"], capture_output=True, text=True).stdout.startswith("folddisco_b200")


def test_worker_pool_runs_every_index_once_and_concurrent_regions():
    """fd_parallel (csrc/fd_ctx.cu): every worker index of a region runs exactly once, from one or several host threads
    at a time (search lanes / a prepare thread next to a search thread share the workers)"""
    import ctypes as C
    import threading
    import folddisco_b200 as fd
    L = fd.lib()
    L.fd_parallel_probe.restype = C.c_int
    L.fd_parallel_probe.argtypes = [C.c_int, C.c_int]
    for nt in (1, 2, 5, 16):
        got = L.fd_parallel_probe(nt, 200)
        assert 1 <= got <= nt  # distinct host threads that took part (a fast thread may take several indices)
    results = []

    def hammer():
        for _ in range(40):
            results.append(L.fd_parallel_probe(4, 100))

    ths = [threading.Thread(target=hammer) for _ in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert len(results) == 160 and all(1 <= r <= 4 for r in results)


def test_add_many_indexed_equals_add_many(host):
    """fdh_queries_add_many_indexed (few structures / strings referenced by index) builds the same query maps"""
    atoms = F.config1_atoms()
    motifs = [(host.CompactStructure.from_atoms(atoms[p]), q) for p, q, _ in F.MOTIFS]
    n = 23
    which = np.arange(n) % 5
    a = host.QueryBatch()
    a.add_many([motifs[k][0] for k in which], [motifs[k][1] for k in which], threads=2)
    b = host.QueryBatch()
    b.add_many_indexed([m[0] for m in motifs], [m[1] for m in motifs], which, which, threads=2)
    assert len(a) == len(b) == n and a.query_strings == b.query_strings
    for q in range(n):
        ma, mb = a.query_map(q), b.query_map(q)
        assert len(ma["hash"]) == F.MOTIFS[q % 5][2]
        for f in ("hash", "qi", "qj", "primary"):
            assert np.array_equal(ma[f], mb[f]), (q, f)
        assert np.array_equal(a.indices(q), b.indices(q))
    with pytest.raises(Exception):
        b.add_many_indexed([m[0] for m in motifs], [m[1] for m in motifs], [7], [0])


@needs_reference
def test_cif_and_gz_readers_on_reference_files(host):
    """read_structure_from_path (src/controller/io.rs:337-379): the mmCIF reader (src/structure/io/cif.rs: atom_site loop,
    auth ids, first model) and gzip inputs give the same CompactStructure as the PDB file of the same entry, on the
    reference's own io_test files.  The reference's CIF reader does not drop HETATM rows (cif.rs:262-270), so an entry
    with hetero residues that carry N and CA atoms has MORE residues than its PDB file."""
    io = REF + "/data/io_test/"

    def same(a, b):
        da, db = a.soa(), b.soa()
        return all(da[k].shape == db[k].shape and np.array_equal(da[k], db[k]) for k in da)

    af = host.read_structure_from_path(io + "cif/AF-A0A4S3KKF6-F1-model_v4.cif")
    assert af.num_residues == 738
    assert same(af, host.read_structure_from_path(io + "cif/AF-A0A4S3KKF6-F1-model_v4.pdb"))
    assert same(af, host.read_structure_from_path(io + "cif/AF-A0A4S3KKF6-F1-model_v4.cif.gz"))
    _same_compact(af, O.Structure.read_pdb(io + "cif/AF-A0A4S3KKF6-F1-model_v4.pdb").compact())
    g2f = host.read_structure_from_path(io + "cif/1G2F.cif")
    pdb = host.read_structure_from_path(REF + "/query/1G2F.pdb")
    assert g2f.num_residues == pdb.num_residues == 176
    da, db = g2f.soa(), pdb.soa()
    for k in ("ca_xyz", "n_xyz", "cb_xyz", "aa", "serial"):
        assert np.array_equal(da[k], db[k]), k
    # the chain of a residue is the chain of the atom that triggered its flush: after the last protein
    # residue the CIF loop goes on with hetero atoms of another chain, the PDB file does not
    assert np.array_equal(da["chain"][:-1], db["chain"][:-1])
    wnb, wnb_pdb = host.read_structure_from_path(io + "cif/2wnb.cif"), host.read_structure_from_path(io + "cif/2wnb.pdb")
    assert wnb.num_residues >= wnb_pdb.num_residues == 270
    assert same(wnb, host.read_structure_from_path(io + "cif/2wnb.cif.gz"))
    assert same(host.read_structure_from_path(io + "inner/1akha-.pdb.gz"), host.read_structure_from_path(io + "1akha-.pdb"))
    assert same(host.read_structure_from_path(io + "1b72a-.ent.gz"), host.read_structure_from_path(io + "1b72a-.pdb"))
    assert same(host.read_structure_from_path(io + "inner/1b72a-.ent"), host.read_structure_from_path(io + "1b72a-.pdb"))
    with pytest.raises(Exception):
        host.read_structure_from_path(io + "cif/2wnb.xyz")


def test_cif_reader_small_cases(host, tmp_path):
    """quoting, missing values, the model cut and the chain / residue-number fallbacks of cif.rs:239-259"""
    head = "data_t\n#\nloop_\n" + "".join("_atom_site.%s\n" % c for c in (
        "group_PDB", "id", "type_symbol", "label_atom_id", "label_alt_id", "label_comp_id", "label_asym_id", "label_seq_id",
        "Cartn_x", "Cartn_y", "Cartn_z", "B_iso_or_equiv", "auth_seq_id", "auth_asym_id", "pdbx_PDB_model_num"))
    rows = []
    k = 0
    for model in (1, 2):
        for res, (name, auth_seq, auth_asym) in enumerate((("ALA", "10", "A"), ("GLY", ".", "AA"), ("SER", "12", "B"))):
            for atom, (x, y, z) in ((("N"), (0.0, 0.0, 0.0)), ("CA", (1.458, 0.0, 0.0)), ("C", (2.0, 1.4, 0.0)),
                                    ('"O5\'"' if res == 2 else "O", (1.5, 2.4, 0.1)), ("CB", (2.0, -0.8, -1.2))):
                if name == "GLY" and atom == "CB":
                    continue
                k += 1
                rows.append("ATOM %d C %s . %s Z %d %.3f %.3f %.3f %s %s %s %d" % (
                    k, atom, name, res + 1, x + 3.8 * res + 50 * (model - 1), y, z, "?" if res == 1 else "42.50", auth_seq,
                    auth_asym, model))
    p = tmp_path / "t.cif"
    p.write_text(head + "\n".join(rows) + "\n#\n_other.item 1\n")
    d = host.read_structure_from_path(str(p)).soa()
    assert d["aa"].tolist() == [0, 7, 15]                       # second model ignored
    assert d["serial"].tolist() == [10, 2, 12]                  # auth_seq_id, `.` falls back to label_seq_id
    # the atoms' chains are A, Z (a two-character auth chain falls back to label_asym_id), B; a residue takes the chain of
    # the atom that triggered its flush (quirk Q2), the last one its own
    assert d["chain"].tolist() == [ord("Z"), ord("B"), ord("B")]
    assert d["b_factor"].tolist() == [1.0, 42.5, 42.5]          # `?` -> default 1.0; same flush quirk as the chain
    assert d["cb_valid"].tolist() == [1, 1, 1] and abs(d["ca_xyz"][2][0] - (1.458 + 7.6)) < 1e-4
    # gzip inputs: the same rows through zlib; a .gz file that is not a gzip stream is refused, an empty loop is empty
    import gzip
    with gzip.open(str(tmp_path / "t.cif.gz"), "wt") as f:
        f.write(head + "\n".join(rows) + "\n#\n")
    dz = host.read_structure_from_path(str(tmp_path / "t.cif.gz")).soa()
    assert all(np.array_equal(d[k], dz[k]) for k in d)
    (tmp_path / "bad.cif.gz").write_text(head + "\n".join(rows))
    with pytest.raises(Exception):
        host.read_structure_from_path(str(tmp_path / "bad.cif.gz"))
    (tmp_path / "none.cif").write_text("data_x\n_entry.id X\n")
    assert host.read_structure_from_path(str(tmp_path / "none.cif")).num_residues == 0


def test_prepared_query_inputs_equal_add_many(host):
    """QueryInputs + QueryBatch.add_prepared (inputs marshalled once) build the same query maps as add_many_indexed"""
    atoms = F.config1_atoms()
    comps = [host.CompactStructure.from_atoms(atoms[p]) for p, _, _ in F.MOTIFS]
    strings = [q for _, q, _ in F.MOTIFS]
    wc = np.array([0, 1, 2, 3, 4, 0, 2], np.uint32)
    ws = np.array([0, 1, 2, 3, 4, 0, 2], np.uint32)
    a, b = host.QueryBatch(), host.QueryBatch()
    a.add_many_indexed(comps, strings, wc, ws)
    inp = host.QueryInputs(comps, strings, wc, ws)
    b.add_prepared(inp)
    b2 = host.QueryBatch()
    b2.add_prepared(inp)  # reusable
    assert len(a) == len(b) == len(b2) == 7 and b.query_strings == a.query_strings
    for k in range(7):
        for f in ("hash", "qi", "qj", "primary"):
            assert np.array_equal(a.query_map(k)[f], b.query_map(k)[f]) and np.array_equal(a.query_map(k)[f], b2.query_map(k)[f])


def _synthetic_inputs(host, n_queries, first=0, structures=60):
    from folddisco_b200 import synth
    import bench
    db = synth.generate(structures, synth.SEED_BASE + 5)
    return host.QueryInputs(*bench.query_inputs(db, n_queries, first))


def test_query_batch_shares_structures_and_outlives_the_handles(host):
    """a batch shares the caller's CompactStructure instead of copying it (fdh_compact is immutable and reference
    counted inside the library): the query maps are the same as before, and they stay readable after every handle
    was freed"""
    atoms = F.config1_atoms()
    comps = [host.CompactStructure.from_atoms(atoms[p]) for p, _, _ in F.MOTIFS]
    strings = [q for _, q, _ in F.MOTIFS]
    a = host.QueryBatch()
    a.add_many(comps, strings)
    single = host.QueryBatch()
    for c, q in zip(comps, strings):
        single.add(c, q)
    want = [{f: a.query_map(k)[f].copy() for f in ("hash", "qi", "qj", "primary")} for k in range(len(strings))]
    idx = [a.indices(k).copy() for k in range(len(strings))]
    del comps, c  # fdh_compact_free of every handle: the batches hold the last references
    import gc
    gc.collect()
    junk = [host.CompactStructure.from_atoms(atoms[p]) for p, _, _ in F.MOTIFS]  # reuse the freed heap blocks, if any
    for k in range(len(strings)):
        assert np.array_equal(a.indices(k), idx[k]) and np.array_equal(single.indices(k), idx[k])
        for f in want[k]:
            assert np.array_equal(a.query_map(k)[f], want[k][f]) and np.array_equal(single.query_map(k)[f], want[k][f])
    # a batch added after the free of OTHER handles to the same data sees the same maps (nothing dangling is read)
    b = host.QueryBatch()
    b.add_many(junk, strings)
    for k in range(len(strings)):
        assert np.array_equal(b.query_map(k)["hash"], want[k]["hash"])
    del junk


def test_query_map_worker_builds_the_same_batches_concurrently(host):
    """QueryMapWorker: the maps of the next batch on a second host thread, while the caller builds maps of its own on
    the shared worker pool; same maps as a direct add_prepared, errors surface in take()"""
    batches = [_synthetic_inputs(host, 48, first) for first in (0, 48, 96)]
    direct = []
    for inp in batches:
        qb = host.QueryBatch()
        qb.add_prepared(inp)
        direct.append(qb)
    w = host.QueryMapWorker()
    with pytest.raises(host.FdError):
        w.take()
    w.start(batches[0])
    with pytest.raises(host.FdError):
        w.start(batches[1])
    for k in range(3):
        got = w.take()
        if k + 1 < 3:
            w.start(batches[k + 1])
        busy = host.QueryBatch()  # the caller's own parallel region, concurrent with the worker's
        busy.add_prepared(batches[k])
        w.wait()
        assert len(got) == len(direct[k]) == 48 and got.query_strings == direct[k].query_strings
        for q in range(48):
            for f in ("hash", "qi", "qj", "primary"):
                assert np.array_equal(got.query_map(q)[f], direct[k].query_map(q)[f])
                assert np.array_equal(busy.query_map(q)[f], direct[k].query_map(q)[f])
    bad = host.QueryInputs([host.CompactStructure.from_soa(np.zeros((2, 3), np.float32), np.zeros((2, 3), np.float32),
                                                           np.zeros((2, 3), np.float32), np.zeros(2, np.uint8))], ["A1-x"])
    w.start(bad)
    w.wait()
    with pytest.raises(host.FdError):
        w.take()
    w.close()


def test_bench_serving_loop_builds_one_batch_per_step(host):
    """bench.serving_loop (the e2e serving loop of the bench line) with stand-ins for the device half: every step takes a
    batch whose maps were built on the worker thread, starts exactly one more, and the batches are complete"""
    import bench
    inp = _synthetic_inputs(host, 32)
    seen = []

    class R:
        struct_offsets = [0, 7]
        match_offsets = [0, 9]

    def search(qb):
        seen.append(len(qb))
        assert qb.query_map(5)["hash"].size > 0
        return R()

    finalized = []
    ts, res = bench.serving_loop(host, None, inp, finalized.append, search, lambda fn: (fn(), 0.002, 0.002), 2, 3, (7, 9))
    assert ts == [0.002] * 3 and isinstance(res, R) and seen == [32] * 5 and len(finalized) == 5
    with pytest.raises(RuntimeError):
        bench.serving_loop(host, None, inp, finalized.append, search, lambda fn: (fn(), 0.002, 0.002), 1, 1, (1, 1))
    blk = bench.e2e_block(type("A", (), {"batch": 32})(), 2, 0.010, 0.005, None, 100, 200)
    assert blk["value"] == 64 / 0.005 and blk["one_batch_at_a_time"]["value"] == 64 / 0.010 and blk["h2d_bytes_per_step"] == 200
    blk = bench.e2e_block(type("A", (), {"batch": 32})(), 1, 0.010, None, "X: y", 100, 200)
    assert blk["value"] == 32 / 0.010 and blk["serving_loop_error"] == "X: y" and "one_batch_at_a_time" not in blk


# ---- Foldcomp input (src/structure/io/fcz.rs) ----
@pytest.fixture()
def foldcomp_codec(monkeypatch):
    try:
        so = O.build_ref()
    except Exception:
        so = None
    if so is None:
        pytest.skip("no Foldcomp codec (oracle/_ref/libfoldcomp_ffi.so) and no reference tree to build it from")
    monkeypatch.setenv("FD_FOLDCOMP_LIB", so)
    return so


@needs_reference
def test_foldcomp_database_reader(host, foldcomp_codec):
    """FoldcompDb on the reference's data/foldcomp/example_db (the database of fcz.rs:366-392): entries in key order with
    their names and keys, every entry = the oracle's CompactStructure of the atoms the codec itself returns, lookup by
    name, "DB:name" paths (controller/io.rs:303-334), the single-entry file 7m0y.fcz"""
    dbp = os.path.join(REF, "data", "foldcomp", "example_db")
    index = [ln.split("\t") for ln in open(dbp + ".index").read().splitlines()]
    lookup = dict((int(r[0]), r[1]) for r in (ln.split("\t") for ln in open(dbp + ".lookup").read().splitlines()))
    raw = open(dbp, "rb").read()
    db = host.FoldcompDb(dbp)
    assert len(db) == len(index) == 24
    assert db.keys() == sorted(int(r[0]) for r in index) and db.names() == [lookup[k] for k in db.keys()]
    assert db.find("d1asha_") == 0 and db.find("nope") == -1
    for k, (key, start, length) in enumerate(index):
        want = O.Structure.from_atoms(O.foldcomp_atoms(raw[int(start):int(start) + int(length)])).compact()
        got = db.read(k)
        assert got.num_residues == want.nres > 50
        _same_compact(got, want)
    _same_compact(db.read_by_name("d1cg5b_"), O.Structure.from_atoms(O.foldcomp_atoms(
        raw[int(index[3][1]):int(index[3][1]) + int(index[3][2])])).compact())
    with pytest.raises(host.FdError):
        db.read_by_name("nope")
    with pytest.raises(host.FdError):
        db.read(24)
    # read_compact_structure: "DB:name"
    _same_compact(host.read_structure_from_path(dbp + ":d1asha_"), O.Structure.from_atoms(O.foldcomp_atoms(
        raw[:int(index[0][2])])).compact())
    with pytest.raises(host.FdError):
        host.read_structure_from_path(dbp + ":nope")
    # one entry in a file of its own (the file of fcz.rs:300-316): 2 187 atoms, residues 449 .. 721 of chain A
    one = open(os.path.join(REF, "data", "foldcomp", "7m0y.fcz"), "rb").read()
    atoms = O.foldcomp_atoms(one)
    assert len(atoms["x"]) == 2187 and int(atoms["res_serial"][0]) == 449 and int(atoms["res_serial"][-1]) == 721
    want = O.Structure.from_atoms(atoms).compact()
    _same_compact(host.compact_from_fcz(one), want)
    _same_compact(host.read_structure_from_path(os.path.join(REF, "data", "foldcomp", "7m0y.fcz")), want)
    with pytest.raises(host.FdError):
        host.compact_from_fcz(b"not a foldcomp entry")
    # a truncated entry makes the codec throw (std::length_error through its C wrapper): an error here, not an abort.
    # In a child process: what the third-party codec does with corrupt bytes is its own (it can also crash).
    import subprocess
    import sys
    bad = str(os.path.join(os.environ.get("TMPDIR", "/tmp"), "fd_truncated_%d.fcz" % os.getpid()))
    open(bad, "wb").write(one[:100])
    code = ("import sys; sys.path.insert(0, %r)\nfrom folddisco_b200 import host\n"
            "try:\n    host.compact_from_fcz(open(%r, 'rb').read())\n    print('decoded')\n"
            "except host.FdError as e:\n    print('refused', e)\n" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), bad))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    os.remove(bad)
    assert r.returncode == 0 and "refused" in r.stdout and "codec failed" in r.stdout, (r.returncode, r.stdout, r.stderr)
    with pytest.raises(host.FdError):
        host.FoldcompDb(os.path.join(REF, "data", "foldcomp", "missing_db"))


def test_foldcomp_database_files_are_validated(host, tmp_path):
    """entries outside the database file, malformed rows and entries without a name (no codec needed)"""
    p = str(tmp_path / "db")
    open(p, "wb").write(b"FCMP" + b"\0" * 60)
    open(p + ".lookup", "w").write("0\tfirst\t0\n2\tthird\t0\n")
    open(p + ".index", "w").write("2\t32\t32\n0\t0\t32\n1\t16\t8\n")
    db = host.FoldcompDb(p)
    assert db.names() == ["first", "third"] and db.keys() == [0, 2]  # key 1 has no name: not a path (fcz.rs:211-217)
    open(p + ".index", "w").write("0\t0\t65\n")
    with pytest.raises(host.FdError):
        host.FoldcompDb(p)
    open(p + ".index", "w").write("0\tx\t6\n")
    with pytest.raises(host.FdError):
        host.FoldcompDb(p)


def test_index_of_a_foldcomp_database_records_keys_and_format(host, tmp_path):
    """an index whose structures came from a Foldcomp database: the database keys are the 5th lookup column
    (lookup.rs:36-40), PREFIX.type says input_format = "FCZDB" and names the database (config.rs:64-97); read back"""
    import ctypes as C
    from folddisco_b200 import capi
    atoms = F.config1_atoms()
    names = F.serine_names()
    store = host.Store()
    comps = []
    for n in names:
        store.add(host.CompactStructure.from_atoms(atoms[n]), n)
        comps.append(O.Structure.from_atoms(atoms[n]).compact())
    oix = O.Index.build(comps)
    hashes, offsets, values = oix.hashes, oix.offsets, oix.values
    b = capi._IndexBuffers(len(hashes), hashes.ctypes.data_as(C.POINTER(C.c_uint32)),
                           offsets.ctypes.data_as(C.POINTER(C.c_uint64)), len(values),
                           values.ctypes.data_as(C.POINTER(C.c_uint8)))
    p = capi.HashParams(0, 0, 20.0)
    ix = host.FolddiscoIndex(host._lib().fdh_index_from_buffers(C.byref(b), store.h, C.byref(p)))
    assert ix.foldcomp_db == ""
    with pytest.raises(host.FdError):
        ix.set_db_keys([1, 2])
    keys = [100, 110, 200, 201, 7]
    ix.set_db_keys(keys)
    prefix = str(tmp_path / "fcz_index")
    ix.save(store, prefix, foldcomp_db="data/foldcomp/example_db")
    rows = [ln.split("\t") for ln in open(prefix + ".lookup").read().splitlines()]
    assert [int(r[0]) for r in rows] == [0, 1, 2, 3, 4] and [int(r[4]) for r in rows] == keys
    t = open(prefix + ".type").read()
    assert 'input_format = "FCZDB"' in t and 'foldcomp_db = "data/foldcomp/example_db"' in t
    back = host.load_folddisco_index(prefix)
    assert back.foldcomp_db == "data/foldcomp/example_db"
    assert [int(host._lib().fdh_index_db_key(back.h, k)) for k in range(5)] == keys
    # a directory index names its -p argument too (the reference's default build does), but is not a Foldcomp index
    plain = host.FolddiscoIndex(host._lib().fdh_index_from_buffers(C.byref(b), store.h, C.byref(p)))
    plain.save(store, prefix + "2", foldcomp_db="data/serine_peptidases")
    assert 'input_format = "PDB"' in open(prefix + "2.type").read()
    assert host.load_folddisco_index(prefix + "2").foldcomp_db == ""


def test_foldcomp_input_fails_loudly_without_the_codec(tmp_path):
    """no codec library: decoding a Foldcomp entry is an error that says what is missing (never a silent skip)"""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r)\n"
            "from folddisco_b200 import host\n"
            "try:\n"
            "    host.compact_from_fcz(b'FCMP' + bytes(64))\n"
            "except host.FdError as e:\n"
            "    print('ERR', e)\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, FD_FOLDCOMP_LIB=str(tmp_path / "no_such_codec.so"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "ERR" in r.stdout and "FD_FOLDCOMP_LIB" in r.stdout, (r.stdout, r.stderr)


@needs_reference
def test_cli_index_reads_a_foldcomp_database(foldcomp_codec, tmp_path):
    """`index -p FOLDCOMP_DB` (build_index.rs:104-123): the 24 entries are decoded on the host, then the build needs the
    GPU; without the codec library, or for a file that is not a database, the command says so"""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU (the GPU run of this path is not part of the CPU suite)")
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "folddisco_b200", "folddisco-b200")
    dbp = os.path.join(REF, "data", "foldcomp", "example_db")
    r = subprocess.run([cli, "index", "-p", dbp, "-i", str(tmp_path / "ix"), "-v"], capture_output=True, text=True)
    assert r.returncode != 0 and "Indexing 24 files" in r.stderr and "no CPU fallback" in r.stderr
    env = dict(os.environ, FD_FOLDCOMP_LIB=str(tmp_path / "none.so"))
    r = subprocess.run([cli, "index", "-p", dbp, "-i", str(tmp_path / "ix")], capture_output=True, text=True, env=env)
    assert r.returncode != 0 and "Foldcomp codec library" in r.stderr
    r = subprocess.run([cli, "index", "-p", os.path.join(REF, "data", "foldcomp", "7m0y.fcz"), "-i", str(tmp_path / "ix")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "neither a directory nor a Foldcomp database" in r.stderr


def test_search_batches_generator_with_stand_ins(host):
    """host.search_batches: one result per batch, in order, the next batch's maps started before the current one is
    finalized; stand-ins replace the device half (finalize / search)"""
    batches = [_synthetic_inputs(host, n, first) for n, first in ((8, 0), (12, 8), (5, 20))]
    order = []
    out = list(host.search_batches(None, batches, finalize=lambda qb: order.append(("f", len(qb))),
                                   search_fn=lambda qb: (order.append(("s", len(qb))), qb.query_strings)[1]))
    assert order == [("f", 8), ("s", 8), ("f", 12), ("s", 12), ("f", 5), ("s", 5)]
    assert [len(o) for o in out] == [8, 12, 5] and out[1] == batches[1].per_query
    assert list(host.search_batches(None, [], finalize=order.append, search_fn=order.append)) == []
    # a consumer that stops early leaves no thread behind
    gen = host.search_batches(None, batches, finalize=lambda qb: None, search_fn=len)
    assert next(gen) == 8
    gen.close()


def test_pdb_decimal_fields_are_correctly_rounded(host, tmp_path):
    """the PDB reader's short-decimal route (m / 10^e in binary64, one rounding) returns the correctly rounded f32 of
    every coordinate / B-factor column: 60 000 random %8.3f / %6.2f fields, edge values included, against Python's
    decimal -> binary64 -> f32 (exact for <= 7 digits, same argument) and against the oracle's reader (strtof)"""
    rng = np.random.default_rng(7)
    n_res = 6000
    vals = rng.uniform(-999.999, 9999.999, size=(n_res, 3, 3))
    vals[:50] = rng.uniform(-0.01, 0.01, size=(50, 3, 3))
    vals[50:60] = [[[-0.0004, 0.0005, 9999.999], [-999.999, 0.001, 1.0], [16777.216 % 9999, 8388.608, 0.125]]] * 10
    bf = rng.uniform(0, 999.99, size=n_res)
    lines, want = [], np.zeros((n_res, 3, 3), np.float32)
    wb = np.zeros(n_res, np.float32)
    serial = 1
    for r in range(n_res):
        for a, name in enumerate((" N  ", " CA ", " CB ")):
            xs = ["%8.3f" % v for v in vals[r, a]]
            b = "%6.2f" % bf[r]
            lines.append("ATOM  %5d %s ALA A%4d    %s%s%s  1.00%s           C" % (serial % 100000, name, r % 10000, *xs, b))
            want[r, a] = [np.float32(float(x)) for x in xs]
            wb[r] = np.float32(float(b))
            serial += 1
    lines.append("ATOM  %5d  N   GLY A%4d    %8.3f%8.3f%8.3f  1.00  0.00           N" % (0, n_res % 10000, 0, 0, 0))
    p = str(tmp_path / "decimals.pdb")
    open(p, "w").write("\n".join(lines) + "\nEND\n")
    got = host.read_structure_from_path(p).soa()
    assert len(got["aa"]) == n_res
    assert np.array_equal(got["n_xyz"], want[:, 0]) and np.array_equal(got["ca_xyz"], want[:, 1])
    assert np.array_equal(got["cb_xyz"], want[:, 2])
    assert np.array_equal(np.signbit(got["ca_xyz"]), np.signbit(want[:, 1]))  # "-0.000" stays a negative zero
    _same_compact(host.read_structure_from_path(p), O.Structure.read_pdb(p).compact())
    # fields the short route declines go through strtof as before: exponent form, more than 7 digits, junk
    def atom(serial, name, res, resi, f1, f2, f3):
        return "ATOM  %5d %s %s A%4d    %8s%8s%8s  1.00 10.00           C" % (serial, name, res, resi, f1, f2, f3)
    odd = [atom(1, " N  ", "ALA", 1, "1.0e+1", "2.500000", "0.25"), atom(2, " CA ", "ALA", 1, "1x.0", "2.000", "3.000"),
           atom(3, " CA ", "ALA", 1, "+4.50", ".5", "6."), atom(4, " CB ", "ALA", 1, "123456.7", "12345678", "-0.00001"),
           atom(5, " N  ", "GLY", 2, "0.000", "0.000", "0.000")]
    p2 = str(tmp_path / "odd.pdb")
    open(p2, "w").write("\n".join(odd) + "\n")
    g2 = host.read_structure_from_path(p2).soa()
    assert g2["n_xyz"][0].tolist() == [10.0, 2.5, 0.25] and g2["ca_xyz"][0].tolist() == [4.5, 0.5, 6.0]
    assert g2["cb_xyz"][0].tolist() == [float(np.float32(x)) for x in (123456.7, 12345678.0, -0.00001)]
    _same_compact(host.read_structure_from_path(p2), O.Structure.read_pdb(p2).compact())


def test_parse_path_by_id_type(host):
    """src/controller/mode.rs:19-31, 70-125 (the regex AF-.+-model_v\\d: leftmost start, greedy middle)"""
    import re
    rx = re.compile(r"AF-.+-model_v\d")
    P = host.parse_path_by_id_type
    af = "data/afdb/AF-P12345-F1-model_v4.pdb"
    assert P(af, "afdb") == "AF-P12345-F1-model_v4" and P(af, "uniprot") == "P12345" and P(af, "UniProt") == "P12345"
    assert P("x/AF-A0A4S3KKF6-F1-model_v4.cif.gz", "afdb") == "AF-A0A4S3KKF6-F1-model_v4"  # stem = "...model_v4.cif"
    assert P("x/AF-A0A4S3KKF6-F1-model_v4.cif.gz", "uniprot") == "A0A4S3KKF6"
    assert P("data/serine_peptidases/4cha.pdb", "afdb") == "4cha" and P("data/serine_peptidases/4cha.pdb", "uniprot") == "4cha"
    assert P("d/pdb1abc.ent", "pdb") == "1abc" and P("d/1abc.ent", "PDB") == "1abc"
    assert P("d/e/1abc.pdb.gz", "filename") == "1abc.pdb" and P("d/e/1abc.pdb.gz", "basename") == "1abc.pdb.gz"
    assert P("d/e/1abc.pdb", "relpath") == "d/e/1abc.pdb" and P("d/e/1abc.pdb", "whatever") == "d/e/1abc.pdb"
    assert P("/", "abspath") == "/" and P("d/.hidden", "filename") == ".hidden"
    for stem in ("AF-AF-Q1-F1-model_v2-model_v3x", "zzAF-Q9-F2-model_v1_AF-Q8-F1-model_v7", "AF--model_v1", "AF-model_v1",
                 "AF-X-model_v", "AF-X-model_vv-model_v9-model_vx", "noAF", "AF-Q1-F1-model_v12"):
        m = rx.search(stem)
        want = m.group(0) if m else stem
        assert P("dir/" + stem + ".pdb", "afdb") == want, stem
        assert P("dir/" + stem + ".pdb", "uniprot") == (want.split("-")[1] if m else stem), stem


@needs_reference
def test_cif_reader_terminates_on_vertical_tab_and_form_feed(tmp_path):
    """a value ends at \\v / \\f (C isspace), so the gap between values must skip them too: a file with such a byte used
    to spin forever in the tokenizer (found by mutating the reference's 2wnb.cif); run in a child with a timeout"""
    import subprocess
    import sys
    src = open(os.path.join(REF, "data", "io_test", "cif", "2wnb.cif"), "rb").read()
    cut = src.index(b"_atom_site.group_PDB")
    cases = {"a": src[:cut - 40] + b"\x0b" + src[cut - 40:], "b": src[:cut + 3000] + b"\x0c\x0b \x0c" + src[cut + 3000:]}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name, data in cases.items():
        p = str(tmp_path / (name + ".cif"))
        open(p, "wb").write(data)
        code = ("import sys; sys.path.insert(0, %r)\nfrom folddisco_b200 import host\n"
                "try:\n    print('nres', host.read_structure_from_path(%r).num_residues)\n"
                "except host.FdError as e:\n    print('refused', e)\n" % (root, p))
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0 and ("nres" in r.stdout or "refused" in r.stdout), (r.stdout, r.stderr)
