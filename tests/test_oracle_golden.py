"""Pins the CPU oracle to every known-answer vector the reference holds for the hot path.

CPU only.  The GPU parity tests (tests/test_gpu_*.py) then compare the CUDA path with this oracle.
"""
import glob
import os

import numpy as np
import pytest

import fixtures as F
import oracle_lib as O
from conftest import REF, needs_reference


@pytest.fixture(scope="module")
def config1():
    atoms = F.config1_atoms()
    names = F.serine_names()
    comps = [O.Structure.from_atoms(atoms[n]).compact() for n in names]
    ix = O.Index.build(comps)
    nres = np.array([c.nres for c in comps], np.uint64)
    plddt = np.array([c.avg_plddt for c in comps], np.float32)
    qs = O.Structure.from_atoms(atoms["query/4CHA.pdb"])
    qc = qs.compact()
    ch, se, subs = O.parse_query_string("B57,B102,C195", qs.first_chain)
    qm = O.QueryMap(qc, ch, se, subs, index=ix, total_structures=len(comps))
    return dict(names=names, comps=comps, ix=ix, nres=nres, plddt=plddt, qc=qc, qm=qm)


def test_kat_edge_hashes(config1):
    """src/controller/graph.rs:71-79"""
    qc = config1["qc"]
    idx = {"B57": qc.get_index(ord("B"), 57), "B102": qc.get_index(ord("B"), 102), "C195": qc.get_index(ord("C"), 195)}
    for (a, b), want in F.KAT_EDGE_HASHES.items():
        f = qc.pair_feature(idx[a], idx[b])
        assert f is not None
        assert O.perfect_hash(f) == want
        assert O.perfect_hash(f, 16, 4) == want


def test_kat_hashes_same_with_libm(config1):
    """The six KATs do not depend on which libm rounds the last ulp."""
    qc = config1["qc"]
    O.lib().fdo_set_math_mode(1)
    try:
        idx = {"B57": qc.get_index(ord("B"), 57), "B102": qc.get_index(ord("B"), 102),
               "C195": qc.get_index(ord("C"), 195)}
        for (a, b), want in F.KAT_EDGE_HASHES.items():
            assert O.perfect_hash(qc.pair_feature(idx[a], idx[b])) == want
    finally:
        O.lib().fdo_set_math_mode(0)


def test_config1_index_sizes(config1, tmp_path):
    """SURVEY section 4 golden 4: 217 612 hashes, 225 674 posting bytes, .offset = 2 611 360 B"""
    ix = config1["ix"]
    assert ix.count == F.CONFIG1_NUM_HASHES
    assert ix.value_bytes == F.CONFIG1_VALUE_BYTES
    prefix = str(tmp_path / "idx")
    ix.save(prefix)
    assert os.path.getsize(prefix) == F.CONFIG1_VALUE_BYTES
    assert os.path.getsize(prefix + ".offset") == F.CONFIG1_OFFSET_FILE_BYTES
    back = O.Index.load(prefix)
    assert np.array_equal(back.hashes, ix.hashes)
    assert np.array_equal(back.offsets, ix.offsets)
    assert np.array_equal(back.values, ix.values)
    h = ix.hashes
    assert np.all(h[1:] > h[:-1])


def test_index_codec_roundtrip(config1):
    """indextable.rs:397-463: every posting list decodes to strictly ascending ids < N and the CSR and
    two-pass builders agree."""
    ix, comps = config1["ix"], config1["comps"]
    rows = [c.hashes(sorted_unique=True) for c in comps]
    ro = np.zeros(len(rows) + 1, np.uint64)
    ro[1:] = np.cumsum([len(r) for r in rows])
    ix2 = O.Index.from_csr(np.concatenate(rows), ro)
    assert np.array_equal(ix2.hashes, ix.hashes) and np.array_equal(ix2.values, ix.values)
    member = {}
    for sid, r in enumerate(rows):
        for h in r[:: max(1, len(r) // 50)]:
            member.setdefault(int(h), None)
    for h in list(member)[:300]:
        ids = ix.entries(h)
        want = [sid for sid, r in enumerate(rows) if h in set(r.tolist())]
        assert ids.tolist() == want


def test_varint_multibyte():
    """ids >= 128 need two LEB128 bytes; id 0 encodes as a single 0x00 (indextable.rs:397-418)."""
    S = 40000
    hashes = np.array([7, 7, 7, 9], np.uint32)
    ro = np.zeros(S + 1, np.uint64)
    sids = [0, 200, 39999]
    # rows: structure 0 -> {7}, 200 -> {7}, 39999 -> {7}, and structure 300 -> {9}
    counts = np.zeros(S, np.int64)
    for s in sids:
        counts[s] = 1
    counts[300] = 1
    ro[1:] = np.cumsum(counts)
    flat = np.array([7, 7, 9, 7], np.uint32)  # in row order: 0, 200, 300, 39999
    ix = O.Index.from_csr(flat, ro)
    assert ix.hashes.tolist() == [7, 9]
    assert ix.entries(7).tolist() == [0, 200, 39999]
    assert ix.entries(9).tolist() == [300]
    v = ix.values.tolist()
    # 0 -> 00 ; 200 -> C8 01 ; 39799 = 0x9B77 -> F7 B6 02 ; 300 -> AC 02
    assert v == [0x00, 0xC8, 0x01, 0xF7, 0xB6, 0x02, 0xAC, 0x02]
    assert ix.offsets.tolist() == [0, 6, 8]


def test_config1_query_map(config1):
    qm = config1["qm"]
    e = qm.entries()
    assert len(e["hash"]) == F.CONFIG1_NUM_QUERY_HASHES
    assert sorted(int(h) for h, p in zip(e["hash"], e["primary"]) if p) == sorted(F.KAT_EDGE_HASHES.values())
    assert len(qm.indices()) == 3


def test_readme_structure_rows(config1):
    """README.md:237-241"""
    hits = O.count_query(config1["qm"], config1["ix"], config1["nres"], config1["plddt"])
    got = {}
    for k in range(len(hits["nid"])):
        nid = int(hits["nid"][k])
        tid = os.path.basename(config1["names"][nid])
        got[tid] = ("%.4f" % hits["idf"][k], int(hits["match_count"][k]), int(hits["node_count"][k]),
                    int(hits["edge_count"][k]), int(config1["nres"][nid]), "%.4f" % config1["plddt"][nid], nid)
    want = {t: ("%.4f" % r[0], r[1], r[2], r[3], r[4], "%.4f" % r[5], r[6]) for t, r in F.README_STRUCT_ROWS.items()}
    assert got == want
    # query_pdb.rs:404: sorted by idf descending
    assert np.all(np.diff(hits["idf"]) <= 0)


def _match_rows(config1, ca_cutoff):
    rows = []
    for nid, comp in enumerate(config1["comps"]):
        r = O.retrieve(config1["qm"], comp, ca_cutoff=ca_cutoff)
        for m in range(len(r["rmsd"])):
            rows.append((os.path.basename(config1["names"][nid]), int(r["some"][m].sum()), "%.4f" % r["idf"][m],
                         "%.4f" % r["rmsd"][m], O.residues_to_string(r["some"][m], r["chain"][m], r["serial"][m])))
    return rows


def test_readme_match_rows(config1):
    """README.md:218-224 at default flags (six rows) and the 1azw row at --ca-distance 1.5"""
    want = sorted((t, n, "%.4f" % i, "%.4f" % r, s) for t, n, i, r, s in F.README_MATCH_ROWS_DEFAULT)
    assert sorted(_match_rows(config1, 1.0)) == want
    t, n, i, r, s = F.README_MATCH_ROW_1AZW_CA15
    assert (t, n, "%.4f" % i, "%.4f" % r, s) in _match_rows(config1, 1.5)


def test_motif_query_hash_counts():
    """16 / 42 / 172 / 97 / 67 query hashes for the five shipped motifs at -d 0.5 -a 5"""
    atoms = F.config1_atoms()
    for path, q, want in F.MOTIFS:
        s = O.Structure.from_atoms(atoms[path])
        ch, se, subs = O.parse_query_string(q, s.first_chain)
        qm = O.QueryMap(s.compact(), ch, se, subs)
        assert len(qm.entries()["hash"]) == want, path


def test_parse_query_string():
    """src/controller/query.rs:425-465"""
    ch, se, subs = O.parse_query_string("A250,A232,A269", ord("A"))
    assert list(zip(ch.tolist(), se.tolist())) == [(65, 250), (65, 232), (65, 269)] and subs == [None] * 3
    ch, se, subs = O.parse_query_string("A250-252,B232:H,269:NDp", ord("C"))
    assert list(zip(ch.tolist(), se.tolist())) == [(65, 250), (65, 251), (65, 252), (66, 232), (67, 269)]
    assert subs == [None, None, None, [8], [2, 3, 1, 8, 11]]
    ch, se, subs = O.parse_query_string("1-3:X", ord("1"))  # non-alphabetic default chain -> 'A'
    assert ch.tolist() == [65, 65, 65] and subs[0] == list(range(20))
    assert O.parse_query_string("", ord("A"))[0].size == 0


def test_kabsch_kat():
    """src/structure/kabsch.rs:560-615"""
    src = [[6.994, 8.354, 42.405], [9.429, 7.479, 48.266], [5.547, 0.158, 42.050]]
    t1 = [[-13.958, -1.741, -4.223], [-12.833, 3.134, -7.780], [-5.720, -2.218, -3.368]]
    t2 = [[-4.924, 5.813, -9.485], [-0.499, 10.073, -8.059], [-0.792, 0.658, -4.430]]
    for t in (t1, t2):
        U, tr, rmsd = O.kabsch(t, src)  # set_atoms(fixed=src, moving=t) rotates t onto src
        assert rmsd < 0.2
        moved = np.asarray(t, np.float64) @ U.T.astype(np.float64) + tr
        assert abs(np.sqrt(((moved - np.asarray(src)) ** 2).sum() / 3) - rmsd) < 1e-4
        assert abs(np.linalg.det(U.astype(np.float64)) - 1) < 1e-5
    c = [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]]
    assert O.kabsch(c, c)[2] < 1e-6


def test_aa_map():
    """src/utils/convert.rs:53-81, 135-166"""
    assert O.lib().fdo_map_aa_to_u8(b"ALA") == 0 and O.lib().fdo_map_aa_to_u8(b"VAL") == 19
    assert O.lib().fdo_map_aa_to_u8(b"MSE") == 12 and O.lib().fdo_map_aa_to_u8(b"SEC") == 4
    assert O.lib().fdo_map_aa_to_u8(b"HOH") == 255 and O.lib().fdo_map_aa_to_u8(b"UNK") == 255


def test_exact_math_vs_libm_hash_agreement(config1):
    """Honest statement of the unpinned last ulp: hashes of every config-1 structure computed with
    the exact binary64-evaluated functions and with this machine's glibc must agree."""
    exact = [c.hashes() for c in config1["comps"]]
    O.lib().fdo_set_math_mode(1)
    try:
        libm = [c.hashes() for c in config1["comps"]]
    finally:
        O.lib().fdo_set_math_mode(0)
    diff = sum(int((a != b).sum()) for a, b in zip(exact, libm))
    total = sum(len(a) for a in exact)
    assert total > 250000
    assert diff == 0, "%d of %d pair hashes differ between exact math and glibc" % (diff, total)


def test_exact_math_is_correctly_rounded():
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 160
    rng = np.random.default_rng(7)
    L = O.lib()

    def cr(v):
        with mp.workprec(24):
            return np.float32(float(+v))

    for x in rng.uniform(-3.5, 3.5, 2000).astype(np.float32):
        assert np.float32(L.fdo_math_sinf(float(x))) == cr(mp.sin(mp.mpf(float(x))))
        assert np.float32(L.fdo_math_cosf(float(x))) == cr(mp.cos(mp.mpf(float(x))))
    for x in rng.uniform(-1, 1, 2000).astype(np.float32):
        assert np.float32(L.fdo_math_acosf(float(x))) == cr(mp.acos(mp.mpf(float(x))))
    for y, x in rng.normal(0, 1, (2000, 2)).astype(np.float32):
        assert np.float32(L.fdo_math_atan2f(float(y), float(x))) == cr(mp.atan2(mp.mpf(float(y)), mp.mpf(float(x))))
    assert np.isnan(L.fdo_math_acosf(1.0000001)) and np.isnan(L.fdo_math_sinf(float("nan")))
    assert L.fdo_math_acosf(1.0) == 0.0 and np.float32(L.fdo_math_acosf(-1.0)) == np.float32(np.pi)


# ---- checks that need the reference tree (run in the build container) -------------------------


@needs_reference
def test_pdb_parser_matches_fixture():
    atoms = F.config1_atoms()
    for name, want in atoms.items():
        got = O.Structure.read_pdb(os.path.join(REF, name)).atoms()
        for k in want:
            assert np.array_equal(got[k], want[k]), (name, k)


@needs_reference
def test_pdb_residue_count_1akha():
    """src/structure/io/pdb.rs:142 and core.rs:514: 49 residues, every GLY gets a virtual CB"""
    s = O.Structure.read_pdb(REF + "/data/homeobox/1akha-.pdb")
    c = s.compact()
    assert s.num_residues == 49 and c.nres == 49
    d = c.soa()
    gly = [i for i in range(c.nres) if bytes(d["res_name"][i]) == b"GLY"]
    assert gly and all(d["cb_valid"][i] for i in gly)


@needs_reference
def test_long_structures_hash_agreement_libm():
    """data/long: 16k-residue inputs; exact-vs-glibc disagreement count on a bounded sample of rows."""
    p = sorted(glob.glob(REF + "/data/long/*.pdb"))[0]
    c = O.Structure.read_pdb(p).compact()
    d = c.soa()
    sub = O.Compact.from_soa(d["n_xyz"][:1500], d["ca_xyz"][:1500], d["cb_xyz"][:1500], d["aa"][:1500],
                             d["cb_valid"][:1500])
    a = sub.hashes()
    O.lib().fdo_set_math_mode(1)
    try:
        b = sub.hashes()
    finally:
        O.lib().fdo_set_math_mode(0)
    assert len(a) == len(b) and int((a != b).sum()) == 0
