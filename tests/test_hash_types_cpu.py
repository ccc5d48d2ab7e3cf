"""The reference's other encodings (`--type`) and `--multiple-bins`: the header the kernels compile
(csrc/fd_hashtypes.cuh), built for the host, and the host's make_query_map against the oracle's independent restatement
of src/controller/feature.rs:11-190, src/geometry/{pdb_motif,pdb_motif_sincos,trrosetta,ppf,folddisco_angle,
folddisco_dist}.rs and src/controller/query.rs:53-206.  No GPU needed: K1 / K4 call these same functions per pair
(tests/test_gpu_hash_types.py runs them on the device).

Pins the reference holds for these encodings: only the amino-acid fields of two decoded hashes
(src/geometry/ppf.rs:173-183, src/geometry/trrosetta.rs:186-208) -- checked below; everything else about them is
"parity unpinned" (oracle == product, both written from the reference's source)."""
import numpy as np
import pytest

import fixtures as F
import oracle_lib as O

# (FD_HASH_* of the library, HashType index of the reference / oracle, bin pairs to try)
TYPES = [
    ("PDBMotif", 1, 0, [(0, 0), (12, 6), (40, 40)]),
    ("PDBMotifSinCos", 2, 1, [(0, 0), (16, 4), (5, 0)]),
    ("TrRosetta", 3, 2, [(0, 0), (6, 3), (20, 9)]),
    ("PDBTrRosetta", 4, 3, [(0, 0), (8, 3), (0, 3)]),
    ("PointPairFeature", 5, 4, [(0, 0), (12, 5)]),
    ("TertiaryInteraction", 6, 5, [(0, 0), (12, 6), (20, 9)]),
    ("Hybrid", 7, 6, [(0, 0), (8, 3)]),
    ("FolddiscoAngle", 8, 7, [(0, 0), (6, 12)]),
    ("FolddiscoDist", 9, 8, [(0, 0), (20, 10)]),
]


@pytest.fixture(scope="module")
def structs():
    atoms = F.config1_atoms()
    out = []
    for name in ("data/serine_peptidases/1pq5.pdb", "query/1G2F.pdb", "query/2N6N.pdb"):
        c = O.Structure.from_atoms(atoms[name]).compact()
        out.append((name, c, c.soa()))
    return out


@pytest.mark.parametrize("name,fd_type,ref_type,bins", TYPES, ids=[t[0] for t in TYPES])
def test_typed_hashes_equal_oracle(structs, name, fd_type, ref_type, bins):
    """every ordered pair of three shipped structures: same hashes in the same order, for default and odd bin counts"""
    from folddisco_b200 import capi
    total = 0
    for sname, c, d in structs:
        for nbd, nba in bins:
            with O.hash_mode(ref_type):
                want = c.hashes(nbd, nba, 20.0)
            got = capi.typed_hash_host(d["n_xyz"], d["ca_xyz"], d["cb_xyz"], d["aa"], d["cb_valid"],
                                       capi.HashParams(nbd, nba, 20.0, fd_type))
            assert len(got) == len(want) and np.array_equal(got, want), (sname, nbd, nba)
            total += len(got)
    assert total > 50_000


@pytest.mark.parametrize("fd_type,ref_type", [(0, 3), (2, 1), (3, 2), (6, 5), (7, 6), (9, 8)])
def test_multiple_bins_equal_oracle(structs, fd_type, ref_type):
    from folddisco_b200 import capi
    mb = [(16, 4), (8, 3), (4, 2)]
    sname, c, d = structs[1]
    with O.hash_mode(ref_type, mb):
        want = c.hashes(0, 0, 20.0)
    got = capi.typed_hash_host(d["n_xyz"], d["ca_xyz"], d["cb_xyz"], d["aa"], d["cb_valid"],
                               capi.HashParams(0, 0, 20.0, fd_type, multiple_bins=mb))
    assert len(got) % 3 == 0 and np.array_equal(got, want)
    with O.hash_mode(ref_type):  # bin pair k of the list == a single-bin index with that pair
        single = c.hashes(8, 3, 20.0)
    assert np.array_equal(got[1::3], single)


@pytest.mark.parametrize("name,fd_type,ref_type,bins", TYPES, ids=[t[0] for t in TYPES])
def test_is_symmetric_equals_oracle(structs, name, fd_type, ref_type, bins):
    from folddisco_b200 import capi
    sname, c, d = structs[0]
    with O.hash_mode(ref_type):
        hashes = np.unique(c.hashes(0, 0, 20.0))
        want = np.array([O.lib().fdo_hash_is_symmetric(int(h)) for h in hashes])
    got = np.array([capi.typed_is_symmetric_host(fd_type, h) for h in hashes])
    assert np.array_equal(got, want)
    if name not in ("PDBMotif", "PDBMotifSinCos", "TertiaryInteraction"):
        assert 0 < want.sum() < len(want)


def test_reference_pins():
    """src/geometry/ppf.rs:173-183 and src/geometry/trrosetta.rs:186-208: the amino acids decode back"""
    rad = np.float32(np.pi) / np.float32(180.0)
    with O.hash_mode(4):
        f = np.zeros(9, np.float32)
        f[:6] = [0, 7, 7.5, np.float32(120.0) * rad, np.float32(45.0) * rad, np.float32(-60.0) * rad]
        h = O.perfect_hash_raw(f, 8, 3)
        assert (h >> 27) & 31 == 0 and (h >> 22) & 31 == 7
    with O.hash_mode(2):
        f = np.zeros(9, np.float32)
        f[:8] = [0, 1, 5.0] + [np.float32(v) * rad for v in (-10.0, 0.0, 10.0, 45.0, 15.0)]
        h = O.perfect_hash_raw(f, 8, 3)
        assert ((h >> 23) & 511) // 20 == 0 and ((h >> 23) & 511) % 20 == 1


def test_refused_types():
    from folddisco_b200 import capi
    for t in (10, 99):
        with pytest.raises(capi.FdError):
            capi.typed_hash_host(np.zeros((2, 3)), np.zeros((2, 3)), np.zeros((2, 3)), np.zeros(2, np.uint8), None,
                                 capi.HashParams(0, 0, 20.0, t))


QUERY_CASES = [("query/4CHA.pdb", "B57,B102,C195"), ("query/1G2F.pdb", "F207,F212,F225,F229"),
               ("query/2MNR.pdb", "164:H,195,221,247:ND,297:H"), ("query/4CHA.pdb", "B57:X,B102,C195:ST")]


@pytest.mark.parametrize("name,fd_type,ref_type,bins", TYPES, ids=[t[0] for t in TYPES])
@pytest.mark.parametrize("multi", [False, True])
def test_query_map_equals_oracle(name, fd_type, ref_type, bins, multi):
    """make_query_map (query.rs:208-329) per encoding: same hashes, residues and insertion order; the distance /
    angle expansions perturb the encoding's own feature slots (feature.rs:269-289)"""
    from folddisco_b200 import capi, host
    atoms = F.config1_atoms()
    mb = [(0, 0), (8, 3)] if multi else []
    nbd, nba = bins[1]
    for dist_thr, angle_thr in (((0.5,), (5.0,)), ((0.5, 1.0), (5.0, 10.0))):
        qb = host.QueryBatch(capi.HashParams(nbd, nba, 20.0, fd_type, multiple_bins=mb), dist_thr=dist_thr,
                             angle_thr=angle_thr)
        for path, q in QUERY_CASES:
            qb.add(host.CompactStructure.from_atoms(atoms[path]), q)
        for k, (path, q) in enumerate(QUERY_CASES):
            s = O.Structure.from_atoms(atoms[path])
            ch, se, subs = O.parse_query_string(q, s.first_chain)
            with O.hash_mode(ref_type, mb):
                om = O.QueryMap(s.compact(), ch, se, subs, nbin_dist=nbd, nbin_angle=nba, dist_thr=dist_thr,
                                angle_thr=angle_thr)
                e = om.entries()
            g = qb.query_map(k)
            assert len(e["hash"]) > 0
            for f in ("hash", "qi", "qj", "primary"):
                assert np.array_equal(e[f], g[f]), (name, path, f)
