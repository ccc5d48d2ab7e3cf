"""`folddisco-b200 index | query` (csrc/host/fd_cli.cpp) end to end on the GPU: the reference's config 1
(README.md:200-241) through the command line -- files on disk in the reference's layout, TSV rows equal to the
README's.  PDB files are re-written from the committed atom fixtures (/root/reference does not exist on the GPU box).
"""
import os
import subprocess

import numpy as np
import pytest

import fixtures as F
import oracle_lib as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "folddisco_b200", "folddisco-b200")


def write_pdb(path, a):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        for k in range(len(a["x"])):
            f.write("ATOM  %5d %4s %3s %c%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n" % (
                (k + 1) % 100000, bytes(a["atom_name"][k]).decode(), bytes(a["res_name"][k]).decode(),
                chr(int(a["chain"][k])), int(a["res_serial"][k]), a["x"][k], a["y"][k], a["z"][k], 1.0,
                a["b_factor"][k]))
        f.write("END\n")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("cli"))
    for name, atoms in F.config1_atoms().items():
        write_pdb(os.path.join(d, name), atoms)
    os.makedirs(os.path.join(d, "idx"))
    r = subprocess.run([CLI, "index", "-p", "data/serine_peptidases", "-i", "idx/serine", "-t", "4", "-v"], cwd=d,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return d


def run(d, *args):
    r = subprocess.run([CLI] + list(args), cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return [ln.split("\t") for ln in r.stdout.splitlines()]


def test_index_files_have_the_reference_layout(workdir):
    p = os.path.join(workdir, "idx", "serine")
    assert os.path.getsize(p) == F.CONFIG1_VALUE_BYTES
    assert os.path.getsize(p + ".offset") == F.CONFIG1_OFFSET_FILE_BYTES
    with open(p + ".offset", "rb") as f:
        assert int(np.frombuffer(f.read(8), np.uint64)[0]) == F.CONFIG1_NUM_HASHES
    lookup = [ln.rstrip("\n").split("\t") for ln in open(p + ".lookup")]
    assert [r[1] for r in lookup] == ["data/serine_peptidases/%s" % n for n in
                                      ("1azw.pdb", "1ju3.pdb", "1l7a.pdb", "1pq5.pdb", "4cha.pdb")]
    for r in lookup:  # id, path, nres, plddt, db_key (lookup.rs:17-58)
        want = F.README_STRUCT_ROWS[os.path.basename(r[1])]
        assert int(r[2]) == want[4] and abs(float(r[3]) - want[5]) < 1e-3 and int(r[4]) == want[6] == int(r[0])
    t = open(p + ".type").read()
    assert 'hash_type = "PDBTrRosetta"' in t and "num_bin_dist = 0" in t and "grid_width = 20.0" in t
    assert os.path.exists(p + ".store")


def test_query_match_rows_equal_readme(workdir):
    rows = run(workdir, "query", "-p", "query/4CHA.pdb", "-q", "B57,B102,C195", "-i", "idx/serine", "--header")
    assert rows[0] == ["tid", "node_count", "idf", "rmsd", "matching_residues", "query_residues"]
    want = [["data/serine_peptidases/" + t, str(n), "%.4f" % i, "%.4f" % r, s, "B57,B102,C195"]
            for t, n, i, r, s in F.README_MATCH_ROWS_DEFAULT]
    assert rows[1:] == want  # README.md:218-223, in the README's order
    rows = run(workdir, "query", "-p", "query/4CHA.pdb", "-q", "B57,B102,C195", "-i", "idx/serine", "--ca-distance", "1.5")
    t, n, i, r, s = F.README_MATCH_ROW_1AZW_CA15
    assert ["data/serine_peptidases/" + t, str(n), "%.4f" % i, "%.4f" % r, s, "B57,B102,C195"] in rows
    rows = run(workdir, "query", "-p", "query/4CHA.pdb", "-q", "B57,B102,C195", "-i", "idx/serine", "--top", "2")
    assert rows == want[:2]  # --top limits the candidates (query_pdb.rs:404-411) and the printed rows (result.rs:466-471)


def test_query_structure_rows_equal_readme(workdir):
    rows = run(workdir, "query", "-p", "query/4CHA.pdb", "-q", "B57,B102,C195", "-i", "idx/serine", "--per-structure",
               "--header")
    assert rows[0] == ["tid", "idf", "total_match_count", "node_count", "edge_count", "max_node_cov", "min_rmsd", "nres",
                       "plddt", "matching_residues", "db_key", "query_residues"]
    got = {os.path.basename(r[0]): r for r in rows[1:]}
    # the 1azw match needs --ca-distance 1.5 under the current code (golden 3): its row stays, without
    # matches, only if filter_after_matching keeps it -- it does not at the defaults
    for tid, (idf, tot, nc, ec, nres, plddt, key) in F.README_STRUCT_ROWS.items():
        if tid == "1azw.pdb":
            continue
        r = got[tid]
        assert r[1] == "%.4f" % idf and [int(x) for x in r[2:5]] == [tot, nc, ec]
        assert int(r[7]) == nres and r[8] == "%.2f" % plddt and int(r[10]) == key and r[11] == "B57,B102,C195"
    assert got["4cha.pdb"][9] == "B57,B102,C195:0.0000;F57,F102,G195:0.0874" and got["4cha.pdb"][5:7] == ["3", "0.0000"]
    assert got["1l7a.pdb"][9] == "_,A146,A127:0.7883;_,B146,B127:0.8078"
    # idf descending (query_pdb.rs:404, sort.rs:454-458)
    idfs = [float(r[1]) for r in rows[1:]]
    assert idfs == sorted(idfs, reverse=True)


def test_query_skip_match_and_query_file(workdir):
    rows = run(workdir, "query", "-p", "query/4CHA.pdb", "-q", "B57,B102,C195", "-i", "idx/serine", "--skip-match")
    assert len(rows) == 5 and all(r[9] == "NA" for r in rows)
    assert {os.path.basename(r[0]): r[1] for r in rows} == {t: "%.4f" % v[0] for t, v in F.README_STRUCT_ROWS.items()}
    with open(os.path.join(workdir, "queries.tsv"), "w") as f:
        f.write("query/4CHA.pdb\tB57,B102,C195\tout_serine.tsv\n")
        f.write("query/1G2F.pdb\tF207,F212,F225,F229\tout_zinc.tsv\n")
    assert run(workdir, "query", "-q", "queries.tsv", "-i", "idx/serine") == []
    got = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(workdir, "out_serine.tsv"))]
    assert [r[:5] for r in got] == [["data/serine_peptidases/" + t, str(n), "%.4f" % i, "%.4f" % r, s]
                                    for t, n, i, r, s in F.README_MATCH_ROWS_DEFAULT]
    assert os.path.exists(os.path.join(workdir, "out_zinc.tsv"))


def test_query_without_store_parses_the_lookup_files(workdir):
    """an index written by the reference has no PREFIX.store: the structures are parsed from the lookup's paths"""
    p = os.path.join(workdir, "idx", "serine.store")
    os.rename(p, p + ".away")
    try:
        rows = run(workdir, "query", "-p", "query/4CHA.pdb", "-q", "B57,B102,C195", "-i", "idx/serine")
    finally:
        os.rename(p + ".away", p)
    assert [r[:5] for r in rows] == [["data/serine_peptidases/" + t, str(n), "%.4f" % i, "%.4f" % r, s]
                                     for t, n, i, r, s in F.README_MATCH_ROWS_DEFAULT]


def test_unsupported_options_fail_loudly(workdir):
    r = subprocess.run([CLI, "query", "-p", "query/4CHA.pdb", "-q", "B57,B102,C195", "-i", "idx/serine", "--mmap-on-disk"],
                       cwd=workdir, capture_output=True, text=True)
    assert r.returncode != 0
    r = subprocess.run([CLI, "benchmark", "-r", "x", "-a", "y"], cwd=workdir, capture_output=True, text=True)
    assert r.returncode != 0 and "outside the ported path" in r.stderr
    r = subprocess.run([CLI, "index", "-p", "data/serine_peptidases", "-i", "idx/x", "-y", "nonsense"], cwd=workdir,
                       capture_output=True, text=True)
    assert r.returncode != 0 and "unknown hash type" in r.stderr


def test_other_encoding_and_multiple_bins_through_the_cli(workdir):
    """`index --type pdb --multiple-bins 8-3,16-4` (PDBMotifSinCos, two bin pairs per residue pair): PREFIX.type records
    both (config.rs:64-97), `query` reads them back and its rows equal the oracle run in the same encoding."""
    mb = [(8, 3), (16, 4)]
    r = subprocess.run([CLI, "index", "-p", "data/serine_peptidases", "-i", "idx/pdbmb", "-y", "pdb", "--multiple-bins",
                        "8-3,16-4", "-t", "2"], cwd=workdir, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    t = open(os.path.join(workdir, "idx", "pdbmb.type")).read()
    assert 'hash_type = "PDBMotifSinCos"' in t and "multiple_bin = [[8, 3], [16, 4]]" in t
    atoms = F.config1_atoms()
    names = F.serine_names()
    comps = [O.Structure.from_atoms(atoms[n]).compact() for n in names]
    with O.hash_mode(1, mb):
        oix = O.Index.build(comps)
        assert os.path.getsize(os.path.join(workdir, "idx", "pdbmb")) == oix.value_bytes
        s = O.Structure.from_atoms(atoms["query/4CHA.pdb"])
        om = O.QueryMap(s.compact(), *O.parse_query_string("B57,B102,C195", s.first_chain), index=oix,
                        total_structures=len(comps))
        nres = np.array([c.nres for c in comps], np.uint64)
        plddt = np.array([c.avg_plddt for c in comps], np.float32)
        hits = O.count_query(om, oix, nres, plddt, O.CountParams.defaults(om.residue_count))
        want = []
        for nid in hits["nid"]:
            m = O.retrieve(om, comps[int(nid)])
            for k in range(len(m["rmsd"])):
                want.append((names[int(nid)], int(m["some"][k].sum()), float(m["idf"][k]), float(m["rmsd"][k]),
                             O.residues_to_string(m["some"][k], m["chain"][k], m["serial"][k])))
    rows = run(workdir, "query", "-p", "query/4CHA.pdb", "-q", "B57,B102,C195", "-i", "idx/pdbmb")
    assert len(rows) == len(want) > 0
    got = sorted((r[0], int(r[1]), r[4]) for r in rows)
    assert got == sorted((w[0], w[1], w[4]) for w in want)
    by_key = {(w[0], w[4], round(w[3], 2)): w for w in want}
    for r in rows:
        w = by_key[(r[0], r[4], round(float(r[3]), 2))]
        assert abs(float(r[2]) - w[2]) <= 1e-4 * max(1.0, w[2]) + 5e-5 and abs(float(r[3]) - w[3]) <= 1e-4 + 5e-5


def test_sort_by_and_format_output(workdir):
    """--sort-by (sort.rs) and --format-output (result.rs column registries)"""
    base = ["query", "-p", "query/4CHA.pdb", "-q", "B57,B102,C195", "-i", "idx/serine"]
    # the legacy order node_count desc, rmsd asc (sort.rs:224-229) is the order the README rows are listed in
    rows = run(workdir, *base, "--sort-by", "node_count,rmsd")
    assert rows == [["data/serine_peptidases/" + t, str(n), "%.4f" % i, "%.4f" % r, s, "B57,B102,C195"]
                    for t, n, i, r, s in F.README_MATCH_ROWS_DEFAULT]
    rows = run(workdir, *base, "--ca-distance", "1.5", "--sort-by", "node_count,rmsd")
    keys = [(-int(r[1]), float(r[3])) for r in rows]
    assert keys == sorted(keys) and len(rows) >= 7
    t, n, i, r, s = F.README_MATCH_ROW_1AZW_CA15
    assert ["data/serine_peptidases/" + t, str(n), "%.4f" % i, "%.4f" % r, s, "B57,B102,C195"] in rows
    rows = run(workdir, *base, "--sort-by", "rmsd:desc")
    rmsd = [float(r[3]) for r in rows]
    assert rmsd == sorted(rmsd, reverse=True) and len(rows) == 6
    rows = run(workdir, *base, "--format-output", "qid,tid,rmsd,u_matrix,t_vector,nid", "--header")
    assert rows[0] == ["qid", "tid", "rmsd", "u_matrix", "t_vector", "nid"]
    first = rows[1]
    assert first[0] == "query/4CHA.pdb" and first[1].endswith("4cha.pdb") and first[2] == "0.0000" and first[5] == "4"
    u, t = [float(x) for x in first[3].split(",")], [float(x) for x in first[4].split(",")]
    assert len(u) == 9 and len(t) == 3   # the query motif found in its own structure: identity, no shift
    assert np.allclose(u, np.eye(3).ravel(), atol=2e-3) and np.allclose(t, 0, atol=5e-2)
    rows = run(workdir, *base, "--per-structure", "--sort-by", "nres:asc", "--format-output", "tid,nres,plddt")
    assert [int(r[1]) for r in rows] == sorted(int(r[1]) for r in rows) and all(len(r) == 3 for r in rows)
    rows = run(workdir, *base, "--per-structure", "--sort-by", "plddt")
    p = [float(r[8]) for r in rows]
    assert p == sorted(p, reverse=True)


def test_stale_store_is_refused_and_no_store_removes_it(workdir):
    """PREFIX.store must belong to the index beside it: a store with the same structure count but other structures is
    refused (not silently used for verification), and `index --no-store` removes a store left by an earlier build."""
    import shutil
    d = workdir
    # a second index over the same number of structures in another order/naming: its store has the same count
    os.makedirs(os.path.join(d, "data", "renamed"), exist_ok=True)
    names = sorted(os.listdir(os.path.join(d, "data", "serine_peptidases")))
    for k, n in enumerate(names):
        shutil.copy(os.path.join(d, "data", "serine_peptidases", n),
                    os.path.join(d, "data", "renamed", names[(k + 1) % len(names)]))
    r = subprocess.run([CLI, "index", "-p", "data/renamed", "-i", "idx/renamed", "-t", "2"], cwd=d, capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stderr
    shutil.copy(os.path.join(d, "idx", "serine.store"), os.path.join(d, "idx", "serine.store.bak"))
    try:
        shutil.copy(os.path.join(d, "idx", "renamed.store"), os.path.join(d, "idx", "serine.store"))
        r = subprocess.run([CLI, "query", "-p", "query/4CHA.pdb", "-q", "B57,B102,C195", "-i", "idx/serine"], cwd=d,
                           capture_output=True, text=True)
        assert r.returncode != 0 and "does not belong to this index" in r.stderr
    finally:
        shutil.move(os.path.join(d, "idx", "serine.store.bak"), os.path.join(d, "idx", "serine.store"))
    assert os.path.exists(os.path.join(d, "idx", "renamed.store"))
    r = subprocess.run([CLI, "index", "-p", "data/renamed", "-i", "idx/renamed", "-t", "2", "--no-store"], cwd=d,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert not os.path.exists(os.path.join(d, "idx", "renamed.store"))


def test_whole_structure_query_with_skip_match(workdir):
    """no -q: every residue of the query structure is a query residue (query.rs:226-233).  count_query answers it
    through its global-memory path: with --skip-match the five structures come back with the oracle's idf values in
    the oracle's order; without it the candidates are verified through the general path."""
    rows = run(workdir, "query", "-p", "query/1G2F.pdb", "-i", "idx/serine", "--skip-match")
    assert 1 <= len(rows) <= 5 and all(r[9] == "NA" for r in rows)
    atoms = F.config1_atoms()
    names = F.serine_names()
    comps = [O.Structure.from_atoms(atoms[n]).compact() for n in names]
    oix = O.Index.build(comps)
    s = O.Structure.from_atoms(atoms["query/1G2F.pdb"])
    om = O.QueryMap(s.compact(), *O.parse_query_string("", s.first_chain), index=oix, total_structures=len(comps))
    nres = np.array([c.nres for c in comps], np.uint64)
    plddt = np.array([c.avg_plddt for c in comps], np.float32)
    want = O.count_query(om, oix, nres, plddt, O.CountParams.defaults(len(om.indices())))
    got = {os.path.basename(r[0]): float(r[1]) for r in rows}
    exp = {os.path.basename(names[int(n)]): float(i) for n, i in zip(want["nid"], want["idf"])}
    assert set(got) == set(exp)
    for k in got:
        assert abs(got[k] - exp[k]) <= 1e-4 * max(1.0, exp[k]) + 5e-5  # printed with four decimals
    idfs = [float(r[1]) for r in rows]
    assert idfs == sorted(idfs, reverse=True)


def test_whole_structure_query_verification(workdir):
    """no -q and no --skip-match: the candidates of a whole-chain query are verified through the general path (K4 with the
    query's ~5 000 observed pairs in global memory, grouped by amino-acid pair -> host graph step -> K5); rows equal the
    oracle's retrieval of the same query, with matches of up to ~57 residues (retrieve.rs:364-552)."""
    atoms = F.config1_atoms()
    names = F.serine_names()
    comps = [O.Structure.from_atoms(atoms[n]).compact() for n in names]
    oix = O.Index.build(comps)
    s = O.Structure.from_atoms(atoms["query/1G2F.pdb"])
    om = O.QueryMap(s.compact(), *O.parse_query_string("", s.first_chain), index=oix, total_structures=len(comps))
    nres = np.array([c.nres for c in comps], np.uint64)
    plddt = np.array([c.avg_plddt for c in comps], np.float32)
    want = O.count_query(om, oix, nres, plddt, O.CountParams.defaults(len(om.indices())))
    rows = run(workdir, "query", "-p", "query/1G2F.pdb", "-i", "idx/serine", "--format-output",
               "tid,node_count,idf,rmsd,matching_residues")
    hit_ids = [int(n) for n in want["nid"]]
    want = []
    for nid in hit_ids:
        m = O.retrieve(om, comps[nid])
        for k in range(len(m["rmsd"])):
            want.append((names[nid], int(m["some"][k].sum()), float(m["idf"][k]), float(m["rmsd"][k]),
                         O.residues_to_string(m["some"][k], m["chain"][k], m["serial"][k])))
    assert len(rows) == len(want) > 100
    assert sorted((r[0], int(r[1]), r[4]) for r in rows) == sorted((w[0], w[1], w[4]) for w in want)
    assert max(int(r[1]) for r in rows) > 16  # more matched residues than the fused kernels hold
    by_key = {}
    for w in want:
        by_key.setdefault((w[0], w[4]), []).append(w)
    for r in rows:
        w = min(by_key[(r[0], r[4])], key=lambda e: abs(e[3] - float(r[3])))
        assert abs(float(r[2]) - w[2]) <= 1e-4 * max(1.0, w[2]) + 5e-5 and abs(float(r[3]) - w[3]) <= 1e-4 * max(1.0, w[3]) + 5e-5
