"""GPU test of Foldcomp input through the command line: `index -p FOLDCOMP_DB`, `query -p DB:name`, and the query-side
fallback that reads an index's structures from its Foldcomp database (src/cli/workflows/build_index.rs:104-123,
query_pdb.rs:320-341).  The codec is oracle/_ref/libfoldcomp_ffi.so (built from the reference tree in the build
container, shipped with the snapshot); the database is the reference's example_db, committed under tests/golden/.
(Named to run last: written after the round's GPU minutes were spent.)"""
import os
import shutil
import subprocess

import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "folddisco_b200", "folddisco-b200")
GOLD = os.path.join(ROOT, "tests", "golden", "foldcomp")


def test_cli_index_and_query_of_a_foldcomp_database(tmp_path, monkeypatch):
    try:
        so = O.build_ref()
    except Exception:
        so = None
    if so is None or not os.path.exists(so):
        pytest.skip("no Foldcomp codec (oracle/_ref/libfoldcomp_ffi.so)")
    monkeypatch.setenv("FD_FOLDCOMP_LIB", so)
    from folddisco_b200 import host
    d = str(tmp_path)
    os.makedirs(os.path.join(d, "db"))
    os.makedirs(os.path.join(d, "idx"))
    for ext in ("", ".index", ".lookup"):
        shutil.copyfile(os.path.join(GOLD, "example_db" + ext), os.path.join(d, "db", "example_db" + ext))

    def run(*args):
        r = subprocess.run([CLI] + list(args), cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return [ln.split("\t") for ln in r.stdout.splitlines()]

    run("index", "-p", "db/example_db", "-i", "idx/fcz", "-t", "4")
    db = host.FoldcompDb(os.path.join(d, "db", "example_db"))
    lookup = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(d, "idx", "fcz.lookup"))]
    assert [r[1] for r in lookup] == db.names() and [int(r[4]) for r in lookup] == db.keys() and len(lookup) == 24
    assert [int(r[2]) for r in lookup] == [db.read(k).num_residues for k in range(24)]
    t = open(os.path.join(d, "idx", "fcz.type")).read()
    assert 'input_format = "FCZDB"' in t and 'foldcomp_db = "db/example_db"' in t
    # a motif of the first entry, queried from the database itself: the entry finds itself, residue by residue
    s = db.read(0).soa()
    i = next(i for i in range(5, 60) if all(s["aa"][j] != 255 and s["cb_valid"][j] for j in (i, i + 3, i + 7)))
    q = ",".join("%s%d" % (chr(int(s["chain"][j])), int(s["serial"][j])) for j in (i, i + 3, i + 7))
    rows = run("query", "-p", "db/example_db:" + db.names()[0], "-q", q, "-i", "idx/fcz", "--header")
    head, body = rows[0], rows[1:]
    col = {c: head.index(c) for c in ("tid", "node_count", "rmsd", "matching_residues")}
    own = [r for r in body if r[col["tid"]] == db.names()[0] and r[col["matching_residues"]] == q]
    assert own and int(own[0][col["node_count"]]) == 3 and float(own[0][col["rmsd"]]) < 0.01
    # without PREFIX.store the structures come from the Foldcomp database named in PREFIX.type: same rows
    os.remove(os.path.join(d, "idx", "fcz.store"))
    assert run("query", "-p", "db/example_db:" + db.names()[0], "-q", q, "-i", "idx/fcz", "--header") == rows
