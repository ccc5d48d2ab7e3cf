"""Similarity metrics of verified matches on the GPU: fd_metrics_store_batch (K7, one thread per match)
through the host search (SearchParams(want_metrics=True)) and the command line (--format-output tm_score ...,
--sort-by, --tm-score / --chamfer cutoffs, --superpose), against the oracle's restatement of
src/structure/metrics.rs:44-345 as rmsd_with_calpha_and_rottran calls it (src/controller/retrieve.rs:776-831)."""
import os
import subprocess

import numpy as np
import pytest

import fixtures as F
import oracle_lib as O

pytestmark = pytest.mark.gpu
TOL = 2e-4  # U / t come from two Kabsch implementations that agree to 1e-4 (north_star); GDT terms are counts


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
    comps = [O.Structure.from_atoms(atoms[n]).compact() for n in names]
    ix = host.FolddiscoIndex.build(ctx, store)
    ix.attach(ctx)
    store.attach(ctx)
    yield dict(ctx=ctx, fd=fd, host=host, atoms=atoms, names=names, store=store, comps=comps, ix=ix,
               oix=O.Index.build(comps))
    ctx.close()


@pytest.mark.parametrize("verify_mode", [0, 1])
def test_metrics_of_every_match_vs_oracle(env, verify_mode):
    host, ctx, store, comps = env["host"], env["ctx"], env["store"], env["comps"]
    qb = host.QueryBatch(env["ix"].params)
    oqms = []
    for path, q, _ in F.MOTIFS:
        a = env["atoms"][path]
        qb.add(host.CompactStructure.from_atoms(a), q)
        s = O.Structure.from_atoms(a)
        ch, se, subs = O.parse_query_string(q, s.first_chain)
        oqms.append(O.QueryMap(s.compact(), ch, se, subs, index=env["oix"], total_structures=len(comps)))
    qb.finalize(ctx)
    plain = host.search(ctx, qb, host.SearchParams(verify_mode=verify_mode), labels=store)
    assert plain.metrics is None
    res = host.search(ctx, qb, host.SearchParams(verify_mode=verify_mode, want_metrics=True), labels=store)
    assert res.metrics is not None and res.metrics.shape == (len(res.matches), 5)
    assert np.array_equal(res.matches["nid"], plain.matches["nid"]) and np.allclose(res.matches["rmsd"], plain.matches["rmsd"])
    checked = 0
    for k, om in enumerate(oqms):
        nq = len(om.indices())
        want = {}
        for nid in range(len(comps)):
            r = O.retrieve(om, comps[nid])
            for m in range(len(r["rmsd"])):
                key = (nid, O.residues_to_string(r["some"][m], r["chain"][m], r["serial"][m]))
                want.setdefault(key, []).append((float(r["rmsd"][m]), r["metrics"][m]))
        lo, hi = int(res.match_offsets[k]), int(res.match_offsets[k + 1])
        for m in range(lo, hi):
            row = res.matches[m]
            key = (int(row["nid"]), res.residue_string(row, nq))
            assert key in want, key
            rm, wm = min(want[key], key=lambda e: abs(e[0] - float(row["rmsd"])))  # same residues, nearest RMSD
            assert abs(rm - float(row["rmsd"])) <= 1e-4 * max(1.0, rm)
            assert np.allclose(res.metrics[m], wm, atol=TOL, rtol=TOL), (key, res.metrics[m], wm)
            # the residue indices behind the labels are kept with the metrics
            idx = res.residue_index[int(row["res_begin"]):int(row["res_begin"]) + nq]
            assert int((idx != 0).sum()) == int(row["node_count"])
            checked += 1
    assert checked >= 8
    # the exact self match of 4CHA: every metric at its best value
    best = res.metrics[[m for m in range(len(res.matches)) if res.matches[m]["rmsd"] < 1e-4][0]]
    assert abs(best[0] - 1) < 1e-4 and best[1] == 1 and best[2] == 1 and best[3] < 1e-3 and best[4] < 1e-3


def test_partial_fit_vs_oracle():
    """`--partial-fit` (retrieve.rs:773-814): matches above three residues are superposed by LMS-QCP on the device
    (fd_lmsqcp_store_batch) and report the RMSD of the inlier core; three residues or fewer keep Kabsch.  Rows (residues,
    idf, RMSD) against the oracle's LMS-QCP restatement on a synthetic database; metrics ride on the LMS superposition."""
    import folddisco_b200 as fd
    import parity
    from folddisco_b200 import host, synth
    ctx = fd.Context(0)
    b = synth.generate(400, 21, mean_len=140.0, max_len=400)
    comps = [O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"],
                                serial=np.arange(1, len(p["aa"]) + 1, dtype=np.uint64)) for p in synth.split(b)]
    store = host.Store()
    store.add_soa(b)
    ix = host.FolddiscoIndex.build(ctx, store)
    ix.attach(ctx)
    store.attach(ctx)
    bufs = ix.buffers()
    oix = O.Index.from_buffers(bufs.hashes, bufs.offsets, bufs.values)
    nres, plddt = ix.lookup()
    atoms = F.config1_atoms()
    qb = host.QueryBatch(ix.params)
    oqms = []
    for path, q, _ in F.MOTIFS:
        qb.add(host.CompactStructure.from_atoms(atoms[path]), q)
        s = O.Structure.from_atoms(atoms[path])
        ch, se, subs = O.parse_query_string(q, s.first_chain)
        oqms.append(O.QueryMap(s.compact(), ch, se, subs, index=oix, total_structures=len(comps)))
    qb.finalize(ctx)
    res = host.search(ctx, qb, host.SearchParams(partial_fit=True, want_metrics=True), labels=store)
    plain = host.search(ctx, qb, host.SearchParams(), labels=store)
    n_rows = n_lms = 0
    with O.partial_fit():
        for k, om in enumerate(oqms):
            hits, rows = parity.oracle_query(om, oix, comps, nres, plddt)
            bad = parity.diff_query(res, k, len(om.indices()), hits, rows)
            assert not bad, bad[:5]
            n_rows += len(rows)
            n_lms += sum(1 for r in rows if r[1] > 3)
    assert n_rows > 20 and n_lms > 0
    # three residues or fewer keep the Kabsch RMSD; above that the reported value is the RMSD of the inlier core under the
    # core's superposition (not comparable with the full fit either way: the core may hold every pair, and its transform
    # is the one computed before the last pair joined, lms_qcp.rs:186-189)
    assert len(res.matches) == len(plain.matches)
    big = res.matches["node_count"] > 3
    assert big.any() and not np.allclose(res.matches["rmsd"][big], plain.matches["rmsd"][big], atol=1e-3)
    assert np.allclose(res.matches["rmsd"][~big], plain.matches["rmsd"][~big], atol=1e-5)
    ctx.close()


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "folddisco_b200", "folddisco-b200")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    from test_gpu_cli import write_pdb
    d = str(tmp_path_factory.mktemp("cli_metrics"))
    for name, atoms in F.config1_atoms().items():
        write_pdb(os.path.join(d, name), atoms)
    os.makedirs(os.path.join(d, "idx"))
    r = subprocess.run([CLI, "index", "-p", "data/serine_peptidases", "-i", "idx/serine", "-t", "4"], cwd=d,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return d


def _run(d, *args):
    r = subprocess.run([CLI, "query", "-p", "query/4CHA.pdb", "-q", "B57,B102,C195", "-i", "idx/serine"] + list(args),
                       cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return [ln.split("\t") for ln in r.stdout.splitlines()]


def test_cli_metric_columns_sort_filter_and_superpose(env, workdir):
    comps, names = env["comps"], env["names"]
    s = O.Structure.from_atoms(env["atoms"]["query/4CHA.pdb"])
    ch, se, subs = O.parse_query_string("B57,B102,C195", s.first_chain)
    om = O.QueryMap(s.compact(), ch, se, subs, index=env["oix"], total_structures=len(comps))
    want = {}
    for nid in range(len(comps)):
        r = O.retrieve(om, comps[nid])
        for m in range(len(r["rmsd"])):
            want[("%s" % names[nid], O.residues_to_string(r["some"][m], r["chain"][m], r["serial"][m]))] = \
                (r["metrics"][m], float(r["idf"][m]))
    cols = "tid,matching_residues,tm_score,gdt_ts,gdt_ha,chamfer_distance,hausdorff_distance,e_value,idf"
    rows = _run(workdir, "--format-output", cols, "--header")
    assert rows[0] == cols.split(",") and len(rows) == 1 + len(F.README_MATCH_ROWS_DEFAULT)
    for r in rows[1:]:
        mt, idf = want[(r[0], r[1])]
        for j in range(5):
            assert abs(float(r[2 + j]) - float(mt[j])) <= TOL + 5e-5, (r, mt)
        # evalue_fitting (result.rs:357-378): index size 5, three query residues; printed like Rust's {:.4e}
        mu, lam = 4.2161 * np.exp(3 * 0.0489) + 3.6661, 0.2894 * np.exp(3 * -0.0762) + 0.0316
        raw = np.exp(lam * mu) / 10546.0 * 5 * 3 * np.exp(-lam * float(r[8]))
        ev = raw * 5 / (raw + 5)
        assert "e" in r[7] and "e+" not in r[7] and "e-0" not in r[7]
        assert abs(float(r[7]) - ev) <= 2e-3 * ev  # idf is read back from four printed decimals
    # --sort-by tm_score: descending TM-score (sort.rs:78-85)
    by_tm = [float(r[1]) for r in _run(workdir, "--format-output", "tid,tm_score", "--sort-by", "tm_score")]
    assert by_tm == sorted(by_tm, reverse=True)
    by_ch = [float(r[1]) for r in _run(workdir, "--format-output", "tid,chamfer_distance", "--sort-by", "chamfer")]
    assert by_ch == sorted(by_ch)
    # MatchFilter cutoffs (filter.rs:217-236)
    cut = sorted(by_tm)[len(by_tm) // 2]
    kept = _run(workdir, "--format-output", "tid,tm_score", "--tm-score", "%.4f" % (cut + 1e-4))
    assert 0 < len(kept) < len(by_tm) and all(float(r[1]) > cut for r in kept)
    kept = _run(workdir, "--format-output", "tid,hausdorff_distance", "--hausdorff", "0.2")
    assert 0 < len(kept) < len(by_tm) and all(float(r[1]) <= 0.2 for r in kept)
    # --superpose: MATCH_RESULT_SUPERPOSE_COLUMNS (result.rs:341-352)
    sup = _run(workdir, "--superpose", "--header")
    assert sup[0] == ["tid", "node_count", "idf", "rmsd", "matching_residues", "u_matrix", "t_vector",
                      "matching_coordinates", "db_key", "query_residues"]
    first = sup[1]  # the self match: identity rotation, zero translation, the CA atoms of B57, B102, C195
    assert first[0].endswith("4cha.pdb") and first[4] == "B57,B102,C195"
    U = np.array([float(x) for x in first[5].split(",")]).reshape(3, 3)
    assert np.allclose(U, np.eye(3), atol=1e-3) and np.allclose([float(x) for x in first[6].split(",")], 0, atol=1e-2)
    ca = np.array([float(x) for x in first[7].split(",")]).reshape(-1, 3)
    c = s.compact()
    d = c.soa()
    idx = [c.get_index(ord(a), b) for a, b in (("B", 57), ("B", 102), ("C", 195))]
    assert np.allclose(ca, d["ca_xyz"][idx], atol=1e-3)
    partial = [r for r in sup[1:] if "_" in r[4]][0]  # an unmatched query residue contributes no coordinates
    assert len(partial[7].split(",")) == 3 * int(partial[1])
    # --web = per-match rows with the superposition columns (QueryMode::Web, query_pdb.rs:481-493)
    assert _run(workdir, "--web", "--header") == sup
    # --partial-fit on a three-residue motif: Kabsch (retrieve.rs:774-778), so the rows do not change
    assert _run(workdir, "--partial-fit") == _run(workdir)
