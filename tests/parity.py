"""Row-level comparison of a GPU search result with the CPU oracle (TEST INFRASTRUCTURE; also used by bench.py's
`parity_check`).  Integer fields and matched residues must be identical; idf and RMSD within 1e-4 (north_star)."""
import numpy as np

import oracle_lib as O

TOL = 1e-4


def oracle_query(om, oix, comps, nres, plddt, top_n=None, ca_cutoff=1.0):
    """count_query + filter/sort/top + retrieval_wrapper of one query on the oracle ->
    (hits dict, match rows [(nid, node_count, idf, rmsd, residue string)])"""
    op = O.CountParams.defaults(om.residue_count, top_n=top_n if top_n is not None else O.UINT64_MAX)
    hits = O.count_query(om, oix, np.ascontiguousarray(nres, np.uint64), plddt, op)
    rows = []
    for nid in hits["nid"]:
        r = O.retrieve(om, comps[int(nid)], ca_cutoff=ca_cutoff)
        for m in range(len(r["rmsd"])):
            rows.append((int(nid), int(r["some"][m].sum()), float(r["idf"][m]), float(r["rmsd"][m]),
                         O.residues_to_string(r["some"][m], r["chain"][m], r["serial"][m])))
    return hits, rows


def diff_query(res, k, n_query_residues, hits, rows, top_n=None, id_offset=0):
    """mismatches (list of strings, empty = parity) between query k of a host.Results and the oracle's answer.
    With a top-n cut the two sides may legitimately keep different members of a group of (near-)equal idf at the
    cut: everything strictly above the cut must agree, and all rows of the structures both sides kept."""
    bad = []
    srows = res.structures(k)
    got = {int(r["nid"]) + id_offset: r for r in srows}
    want = {int(n): (int(m), int(nc), int(ec), float(i)) for n, m, nc, ec, i in
            zip(hits["nid"], hits["match_count"], hits["node_count"], hits["edge_count"], hits["idf"])}
    if top_n is None:
        if set(got) != set(want):
            bad.append("q%d: structure sets differ (%d vs %d)" % (k, len(got), len(want)))
    else:
        if len(got) != len(want):
            bad.append("q%d: %d structure rows, oracle %d" % (k, len(got), len(want)))
        if want:
            cut = min(v[3] for v in want.values())
            strict = {n for n, v in want.items() if v[3] > cut * (1 + 2 * TOL) + 1e-7}
            if not strict <= set(got):
                bad.append("q%d: %d structures above the cut are missing" % (k, len(strict - set(got))))
    common = set(got) & set(want)
    for n in common:
        g, w = got[n], want[n]
        if (int(g["total_match_count"]), int(g["node_count"]), int(g["edge_count"])) != w[:3]:
            bad.append("q%d nid %d: counts %s vs %s" % (k, n, (int(g["total_match_count"]), int(g["node_count"]),
                                                                int(g["edge_count"])), w[:3]))
        if abs(float(g["idf"]) - w[3]) > TOL * max(1.0, abs(w[3])):
            bad.append("q%d nid %d: idf %r vs %r" % (k, n, float(g["idf"]), w[3]))
    want_rows = [r for r in rows if r[0] in common]
    got_rows = [(int(m["nid"]) + id_offset, int(m["node_count"]), float(m["idf"]), float(m["rmsd"]),
                 res.residue_string(m, n_query_residues)) for m in res.sorted_matches(k)
                if int(m["nid"]) + id_offset in common]
    if sorted((r[0], r[1], r[4]) for r in got_rows) != sorted((r[0], r[1], r[4]) for r in want_rows):
        bad.append("q%d: matched residues differ (%d vs %d match rows)" % (k, len(got_rows), len(want_rows)))
        return bad
    # two components of one candidate can map to the same residues with the same RMSD but different edge sets, hence
    # different idf: idf is part of the pairing key
    key = lambda r: (r[0], r[4], round(r[3], 3), round(r[2], 2))
    for g, w in zip(sorted(got_rows, key=key), sorted(want_rows, key=key)):
        if abs(g[2] - w[2]) > TOL * max(1.0, abs(w[2])):
            bad.append("q%d nid %d: match idf %r vs %r" % (k, g[0], g[2], w[2]))
        if abs(g[3] - w[3]) > TOL * max(1.0, abs(w[3])):
            bad.append("q%d nid %d: rmsd %r vs %r" % (k, g[0], g[3], w[3]))
    return bad
