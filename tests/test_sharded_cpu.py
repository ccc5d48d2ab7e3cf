"""CPU tests of the multi-GPU host logic (folddisco_b200/sharded.py + fdh_queries_set_shards): shard planning, the
vote-bit assignment that makes the ranks' edge masks disjoint, and the merge itself run by two gloo ranks through the
same all_reduce call the NCCL path uses.  The partial votes are produced here by a numpy restatement of the dense
vote layout of include/folddisco_b200.h over the ORACLE's posting lists (no kernel runs on CPU); the merged result
must equal the oracle's count_query on the whole index."""
import os
import socket
import sys

import numpy as np
import pytest

import fixtures as F
import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))


def test_plan_hash_shards_properties():
    from folddisco_b200 import sharded
    rng = np.random.default_rng(5)
    # skewed sample: few amino-acid pairs dominate
    pref = rng.choice(400, size=20000, p=rng.dirichlet(np.ones(400) * 0.3))
    hashes = ((pref // 20).astype(np.uint32) << 25) | ((pref % 20).astype(np.uint32) << 20) | rng.integers(0, 1 << 20, 20000).astype(np.uint32)
    for world in (1, 2, 3, 4, 8):
        b = sharded.plan_hash_shards(hashes, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == 1 << 32
        assert np.all(np.diff(b.astype(np.int64)) >= 0)
        assert np.all(b % (1 << sharded.SNAP_BITS) == 0)
        owner = sharded.shard_of(b, hashes)
        assert owner.min() >= 0 and owner.max() < world
        for r in range(world):
            assert np.all((hashes[owner == r] >= b[r]) & (hashes[owner == r].astype(np.uint64) < b[r + 1]))
        if world > 1:
            load = np.bincount(owner, minlength=world)
            # never worse than the ideal share plus the largest single amino-acid-pair bucket
            biggest = np.bincount(hashes >> 20).max()
            assert load.max() <= len(hashes) / world + biggest
    # empty sample: equal-width ranges
    b = sharded.plan_hash_shards(np.zeros(0, np.uint32), 4)
    assert np.array_equal(b, np.array([0, 1 << 30, 2 << 30, 3 << 30, 4 << 30], np.uint64))
    assert sharded.query_slice(10, 0, 4) == (0, 2) and sharded.query_slice(10, 3, 4) == (7, 10)
    assert sum(b - a for a, b in (sharded.query_slice(1024, r, 8) for r in range(8))) == 1024


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _database(n_structs=160, seed=31):
    from folddisco_b200 import synth
    db = synth.generate(n_structs, seed, mean_len=120.0, max_len=260, jitter=0.08, mutate=0.03,
                        template_ids=[0, 1, 2, 3, 4, 9])
    parts = synth.split(db)
    comps = [O.Compact.from_soa(p["n_xyz"], p["ca_xyz"], p["cb_xyz"], p["aa"]) for p in parts]
    return parts, comps


MOTIFS = list(F.MOTIFS) + [("query/4CHA.pdb", "B57:X,B102,C195:ST", None)]  # the last one spans many AA pairs


def _scale(n_hashes, n_structs):
    """idf fixed-point scale of the dense vote layout (narrow), fd_query.cu idf_scale + bound mode"""
    bound = np.float32(n_hashes) * np.log2(np.float32(max(n_structs, 2))).astype(np.float32)
    return np.float32(2.0) ** np.floor(np.log2(np.float32(16777215.0) / (bound + np.float32(1.0))))


def _worker(rank, world, port, out_q):
    try:
        sys.path.insert(0, os.path.dirname(HERE))
        sys.path.insert(0, HERE)
        import torch
        import torch.distributed as dist
        from folddisco_b200 import host, sharded
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        parts, comps = _database()
        N = len(comps)
        oix = O.Index.build(comps)
        nres = np.array([c.nres for c in comps], np.uint64)
        plddt = np.zeros(N, np.float32)
        # shard plan from a sample, exactly as ShardedIndex.build does
        bounds = sharded.plan_hash_shards(oix.hashes[:: 7], world)
        mine = sharded.shard_of(bounds, oix.hashes) == rank
        assert 0 < mine.sum() < len(mine)
        lens = {int(h): len(oix.entries(int(h))) for h in []}  # filled lazily below
        atoms = F.config1_atoms()
        qb = host.QueryBatch()
        structs = [host.CompactStructure.from_atoms(atoms[p]) for p, _, _ in MOTIFS]
        qb.add_many(structs, [q for _, q, _ in MOTIFS])
        qb.set_shards(bounds)
        nq = len(MOTIFS)
        # oracle query maps (for the expected answer)
        qms = []
        for path, q, _ in MOTIFS:
            s = O.Structure.from_atoms(atoms[path])
            ch, se, subs = O.parse_query_string(q, s.first_chain)
            qms.append(O.QueryMap(s.compact(), ch, se, subs, index=oix, total_structures=N))
        # ---- step 1: pair counts through the merge call ----
        # (local counts = list length if the observed hash of the pair lives on this shard)
        bits = [qb.vote_bits(q) for q in range(nq)]
        ew = max(1, max((len(b["bit_node"]) + 31) // 32 for b in bits))
        planes = 1 + ew
        straddle = 0
        votes = np.zeros((planes, nq, N), np.uint32)
        for q in range(nq):
            b = bits[q]
            owner = sharded.shard_of(bounds, b["hashes"])
            # bits of one group are consecutive and never cross a word
            g = b["bit_group"]
            for e in range(1, len(g)):
                if g[e] == g[e - 1]:
                    assert e % 32 != 0
                    straddle += 1
            # a vote bit is used by hashes of ONE rank only -> the ranks' masks are disjoint
            for bit in np.unique(b["bit_of_hash"]):
                assert len(np.unique(owner[b["bit_of_hash"] == bit])) == 1
            sc = _scale(len(b["hashes"]), N)
            for h, bit, own in zip(b["hashes"].tolist(), b["bit_of_hash"].tolist(), owner.tolist()):
                if own != rank:
                    continue
                ids = oix.entries(h)
                if len(ids) == 0:
                    continue
                idf = np.log2(np.float32(N) / np.float32(len(ids))).astype(np.float32)
                w = np.uint32(np.float32(max(idf, 0.0)) * sc + np.float32(0.5))
                votes[0, q, ids] += np.uint32((1 << 24) | int(w))
                votes[1 + bit // 32, q, ids] |= np.uint32(1 << (bit % 32))
        t = torch.from_numpy(votes.view(np.int32).reshape(-1))
        sharded.all_reduce_sum(t, dist)  # ---- step 2: the exchange ----
        merged = t.numpy().view(np.uint32).reshape(planes, nq, N)
        # ---- step 3 restated: counts from the merged planes == oracle count_query on the whole index ----
        for q in range(nq):
            b = bits[q]
            want = O.count_query(qms[q], oix, nres, plddt)
            mc = merged[0, q] >> 24
            got_ids = np.nonzero(mc)[0]
            assert got_ids.tolist() == sorted(int(n) for n in want["nid"])
            w = {int(n): (int(m), int(nc), int(ec), float(i)) for n, m, nc, ec, i in
                 zip(want["nid"], want["match_count"], want["node_count"], want["edge_count"], want["idf"])}
            sc = _scale(len(b["hashes"]), N)
            for nid in got_ids.tolist():
                word = [int(merged[1 + k, q, nid]) for k in range(ew)]
                setbits = [e for e in range(len(b["bit_node"])) if word[e // 32] >> (e % 32) & 1]
                ec = len({int(b["bit_group"][e]) for e in setbits})
                nc = len({int(b["bit_node"][e]) for e in setbits})
                idf = float(merged[0, q, nid] & 0xffffff) / float(sc) * float(nres[nid]) ** -0.5
                assert (int(mc[nid]), nc, ec) == w[nid][:3], (q, nid)
                assert abs(idf - w[nid][3]) <= 1e-4 * max(1.0, abs(w[nid][3]))
        # the variable-length gather of the sparse protocol's prepare step
        mine = [np.arange(3 + rank, dtype=np.uint32) + 100 * rank, np.zeros(0, np.uint32),
                np.full(5 * (rank + 1), 7 + rank, np.uint32)]
        g = sharded.all_gather_arrays(mine, dist)
        assert len(g) == world
        for r in range(world):
            assert g[r][0].tolist() == (np.arange(3 + r) + 100 * r).tolist()
            assert len(g[r][1]) == 0 and g[r][2].tolist() == [7 + r] * (5 * (r + 1))
        dist.barrier()
        dist.destroy_process_group()
        out_q.put((rank, "ok", straddle))
    except Exception:  # pragma: no cover - reported to the parent
        import traceback
        out_q.put((rank, traceback.format_exc(), 0))


@pytest.mark.timeout(600)
def test_two_rank_gloo_vote_merge():
    import torch.multiprocessing as mp
    from folddisco_b200 import build
    build.build()
    O.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=540) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg, straddle in sorted(res):
        assert msg == "ok", "rank %d:\n%s" % (rank, msg)
    # the substitution query must actually have exercised edges whose hashes straddle the two shards
    assert max(s for _, _, s in res) > 0
