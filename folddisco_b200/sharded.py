"""Hash-range sharded index across the GPUs of one box: one process per GPU (torch.distributed, NCCL over NVLink).

    rank r holds the posting lists of the hashes in [bounds[r], bounds[r+1])   (a slice of PREFIX / PREFIX.offset)
    every rank holds the compact-structure store (replicated; 38 B per residue) and the lookup

Per batch of queries (the reference's `queries.into_par_iter()` body, src/cli/workflows/query_pdb.rs:348-452) every
rank owns a SLICE of the batch: it builds the query maps of its slice, and finishes those queries.

    prepare   0. QueryBatch.set_shards(bounds)    every (query edge, owning rank) pair gets its own vote bit
              1. all_gather of the flat scan inputs (hashes + vote bit per hash, a few hundred bytes per query)
              2. pair counts -> all_reduce(SUM)   idf of a query edge needs the GLOBAL list length (query.rs:17-32);
                                                  a list lives on exactly one shard, so the sum is that length
    search    3. fd_votes_scan_sparse             this rank's shard scanned for the WHOLE batch (K3); the non-empty
                                                  (query, structure) cells are packed per destination rank
              4. all_to_all of the records        the one exchange step: only non-empty cells travel over NVLink
              5. fd_votes_merge_begin / apply     counts and fixed-point idf add, edge bits OR (disjoint by step 0)
              6. fd_votes_select + verification   node/edge counts, length penalty, filter, idf sort, --top, then
                                                  candidate verification (K6) against the replicated store

`search_dense` is the north-star's literal form of steps 3-5 (dense vote planes + one NCCL all_reduce(SUM)); it moves
planes * batch * structures words per rank whatever the votes are, which is why the sparse form is the default.

Nothing here computes votes or decodes postings on the host: the tensors that torch.distributed moves are views of
library-owned device memory.  The same collectives run under gloo on CPU tensors in tests/test_sharded_cpu.py (host
logic only: shard planning, vote-bit assignment, gather / merge semantics).
"""
import numpy as np

from . import capi, host

SNAP_BITS = 20  # shard boundaries are multiples of 2^20: a boundary never splits one amino-acid pair (hash >> 20)


def plan_hash_shards(sample_hashes, world, snap_bits=SNAP_BITS):
    """bounds[world + 1] (uint64, bounds[0] = 0, bounds[-1] = 2^32), ascending, multiples of 2^snap_bits, chosen so
    that the sampled postings are spread as evenly as the snapping allows.  Deterministic: every rank computes the
    same bounds from the same sample."""
    nbuckets = 1 << (32 - snap_bits)
    hist = np.bincount(np.asarray(sample_hashes, np.uint32) >> np.uint32(snap_bits), minlength=nbuckets).astype(np.float64)
    cum = np.concatenate([[0.0], np.cumsum(hist)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        if total <= 0:
            b = (nbuckets * r) // world
        else:
            b = int(np.searchsorted(cum, total * r / world, side="left"))
        b = min(max(b, bounds[-1]), nbuckets)
        bounds.append(b)
    bounds.append(nbuckets)
    return (np.array(bounds, np.uint64) << np.uint64(snap_bits)).astype(np.uint64)


def shard_of(bounds, hashes):
    """owning rank of each hash"""
    return np.searchsorted(np.asarray(bounds[1:-1], np.uint64), np.asarray(hashes, np.uint64), side="right")


def query_slice(n_queries, rank, world):
    """the contiguous slice of the batch that rank finishes after the merge"""
    return (n_queries * rank) // world, (n_queries * (rank + 1)) // world


def _active(dist):
    return dist is not None and dist.is_initialized() and dist.get_world_size() > 1


def all_gather_arrays(arrays, dist, device="cpu"):
    """Every rank contributes a list of uint32 numpy arrays; returns, per rank, the list of that rank's arrays.
    Two collectives: the lengths, then one padded all_gather_into_tensor."""
    import torch
    arrays = [np.ascontiguousarray(a, np.uint32).ravel() for a in arrays]
    if not _active(dist):
        return [arrays]
    world = dist.get_world_size()
    lens = torch.tensor([len(a) for a in arrays], dtype=torch.int64, device=device)
    all_lens = torch.empty(world * len(arrays), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_lens, lens)
    all_lens = all_lens.cpu().numpy().reshape(world, len(arrays))
    width = int(all_lens.sum(axis=1).max())
    flat = np.zeros(max(width, 1), np.uint32)
    mine = np.concatenate(arrays) if arrays else np.zeros(0, np.uint32)
    flat[:len(mine)] = mine
    send = torch.from_numpy(flat.view(np.int32)).to(device)
    recv = torch.empty(world * len(flat), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy().view(np.uint32).reshape(world, len(flat))
    out = []
    for r in range(world):
        pos, row = 0, []
        for n in all_lens[r]:
            row.append(recv[r, pos:pos + int(n)].copy())
            pos += int(n)
        out.append(row)
    return out


def all_reduce_sum(tensor, dist):
    """the exchange step; separate so that the CPU (gloo) tests run exactly this call"""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


class _BatchView:
    """fd_struct_batch view with the attributes capi.Context expects of a StructBatch"""

    def __init__(self, c, keep=None):
        self.c = c
        self.n_structs = int(c.n_structs)
        self._keep = keep


class ShardedIndex:
    """This rank's hash-range shard of the index, attached to ctx."""

    def __init__(self, index, bounds, rank, world, n_structs):
        self.index, self.bounds, self.rank, self.world, self.n_structs = index, bounds, rank, world, n_structs
        self.merge_ms = 0.0
        self.merge_bytes = 0

    @classmethod
    def build(cls, ctx, store, rank, world, params=None, sample_structs=512, bounds=None):
        """Builds and attaches rank's shard from the (replicated) store: K1 hashes every structure, only the
        hashes of this rank's range are kept, K2 encodes them (fd_build_index with hash_lo / hash_hi)."""
        params = params or capi.HashParams()
        view = _BatchView(store.batch_view())
        if bounds is None:
            bounds = plan_hash_shards(cls._sample(ctx, view, params, sample_structs), world)
        buf = ctx.build_index(view, params, 0, int(bounds[rank]), int(bounds[rank + 1]))
        b = capi._IndexBuffers(buf.count, buf.hashes.ctypes.data_as(capi.C.POINTER(capi.C.c_uint32)),
                               buf.offsets.ctypes.data_as(capi.C.POINTER(capi.C.c_uint64)), buf.value_bytes,
                               buf.values.ctypes.data_as(capi.C.POINTER(capi.C.c_uint8)))
        index = host.FolddiscoIndex(host._lib().fdh_index_from_buffers(capi.C.byref(b), store.h, capi.C.byref(params)))
        index.attach(ctx)
        return cls(index, bounds, rank, world, len(store))

    @staticmethod
    def _sample(ctx, view, params, sample_structs):
        """hashes of the first sample_structs structures (the same on every rank)"""
        c = view.c
        n = int(min(sample_structs, c.n_structs))
        ro = np.ctypeslib.as_array(capi.C.cast(c.row_offsets, capi.C.POINTER(capi.C.c_uint64)), (n + 1,)).copy()
        sub = capi._StructBatch(n, ro.ctypes.data_as(capi.VP), c.n_xyz, c.ca_xyz, c.cb_xyz, c.aa, c.cb_valid)
        hashes, _ = ctx.hash_structures(_BatchView(sub, ro), params)
        return hashes

    # ---- sparse protocol: every rank owns a slice of the batch ----
    def prepare(self, ctx, qb, dist):
        """steps 0-2 for this rank's slice `qb` of the batch"""
        import torch
        nccl = _active(dist) and dist.get_backend() == "nccl"
        dev = "cuda" if nccl else "cpu"
        qb.set_shards(self.bounds)
        per_query, hashes, bits, pairs = qb.scan_arrays()
        g = all_gather_arrays([per_query, hashes, bits, pairs], dist, dev)
        self.per_query = np.concatenate([x[0] for x in g]).reshape(-1, 3)
        self.hashes = np.concatenate([x[1] for x in g])
        self.bits = np.concatenate([x[2] for x in g])
        all_pairs = np.concatenate([x[3] for x in g])
        nq = [len(x[0]) // 3 for x in g]
        self.slice_begin = np.concatenate([[0], np.cumsum(nq)]).astype(np.uint32)
        pair_begin = np.concatenate([[0], np.cumsum([len(x[3]) for x in g])])
        counts = torch.from_numpy(ctx.posting_counts(all_pairs).astype(np.int64)).to(dev)
        all_reduce_sum(counts, dist)
        counts = counts.cpu().numpy().astype(np.uint32)
        r = self.rank if _active(dist) else 0
        qb.finalize_with_counts(counts[pair_begin[r]:pair_begin[r + 1]], self.n_structs)

    def search(self, ctx, qb, sp, dist, labels=None):
        """steps 3-6 -> Results of this rank's slice `qb` (prepared by self.prepare)"""
        import torch
        lay, ptr, off, cnt = host.votes_scan_sparse(ctx, self.per_query, self.hashes, self.bits, sp.prefilter,
                                                    self.slice_begin)
        W = 1 + lay.planes
        r = self.rank if _active(dist) else 0
        sl, dense = host.votes_merge_begin(ctx, lay, len(qb))
        host.votes_apply(ctx, sl, dense, ptr + int(off[r]) * W * 4, int(cnt[r]))
        if _active(dist):
            world = dist.get_world_size()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            mine = torch.from_numpy(cnt.astype(np.int64)).cuda()
            table = torch.empty(world * world, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(table, mine)
            incoming = table.cpu().numpy().reshape(world, world)[:, r].copy()
            incoming[r] = 0
            recv = torch.empty(max(int(incoming.sum()) * W, 1), dtype=torch.int32, device="cuda")
            empty = torch.empty(0, dtype=torch.int32, device="cuda")
            inputs, outputs, pos = [], [], 0
            for d in range(world):
                n_out = int(cnt[d]) * W if d != r else 0
                inputs.append(torch.as_tensor(capi.DeviceWords(ptr + int(off[d]) * W * 4, n_out), device="cuda")
                              if n_out else empty)
                n_in = int(incoming[d]) * W
                outputs.append(recv[pos:pos + n_in] if n_in else empty)
                pos += n_in
            dist.all_to_all(outputs, inputs)
            ev1.record()
            ev1.synchronize()
            self.merge_ms += ev0.elapsed_time(ev1)
            self.merge_bytes += (int(cnt.sum()) - int(cnt[r])) * W * 4
            host.votes_apply(ctx, sl, dense, recv.data_ptr(), int(incoming.sum()))
        return host.search_from_votes(ctx, qb, sp, sl, dense, 0, len(qb), labels)

    # ---- dense protocol (every rank holds the whole batch) ----
    def finalize(self, ctx, qb, dist):
        """steps 0 and 2 when every rank holds the whole batch: vote bits + global per-edge idf"""
        import torch
        qb.set_shards(self.bounds)
        counts = qb.pair_counts(ctx)
        t = torch.from_numpy(counts.astype(np.int64))
        if dist is not None and dist.is_initialized() and dist.get_backend() == "nccl":
            t = t.cuda()
        all_reduce_sum(t, dist)
        qb.finalize_with_counts(t.cpu().numpy().astype(np.uint32), self.n_structs)

    def search_dense(self, ctx, qb, sp, dist, labels=None):
        """dense vote planes + one all_reduce(SUM) -> Results of this rank's slice of the batch (query_slice)"""
        import torch
        lay, ptr = host.votes_scan(ctx, qb, sp.prefilter)
        if lay.words:
            votes = torch.as_tensor(capi.DeviceWords(ptr, lay.words), device="cuda")
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            all_reduce_sum(votes, dist)
            ev1.record()
            ev1.synchronize()
            self.merge_ms += ev0.elapsed_time(ev1)
            self.merge_bytes += int(lay.words) * 4
        q0, q1 = query_slice(len(qb), self.rank, self.world)
        return host.search_from_votes(ctx, qb, sp, lay, ptr, q0, q1, labels)


# ----------------------------------------------------------------------------------------------------------------
# id-range shards (SURVEY 8e ablation -> the default multi-GPU partition): rank r indexes the structures
# [cuts[r], cuts[r + 1]) only; every rank scans its local index for the WHOLE batch with the global list lengths, one
# all-to-all of fixed-size per-query top-n blocks goes to the query's owner, the owner merges and verifies.  All
# collectives run inside the library (csrc/fd_comm.cu, NCCL): Python only hands the 128-byte unique id around.
# ----------------------------------------------------------------------------------------------------------------
def id_range_cuts(n_structs, world):
    """cuts[world + 1]: contiguous, near-equal id ranges"""
    return np.array([(n_structs * r) // world for r in range(world + 1)], np.int64)


def comm_init(ctx, rank, world, dist=None):
    """NCCL communicator of the library on ctx; the unique id travels over torch.distributed when world > 1"""
    uid = capi.Context.comm_unique_id()
    if world > 1:
        import torch
        t = torch.from_numpy(uid.copy())
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        uid = t.cpu().numpy()
    ctx.comm_init(uid, rank, world)


class IdRangeShards:
    """This rank's id-range shard: local index attached to ctx (local ids), the whole structure store attached for
    the verification of the rank's own queries."""

    def __init__(self, index, first_id, n_local, total_structs, rank, world):
        self.index, self.first_id, self.n_local, self.total = index, int(first_id), int(n_local), int(total_structs)
        self.rank, self.world = rank, world
        self.table_bytes, self.table_s = 0, 0.0

    @classmethod
    def build(cls, ctx, db, rank, world, hash_params=None, pair_table=False):
        """db: SoA dict of the whole database (row_offsets, n_xyz, ca_xyz, cb_xyz, aa); the rank builds the index of its
        id range on its GPU.  Returns (shards, full_store)"""
        S = len(db["row_offsets"]) - 1
        cuts = id_range_cuts(S, world)
        lo, hi = int(cuts[rank]), int(cuts[rank + 1])
        ro = db["row_offsets"].astype(np.int64)
        sl = slice(ro[lo], ro[hi])
        local = dict(row_offsets=(db["row_offsets"][lo:hi + 1] - db["row_offsets"][lo]).astype(np.uint64),
                     n_xyz=db["n_xyz"][sl], ca_xyz=db["ca_xyz"][sl], cb_xyz=db["cb_xyz"][sl], aa=db["aa"][sl])
        local_store = host.Store()
        local_store.add_soa(local)
        index = host.FolddiscoIndex.build(ctx, local_store, hash_params)
        index.attach(ctx)
        full = host.Store()
        full.add_soa(db)
        import time
        t0 = time.perf_counter()
        tb = full.attach(ctx, pair_table=pair_table, hash_params=index.params)
        sh = cls(index, lo, hi - lo, S, rank, world)
        sh.table_bytes, sh.table_s = tb, time.perf_counter() - t0
        return sh, full

    def prepare(self, ctx, qb):
        qb.finalize_sharded(ctx, self.first_id, self.total)

    def search(self, ctx, qb, params=None, labels=None):
        return host.search_sharded(ctx, qb, params, labels)
