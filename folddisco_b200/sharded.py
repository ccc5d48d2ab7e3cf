"""Hash-range sharded index across the GPUs of one box: one process per GPU (torch.distributed, NCCL over NVLink).

    rank r holds the posting lists of the hashes in [bounds[r], bounds[r+1])   (a slice of PREFIX / PREFIX.offset)
    every rank holds the compact-structure store (replicated; 38 B per residue) and the lookup

Per batch of queries (the reference's `queries.into_par_iter()` body, src/cli/workflows/query_pdb.rs:348-452):

    0. QueryBatch.set_shards(bounds)           every (query edge, owning rank) pair gets its own vote bit
    1. pair counts  -> all_reduce(SUM)          idf of every query edge needs the GLOBAL list length (query.rs:17-32);
                                                a list lives on exactly one shard, so the sum is that length
    2. fd_votes_scan on every rank              partial per-structure votes of the shard for the WHOLE batch (K3)
       all_reduce(SUM) of the vote planes       the one exchange step: match counts and fixed-point idf add, the
                                                edge bit masks are disjoint by construction so their sum is their OR
    3. rank r finishes queries [q_r, q_{r+1})   node/edge counts, length penalty, filter, idf sort, --top
                                                (fd_votes_select), then candidate verification (K6) against the store

Nothing here computes votes or decodes postings on the host: the tensors that torch.distributed reduces are views of
library-owned device memory.  The same code runs under gloo on CPU tensors in tests/test_sharded_cpu.py (host logic
only: shard planning, vote-bit assignment, merge semantics).
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

    def finalize(self, ctx, qb, dist):
        """steps 0 and 1: vote bits + global per-edge idf"""
        import torch
        qb.set_shards(self.bounds)
        counts = qb.pair_counts(ctx)
        t = torch.from_numpy(counts.astype(np.int64))
        if dist is not None and dist.is_initialized() and dist.get_backend() == "nccl":
            t = t.cuda()
        all_reduce_sum(t, dist)
        qb.finalize_with_counts(t.cpu().numpy().astype(np.uint32), self.n_structs)

    def search(self, ctx, qb, sp, dist, labels=None):
        """steps 2 and 3 -> Results of this rank's slice of the batch (query_slice)"""
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
