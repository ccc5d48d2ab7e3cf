// fd_query.cu -- K3: posting-list scan + per-structure vote (query path (ii), prefilter stage).
//
// Replaces count_query (reference src/controller/count_query.rs:82-220), FolddiscoIndex::get_entries
// (src/index/indextable.rs:44-86, 439-463), StructureFilter::filter_before_matching (src/controller/filter.rs:
// 76-100) and the idf sort + --top truncation of src/cli/workflows/query_pdb.rs:395-411, for a batch of queries.
//
// HBM layout of the attached index (fd_index_attach):
//   hashes[count] u32 ascending | offsets[count+1] u64 | values[value_bytes] raw delta+LEB128 bytes (the file,
//   unmodified) | counts[count] u32 postings per list | dir[2^20+1] u32 bucket directory over hash>>12 |
//   skip_off u8 / skip_id u32 [value_bytes/64+2]: for every 64-byte granule of values[] the offset of the first
//   varint that STARTS at or after the granule boundary and the running structure id of its list at that point,
//   so a list can be entered at any granule without decoding its prefix (the LEB128 stream is otherwise strictly
//   sequential) and every granule can be decoded independently of its neighbours.
//
// Kernels per batch:
//   k3_lookup : one thread per query hash: directory + binary search -> byte range, posting count, idf weight
//   k3_scan   : one CTA per (structure-id tile, query).  The tile's votes live in shared memory (packed
//               match_count|idf fixed point, plus edge-bitmask planes).  Work items are the 64-byte granules of
//               the query's lists that intersect the tile; a granule is decoded by 8 lanes (one aligned 8-byte
//               load each, 4 granules = 256 contiguous-per-granule bytes per warp step): every lane finds the
//               varints that start in its 8 bytes from the terminator bits, decodes them from its own bytes plus
//               the next lane's, an 8-lane scan turns deltas into ids, and votes are shared-memory atomics.
//               The epilogue compacts the non-empty cells warp by warp, derives node/edge counts, applies the
//               length penalty and the structure filter and appends survivors to the query's hit region --
//               or (multi-GPU) dumps the tile's planes to the dense partial-vote buffer.
//   k3_select_dense : the same epilogue over merged dense votes (multi-GPU, after the all-reduce).
//   (cub segmented sort) + k3_gather: order hits by (idf desc, nid asc) and keep top_n.
// Only posting bytes, skip entries and survivors touch HBM; the N-sized per-node-group arrays and O(N * edges)
// bit sweeps of the reference never exist.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>

#include "fd_async.cuh"
#include "fd_common.cuh"

void fd_ctx_release_index(fd_ctx *ctx);

namespace {

constexpr int DIR_SHIFT = 12;
constexpr uint32_t DIR_SIZE = 1u << (32 - DIR_SHIFT);
constexpr int SKIP_SHIFT = 6; // 64-byte granules
constexpr uint32_t SKIP_BYTES = 1u << SKIP_SHIFT;
constexpr uint32_t VALUES_PAD = 256; // readable bytes after the last posting byte (vector loads of the last granule)

constexpr int K3_THREADS = 256;      // k3_select_dense and the default CTA of k3_scan
constexpr int K3_MAX_THREADS = 1024; // k3_scan with one large vote tile per SM
constexpr int K3_WARPS = K3_THREADS / 32;
constexpr int K3_WQ = 160;               // per-warp compaction queue of the epilogue (31 pending + 128 new)
constexpr int K3_MAX_HASHES_NARROW = 255; // match_count fits 8 bits
constexpr int K3_MAX_HASHES = 4095;       // smem prefix array
constexpr int K3_MAX_EDGE_WORDS = 8;      // 256 edges
constexpr int K3_MAX_NODES = 256;

// ------------------------------------------------------------------------------------------------
// attach-time kernels
// ------------------------------------------------------------------------------------------------

// dir[b] = first position whose bucket (hash >> DIR_SHIFT) is >= b, for b in [0, DIR_SIZE]: one thread per bucket,
// lower bound by binary search (a thread per hash walking the empty buckets after it serialises on sparse indexes)
__global__ void k3_build_dir(const uint32_t *hashes, uint64_t count, uint32_t *dir) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > DIR_SIZE) return;
    uint64_t lo = 0, hi = count;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if ((uint64_t)(hashes[mid] >> DIR_SHIFT) < b) lo = mid + 1;
        else hi = mid;
    }
    dir[b] = (uint32_t)lo;
}

// number of varint terminators (bytes with the top bit clear) per granule
__global__ void k3_block_terms(const uint8_t *values, uint64_t nbytes, uint64_t nblocks, uint32_t *terms) {
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const uint64_t p0 = b << SKIP_SHIFT, p1 = min(nbytes, p0 + SKIP_BYTES);
    uint32_t c = 0;
    for (uint64_t p = p0; p < p1; p++) c += (values[p] & 0x80u) ? 0u : 1u;
    terms[b] = c;
}

__device__ __forceinline__ uint64_t terms_before(const uint8_t *values, const uint64_t *block_prefix, uint64_t p) {
    const uint64_t b = p >> SKIP_SHIFT;
    uint64_t c = block_prefix[b];
    for (uint64_t q = b << SKIP_SHIFT; q < p; q++) c += (values[q] & 0x80u) ? 0u : 1u;
    return c;
}

__global__ void k3_list_counts(const uint8_t *values, const uint64_t *offsets, const uint64_t *block_prefix,
                               uint64_t count, uint32_t *counts) {
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= count) return;
    counts[l] = (uint32_t)(terms_before(values, block_prefix, offsets[l + 1]) -
                           terms_before(values, block_prefix, offsets[l]));
}

// largest l with offsets[l] <= p  (offsets strictly increasing: no empty lists in a sparse index)
__device__ __forceinline__ uint64_t list_containing(const uint64_t *offsets, uint64_t count, uint64_t p) {
    uint64_t lo = 0, hi = count; // invariant offsets[lo] <= p < offsets[hi]
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= p) lo = mid;
        else hi = mid;
    }
    return lo;
}

// first varint start >= p inside a list that begins before p
__device__ __forceinline__ uint64_t varint_start_at_or_after(const uint8_t *values, uint64_t p) {
    while (values[p - 1] & 0x80u) p++;
    return p;
}

// For every granule boundary b: skip_off[b] and a (reset flag, partial id sum) pair whose segmented inclusive
// scan is the running structure id at the first varint start at or after the boundary.
__global__ void k3_skip_partials(const uint8_t *values, const uint64_t *offsets, uint64_t count, uint64_t nbytes,
                                 uint64_t nblocks, uint8_t *skip_off, uint64_t *partial) {
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const uint64_t p0 = b << SKIP_SHIFT;
    if (p0 >= nbytes) { // pad entry
        skip_off[b] = 0;
        partial[b] = 1ull << 32;
        return;
    }
    const uint64_t l = list_containing(offsets, count, p0);
    const uint64_t lstart = offsets[l];
    if (lstart == p0) {
        skip_off[b] = 0;
        partial[b] = 1ull << 32; // reset, running id 0
        return;
    }
    const uint64_t pos = varint_start_at_or_after(values, p0);
    skip_off[b] = (uint8_t)(pos - p0);
    uint64_t from;
    uint64_t flag;
    const uint64_t prev0 = p0 - SKIP_BYTES; // b >= 1 here because lstart < p0
    if (lstart > prev0) {
        from = lstart;
        flag = 1ull << 32;
    } else if (lstart == prev0) {
        from = lstart;
        flag = 0; // previous boundary is the list start: its running id is 0 and it carries the reset
    } else {
        from = varint_start_at_or_after(values, prev0);
        flag = 0;
    }
    uint32_t sum = 0, cur = 0;
    int shift = 0;
    for (uint64_t p = from; p < pos; p++) {
        const uint32_t v = values[p];
        cur |= (v & 0x7Fu) << shift;
        if (v & 0x80u) {
            shift += 7;
        } else {
            sum += cur;
            cur = 0;
            shift = 0;
        }
    }
    partial[b] = flag | sum;
}

struct SegScanOp {
    __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const {
        if (b >> 32) return b;
        return (a & ~0xffffffffull) | (uint32_t)((uint32_t)a + (uint32_t)b);
    }
};

__global__ void k3_low32(const uint64_t *in, uint64_t n, uint32_t *out) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (uint32_t)in[k];
}
__global__ void k3_widen32(const uint32_t *in, uint64_t n, uint64_t *out) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[k];
}

// ------------------------------------------------------------------------------------------------
// per-batch kernels
// ------------------------------------------------------------------------------------------------

struct IndexView {
    const uint32_t *hashes;
    const uint64_t *offsets;
    const uint8_t *values;
    const uint32_t *counts;
    const uint32_t *dir;
    const uint8_t *skip_off;
    const uint32_t *skip_id;
    const uint32_t *nres;
    const float *plddt;
    uint64_t count;
    uint32_t n_structs;
};

__device__ __forceinline__ int64_t find_list(const IndexView &ix, uint32_t h) {
    uint32_t lo = ix.dir[h >> DIR_SHIFT], hi = ix.dir[(h >> DIR_SHIFT) + 1];
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        uint32_t v = ix.hashes[mid];
        if (v < h) lo = mid + 1;
        else hi = mid;
    }
    if (lo < ix.count && ix.hashes[lo] == h) return lo;
    return -1;
}

struct QHash {         // one per query hash, produced by k3_lookup
    uint64_t start;    // byte range of the posting list (start == end: absent or filtered)
    uint64_t end;
    uint32_t count;    // postings in the list
    float idf;         // log2(N / count)  (count_query.rs:130)
};

struct QueryDesc {       // one per query (device copy)
    uint32_t hash_begin; // range in the flattened per-hash arrays
    uint32_t n_hashes;
    uint32_t edge_begin; // range in the flattened edge_node array
    uint32_t n_edges;
    uint32_t n_nodes;
    uint32_t expected_node_count;
    uint32_t group_iters; // (largest number of vote bits that stand for one query edge) - 1; 0 = no groups
};

// gcounts / g_structs (id-range shards, fd_count_query_batch_ex): the list length and the structure count of the
// WHOLE database, so that every shard weighs a hash exactly as the unsharded index does (count_query.rs:130); the
// byte range and QHash::count (pool sizing) stay those of the local list, which may be absent.
__global__ void k3_lookup(IndexView ix, const uint32_t *qhashes, uint32_t n_qhashes, float freq_filter,
                          const uint32_t *gcounts, uint32_t g_structs, QHash *out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_qhashes) return;
    QHash r{0, 0, 0, 0.f};
    const int64_t l = find_list(ix, qhashes[k]);
    const uint32_t c_local = l >= 0 ? ix.counts[l] : 0u;
    const uint32_t c = gcounts ? gcounts[k] : c_local;
    const uint32_t n_total = gcounts ? g_structs : ix.n_structs;
    bool keep = c > 0;
    // count_query.rs:124-128: skip hashes more frequent than freq_filter
    if (keep && freq_filter >= 0.f && (float)c / (float)n_total > freq_filter) keep = false;
    if (keep) {
        if (c_local > 0) {
            r.start = ix.offsets[l];
            r.end = ix.offsets[l + 1];
            r.count = c_local;
        }
        r.idf = log2f((float)n_total / (float)c);
    }
    out[k] = r;
}

// per-query totals in a fixed order (deterministic): postings, posting bytes, bound of the idf sum.
// bound_mode: n_hashes * log2(N) -- the same on every rank of a sharded index, whatever the shard holds.
__global__ void k3_query_sums(const QueryDesc *queries, const QHash *qh, uint32_t n_queries, uint32_t n_structs,
                              int bound_mode, unsigned long long *postings, unsigned long long *bytes,
                              float *idf_sum) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const QueryDesc d = queries[q];
    unsigned long long p = 0, b = 0;
    float s = 0.f;
    for (uint32_t k = 0; k < d.n_hashes; k++) {
        const QHash h = qh[d.hash_begin + k];
        p += h.count;
        b += h.end - h.start;
        s += fmaxf(h.idf, 0.f);
    }
    postings[q] = p;
    bytes[q] = b;
    idf_sum[q] = bound_mode ? (float)d.n_hashes * log2f((float)max(n_structs, 2u)) : s;
}

__global__ void k3_counts_only(IndexView ix, const uint32_t *qhashes, uint64_t n, uint32_t *out) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t l = find_list(ix, qhashes[k]);
    out[k] = l >= 0 ? ix.counts[l] : 0u;
}

__global__ void k3_decode_list(IndexView ix, uint32_t hash, uint64_t *out, unsigned long long *n_out, uint64_t cap) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int64_t l = find_list(ix, hash);
    unsigned long long n = 0;
    if (l >= 0) {
        uint64_t id = 0;
        uint32_t cur = 0;
        int shift = 0;
        for (uint64_t p = ix.offsets[l]; p < ix.offsets[l + 1]; p++) {
            const uint32_t v = ix.values[p];
            cur |= (v & 0x7Fu) << shift;
            if (v & 0x80u) {
                shift += 7;
            } else {
                id += cur;
                if (n < cap) out[n] = id;
                n++;
                cur = 0;
                shift = 0;
            }
        }
    }
    *n_out = n;
}

// pen[nid] = (nres as f32).powf(-lp)  (count_query.rs:199), once per batch instead of once per (query, structure)
__global__ void k3_length_penalty(const uint32_t *nres, uint32_t n, float lp, float *pen) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) pen[k] = powf((float)nres[k], -lp);
}

// the same table with the static structure filters folded into the sign: negative where the structure fails
// --num-residue or --plddt (filter.rs:92-98), so that k3_scan_v3 needs one load per non-empty cell
__global__ void k3_length_penalty_signed(const uint32_t *nres, const float *plddt, uint32_t n, float lp,
                                         uint32_t num_res_cutoff, float plddt_cutoff, float *pen) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t nr = nres[k];
    bool pass = true;
    if (num_res_cutoff > 0) pass = pass && nr <= num_res_cutoff;
    if (plddt_cutoff > 0.f) pass = pass && plddt[k] >= plddt_cutoff;
    const float p = powf((float)nr, -lp);
    pen[k] = pass ? p : -p;
}

struct HitRec { // 16 bytes; key/value for the segmented sort are derived from it
    uint32_t nid;
    uint32_t match_count;
    uint32_t node_edge; // node_count << 16 | edge_count
    float idf;
};

struct SparseOut { // MODE 2 of k3_scan
    const uint32_t *slice_begin;        // [world + 1] first query of every rank's slice
    const uint64_t *region_offset;      // [world] first record of every destination's region in the pool
    unsigned long long *region_count;   // [world] records appended so far
    uint32_t world;
    uint32_t rec_planes;                // plane words per record (>= the launch's own planes; the rest is zero)
};

struct FilterParams {
    float length_penalty;
    uint32_t total_match_count, covered_node_count;
    float covered_node_ratio, idf_score_cutoff;
    uint32_t num_res_cutoff;
    float plddt_cutoff;
};

// idf fixed-point scale: the sum over all of the query's hashes must fit the accumulator
template <bool NARROW>
__device__ __forceinline__ float idf_scale(float idf_total) {
    return exp2f(floorf(log2f((NARROW ? 16777215.0f : 4294967040.0f) / (idf_total + 1.0f))));
}

// terminator bits (top bit clear) of the four bytes of w as a 4-bit mask, bit i = byte i
__device__ __forceinline__ uint32_t term_mask4(uint32_t w) {
    return ((((~w) >> 7) & 0x01010101u) * 0x01020408u) >> 24;
}

// LEB128 value starting at the low byte of x0 (4 bytes) with the fifth byte in the low bits of x1
__device__ __forceinline__ uint32_t varint_at(uint32_t x0, uint32_t x1) {
    uint32_t v = x0 & 0x7Fu;
    if (x0 & 0x80u) {
        v |= (x0 >> 1) & 0x3F80u;
        if (x0 & 0x8000u) {
            v |= (x0 >> 2) & 0x1FC000u;
            if (x0 & 0x800000u) {
                v |= (x0 >> 3) & 0xFE00000u;
                if (x0 & 0x80000000u) v |= x1 << 28;
            }
        }
    }
    return v;
}

template <bool NARROW, class F>
__device__ __forceinline__ void compact_cells(const uint32_t *acc, const uint32_t *match, uint32_t T, uint32_t *wq,
                                              F &process);

// Count/filter/append epilogue over a range of vote cells [0, T) (ids lo .. lo+T).  acc / match / edge are the
// planes (shared or global memory); edge plane i is edge + i * edge_stride.  Non-empty cells are compacted warp
// by warp through a small shared-memory queue so that the per-cell work runs with full lanes.
template <bool NARROW, int EW>
__device__ __forceinline__ void emit_cells(const uint32_t *acc, const uint32_t *match, const uint32_t *edge,
                                           size_t edge_stride, uint32_t T, uint32_t lo, const QueryDesc &qd,
                                           const uint32_t *node_mask, const uint32_t *cont_mask /* [EW] */,
                                           uint32_t *wqueue /* [K3_WARPS * K3_WQ] */,
                                           float inv_scale, const IndexView &ix, const float *pen,
                                           const FilterParams &fp, unsigned int *hit_count, HitRec *hits_out) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wq = wqueue + warp * K3_WQ;
    auto process = [&](uint32_t first, uint32_t n_take) { // queue entries [first, first + n_take), one per lane
        bool pass = false;
        HitRec rec{0, 0, 0, 0.f};
        if (lane < n_take) {
            const uint32_t x = wq[first + lane];
            const uint32_t a = acc[x];
            const uint32_t mc = NARROW ? (a >> 24) : match[x];
            const uint32_t fixed = NARROW ? (a & 0xffffffu) : a;
            uint32_t ew[EW];
            uint32_t ec = 0;
#pragma unroll
            for (int i = 0; i < EW; i++) {
                ew[i] = edge[(size_t)i * edge_stride + x];
                // bits of one group (fd_query.edge_group) count once: smear every set bit down to the first
                // bit of its group, then count first bits
                uint32_t g = ew[i];
                const uint32_t cm = cont_mask[i];
                for (uint32_t it = 0; it < qd.group_iters; it++) g |= (g & cm) >> 1;
                ec += __popc(g & ~cm);
            }
            uint32_t nc = 0;
            for (uint32_t nd = 0; nd < qd.n_nodes; nd++) {
                uint32_t any = 0;
#pragma unroll
                for (int i = 0; i < EW; i++) any |= ew[i] & node_mask[nd * EW + i];
                nc += any != 0;
            }
            const uint32_t nid = lo + x;
            const uint32_t nr = ix.nres[nid];
            const float idf = ((float)fixed * inv_scale) * pen[nid];
            pass = true; // filter.rs:76-100
            if (fp.total_match_count > 0) pass = pass && mc >= fp.total_match_count;
            if (fp.covered_node_count > 0) pass = pass && nc >= fp.covered_node_count;
            if (fp.covered_node_ratio > 0.f)
                pass = pass && (float)nc / (float)qd.expected_node_count >= fp.covered_node_ratio;
            if (fp.idf_score_cutoff > 0.f) pass = pass && idf >= fp.idf_score_cutoff;
            if (fp.num_res_cutoff > 0) pass = pass && nr <= fp.num_res_cutoff;
            if (fp.plddt_cutoff > 0.f) pass = pass && ix.plddt[nid] >= fp.plddt_cutoff;
            rec = HitRec{nid, mc, (nc << 16) | (ec & 0xffffu), idf};
        }
        const uint32_t m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(hit_count, (unsigned int)__popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1));
            if (pass) hits_out[pos] = rec;
        }
    };
    compact_cells<NARROW>(acc, match, T, wq, process);
}

// Warp-level compaction of the non-empty vote cells of [0, T): every warp scans 128 cells per step (4 coalesced
// loads per lane), a warp prefix sum places the non-empty ones in the warp's queue wq[K3_WQ], and process(first, n)
// is called warp-wide for queue entries [first, first + n), n <= 32, one per lane.
template <bool NARROW, class F>
__device__ __forceinline__ void compact_cells(const uint32_t *acc, const uint32_t *match, uint32_t T, uint32_t *wq,
                                              F &process) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t pending = 0;
    const uint32_t n_chunks = (T + 127) >> 7;
    for (uint32_t c = warp; c < n_chunks; c += (blockDim.x >> 5)) {
        const uint32_t x0 = (c << 7) + lane;
        uint32_t nz = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t x = x0 + 32 * j;
            uint32_t v = 0;
            if (x < T) v = NARROW ? (acc[x] >> 24) : match[x];
            nz |= (v != 0 ? 1u : 0u) << j;
        }
        const uint32_t cnt = __popc(nz);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        uint32_t pos = pending + incl - cnt;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((nz >> j) & 1u) wq[pos++] = x0 + 32 * j;
        pending += __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        while (pending >= 32) { // take from the back: no queue shifting, hit order is irrelevant (sorted later)
            process(pending - 32, 32);
            pending -= 32;
        }
        __syncwarp();
    }
    if (pending) process(0, pending);
}

// node_mask[nd * EW + w]: vote bits whose source node is nd; cont_mask[w]: vote bits that continue the group of
// the previous bit.  Caller zeroes both and synchronises before and after.
template <int EW>
__device__ __forceinline__ void build_edge_masks(const QueryDesc &qd, const uint16_t *edge_node,
                                                 const uint16_t *edge_group, uint32_t *node_mask,
                                                 uint32_t *cont_mask) {
    for (uint32_t e = threadIdx.x; e < qd.n_edges; e += blockDim.x) {
        atomicOr(&node_mask[edge_node[qd.edge_begin + e] * EW + (e >> 5)], 1u << (e & 31));
        if (qd.group_iters && (e & 31) && edge_group[qd.edge_begin + e] == edge_group[qd.edge_begin + e - 1])
            atomicOr(&cont_mask[e >> 5], 1u << (e & 31));
    }
}

// Shared-memory vote tile.  NARROW: plane 0 = match_count (8 bits) | idf fixed point (24 bits);
// wide: separate idf and match planes.  EW edge-bitmask planes.  MODE 0: count/filter/append epilogue;
// 1: write the planes to the dense partial-vote buffer; 2: pack the non-empty cells as sparse records for the rank
// that finishes the query (multi-GPU).
// partial-vote buffer instead of running the epilogue (multi-GPU).
template <bool NARROW, int EW, int MODE>
__global__ void __launch_bounds__(K3_MAX_THREADS)
    k3_scan(IndexView ix, const QueryDesc *queries, const QHash *qh, const uint16_t *edge_of_hash,
            const uint16_t *edge_node, const uint16_t *edge_group, const float *idf_sum_per_query, const float *pen, uint32_t tile_ids,
            FilterParams fp, const uint64_t *hit_offsets, unsigned int *hit_counts, HitRec *hits,
            uint32_t *dense /* MODE 1: [planes][nq][N]; MODE 2: sparse record pool */, SparseOut sp,
            int stage_lists /* per-list byte range / vote word / edge bit staged in shared memory */,
            const uint32_t *qlist /* MODE 0: the queries of this launch (one launch per edge-word class), or NULL */) {
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t q = qlist ? qlist[blockIdx.y] : blockIdx.y;
    const QueryDesc qd = queries[q];
    const uint32_t lo = blockIdx.x * tile_ids;
    const uint32_t hi = min(ix.n_structs, lo + tile_ids);
    const uint32_t T = hi - lo;
    const bool single_tile = gridDim.x == 1;
    const uint32_t Q = qd.n_hashes;
    constexpr uint32_t PLANES = (NARROW ? 1 : 2) + EW;

    // smem carve-up (tile_ids is a multiple of 32, so every plane is 16-byte aligned)
    uint32_t *w_acc = smem;                                  // [tile_ids] narrow: packed; wide: idf fixed
    uint32_t *w_match = NARROW ? nullptr : w_acc + tile_ids; // [tile_ids] wide only
    uint32_t *w_edge = w_acc + (NARROW ? 1 : 2) * tile_ids;  // EW planes of [tile_ids]
    uint64_t *l_range = reinterpret_cast<uint64_t *>(w_acc + (size_t)PLANES * tile_ids); // [2 Q] start, end (staged)
    uint32_t *l_vote = reinterpret_cast<uint32_t *>(l_range + (stage_lists ? 2 * Q : 0)); // [2 Q] vote word, edge
    uint32_t *item_prefix = l_vote + (stage_lists ? 2 * Q : 0); // [Q + 1]
    uint32_t *seg_lo = item_prefix + (Q + 1);                // [Q] first relevant granule of each list
    uint32_t *node_mask = seg_lo + Q;                        // [n_nodes * EW]
    uint32_t *cont_mask = node_mask + qd.n_nodes * EW;       // [EW]
    uint32_t *wqueue = cont_mask + EW;                       // [K3_WARPS * K3_WQ]
    __shared__ uint32_t s_total_items;

    const float scale = idf_scale<NARROW>(idf_sum_per_query[q]);
    const float inv_scale = 1.0f / scale;
    for (uint32_t i = threadIdx.x; i < qd.n_nodes * EW + EW; i += blockDim.x) node_mask[i] = 0; // + cont_mask

    // ---- which 64-byte granules of each list intersect this tile ----
    for (uint32_t k = threadIdx.x; k < Q; k += blockDim.x) {
        const QHash h = qh[qd.hash_begin + k];
        uint32_t n_items = 0, first = 0;
        if (h.end > h.start) {
            const uint64_t b0 = h.start >> SKIP_SHIFT, b1 = (h.end - 1) >> SKIP_SHIFT;
            const uint32_t nseg = (uint32_t)(b1 - b0) + 1;
            if (single_tile || nseg == 1) {
                first = 0;
                n_items = nseg;
            } else {
                // r(k') = skip_id[b0 + k'] for k' in [1, nseg-1], non-decreasing: the id of the last posting
                // that starts before granule k'
                const uint32_t *r = ix.skip_id + b0;
                uint32_t a = 1, b = nseg; // first k' with r(k') >= lo
                while (a < b) {
                    uint32_t m = (a + b) >> 1;
                    if (r[m] < lo) a = m + 1;
                    else b = m;
                }
                const uint32_t below_lo = a - 1;
                a = 1;
                b = nseg;
                while (a < b) {
                    uint32_t m = (a + b) >> 1;
                    if (r[m] < hi) a = m + 1;
                    else b = m;
                }
                const uint32_t below_hi = a - 1;
                first = below_lo;                  // granule below_lo may still hold ids >= lo
                n_items = below_hi + 1 - below_lo; // granules [below_lo, below_hi]
            }
        }
        seg_lo[k] = first;
        item_prefix[k + 1] = n_items;
        if (stage_lists) {
            l_range[2 * k] = h.start;
            l_range[2 * k + 1] = h.end;
            const uint32_t wgt = (uint32_t)(fmaxf(h.idf, 0.f) * scale + 0.5f);
            l_vote[2 * k] = NARROW ? ((1u << 24) | wgt) : wgt;
            l_vote[2 * k + 1] = edge_of_hash[qd.hash_begin + k];
        }
    }
    if (threadIdx.x == 0) item_prefix[0] = 0;
    __syncthreads();
    build_edge_masks<EW>(qd, edge_node, edge_group, node_mask, cont_mask);
    // inclusive scan of item_prefix[1..Q] (Q is small; one warp, chunked)
    if (threadIdx.x < 32) {
        uint32_t carry = 0;
        for (uint32_t base = 1; base <= Q; base += 32) {
            uint32_t idx = base + threadIdx.x;
            uint32_t v = idx <= Q ? item_prefix[idx] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
                if ((int)threadIdx.x >= o) v += t;
            }
            v += carry;
            if (idx <= Q) item_prefix[idx] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
        if (threadIdx.x == 0) s_total_items = carry;
    }
    __syncthreads();
    const uint32_t total_items = s_total_items;
    // nothing of this query's lists falls into this tile (common for hash-range shards: a rank holds the lists of
    // a few amino-acid pairs only): no votes, no hits, no records -- the tile is neither cleared nor scanned
    if (MODE != 1 && total_items == 0) return;
    {
        uint4 *z = reinterpret_cast<uint4 *>(smem);
        const uint32_t n4 = (PLANES * tile_ids) >> 2;
        for (uint32_t i = threadIdx.x; i < n4; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    // ---- decode + vote: 8 lanes per granule, 32 granules per CTA step ----
    const uint32_t sub = threadIdx.x & 7, grp = threadIdx.x >> 3;
    for (uint32_t it0 = 0; it0 < total_items; it0 += (blockDim.x >> 3)) {
        if (it0 + ((threadIdx.x >> 5) << 2) >= total_items) break; // none of this warp's four granules exists
        const uint32_t it = it0 + grp;
        const bool act = it < total_items;
        uint32_t w0 = 0, w1 = 0, w2 = 0, base = 0, add = 0, ebit = 0, eword = 0;
        uint64_t A = 0, start_pos = 0, lim = 0;
        if (act) {
            uint32_t a = 0, b = Q; // list index: last k with item_prefix[k] <= it
            while (b - a > 1) {
                uint32_t m = (a + b) >> 1;
                if (item_prefix[m] <= it) a = m;
                else b = m;
            }
            const uint32_t k = a;
            uint64_t h_start, h_end;
            uint32_t e;
            if (stage_lists) {
                h_start = l_range[2 * k];
                h_end = l_range[2 * k + 1];
                add = l_vote[2 * k];
                e = l_vote[2 * k + 1];
            } else {
                const QHash h = qh[qd.hash_begin + k];
                h_start = h.start;
                h_end = h.end;
                e = edge_of_hash[qd.hash_begin + k];
                const uint32_t wgt = (uint32_t)(fmaxf(h.idf, 0.f) * scale + 0.5f);
                add = NARROW ? ((1u << 24) | wgt) : wgt;
            }
            const uint32_t seg = seg_lo[k] + (it - item_prefix[k]);
            const uint64_t gi = (h_start >> SKIP_SHIFT) + seg;
            const uint64_t G0 = gi << SKIP_SHIFT;
            if (seg == 0) {
                start_pos = h_start;
            } else {
                start_pos = G0 + ix.skip_off[gi];
                base = ix.skip_id[gi];
            }
            lim = min(h_end, G0 + SKIP_BYTES);
            A = G0 + 8u * sub;
            const uint2 w = *reinterpret_cast<const uint2 *>(ix.values + A);
            w0 = w.x;
            w1 = w.y;
            if (sub == 7) w2 = *reinterpret_cast<const uint32_t *>(ix.values + G0 + SKIP_BYTES);
            ebit = 1u << (e & 31);
            eword = e >> 5;
        }
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, w0, 1, 8);
        if (sub != 7) w2 = nxt;
        const uint32_t tm8 = term_mask4(w0) | (term_mask4(w1) << 4);
        uint32_t prevterm = (__shfl_up_sync(0xffffffffu, tm8, 1, 8) >> 7) & 1u;
        if (sub == 0) prevterm = 0;
        uint32_t sm = ((tm8 << 1) | prevterm) & 0xffu; // byte i starts a varint iff byte i-1 terminates one
        {
            const uint32_t v0 = start_pos > A ? (uint32_t)min(start_pos - A, (uint64_t)8) : 0u;
            const uint32_t v1 = lim > A ? (uint32_t)min(lim - A, (uint64_t)8) : 0u;
            const uint32_t valid = v1 > v0 ? (((1u << v1) - 1u) & ~((1u << v0) - 1u)) : 0u;
            sm &= valid;
            if (start_pos >= A && start_pos < A + 8 && start_pos < lim) sm |= 1u << (uint32_t)(start_pos - A);
            if (!act) sm = 0;
        }
        // running sums of the deltas of the varints that start in this lane's bytes
        uint32_t ds[8];
        uint32_t run = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if ((sm >> i) & 1u) {
                const uint32_t x0 = i == 0 ? w0
                                    : i < 4 ? __funnelshift_r(w0, w1, 8 * i)
                                    : i == 4 ? w1
                                             : __funnelshift_r(w1, w2, 8 * (i - 4));
                const uint32_t x1 = i < 4 ? (w1 >> (8 * i)) : (w2 >> (8 * (i - 4)));
                run += varint_at(x0, x1 & 0xffu);
            }
            ds[i] = run;
        }
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o, 8);
            if ((int)sub >= o) inc += t;
        }
        const uint32_t lane_base = base + inc - run;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if ((sm >> i) & 1u) {
                const uint32_t id = lane_base + ds[i];
                if (id >= lo && id < hi) {
                    const uint32_t x = id - lo;
                    atomicAdd(&w_acc[x], add);
                    if (!NARROW) atomicAdd(&w_match[x], 1u);
                    atomicOr(&w_edge[eword * tile_ids + x], ebit);
                }
            }
        }
    }
    __syncthreads();

    if (MODE == 2) {
        // records {key = (q - first query of the destination) * N + nid, planes...} appended to the region of the
        // rank that finishes q
        __shared__ uint32_t s_dest;
        if (threadIdx.x == 0) {
            uint32_t d = 0;
            while (d + 1 < sp.world && q >= sp.slice_begin[d + 1]) d++;
            s_dest = d;
        }
        __syncthreads();
        const uint32_t dest = s_dest;
        const uint32_t key0 = (q - sp.slice_begin[dest]) * ix.n_structs + lo;
        const uint32_t RP = max(sp.rec_planes, PLANES); // record width of the batch; this launch's class may be narrower
        uint32_t *region = dense + sp.region_offset[dest] * (1 + RP);
        unsigned long long *counter = sp.region_count + dest;
        const uint32_t lane = threadIdx.x & 31;
        uint32_t *wq = wqueue + (threadIdx.x >> 5) * K3_WQ;
        auto pack = [&](uint32_t first, uint32_t n_take) {
            unsigned long long pos = 0;
            if (lane == 0) pos = atomicAdd(counter, (unsigned long long)n_take);
            pos = __shfl_sync(0xffffffffu, pos, 0) + lane;
            if (lane < n_take) {
                const uint32_t x = wq[first + lane];
                uint32_t *r = region + pos * (1 + RP);
                r[0] = key0 + x;
#pragma unroll
                for (uint32_t p = 0; p < PLANES; p++) r[1 + p] = w_acc[(size_t)p * tile_ids + x];
                for (uint32_t p = PLANES; p < RP; p++) r[1 + p] = 0;
            }
        };
        compact_cells<NARROW>(w_acc, w_match, T, wq, pack);
    } else if (MODE == 1) {
        // planes of this tile -> dense[plane][q][lo .. hi)
        const size_t plane_stride = (size_t)gridDim.y * ix.n_structs;
        uint32_t *dst = dense + (size_t)q * ix.n_structs + lo;
        for (uint32_t p = 0; p < PLANES; p++)
            for (uint32_t x = threadIdx.x; x < T; x += blockDim.x)
                dst[p * plane_stride + x] = w_acc[(size_t)p * tile_ids + x];
    } else {
        emit_cells<NARROW, EW>(w_acc, w_match, w_edge, tile_ids, T, lo, qd, node_mask, cont_mask, wqueue, inv_scale, ix,
                               pen, fp, &hit_counts[q], hits + hit_offsets[q]);
    }
}

// ------------------------------------------------------------------------------------------------
// k3_scan_v3: the single-rank / id-range-shard scan (fd_count_query_batch[_ex], fd_count_query_sharded).
//
// A work item is (query, run of consecutive id tiles); persistent CTAs (two per SM) take items, costliest query
// first, from an atomic counter.  What changed against k3_scan_v2 and why (profiles/r03a_ncu_full.txt: 70 k warp
// instructions per 6 k postings, 55 % of the stall samples at CTA barriers, issue slots 37 % busy):
//   * the query's list descriptors are staged ONCE per item and the tiles of the run are walked in ascending id
//     order, so the granule range of a list in the next tile is found by a short galloping search from the current
//     position (one dependent skip-table load or two) instead of two full binary searches per (list, tile);
//   * the epilogue no longer compacts the voted cells through queues and global scratch lists: two dense passes over
//     the tile, four cells per thread (one LDS.128 of the occupancy plane + one coalesced LDG.128 of the signed
//     length-penalty table): pass A bins the idf of the passing cells into the item's histogram, pass B emits the
//     cells at or above the threshold bin and zeroes the planes with vector stores;
//   * the histogram is CUMULATIVE over the tiles of the run: the threshold bin only rises, so later tiles emit
//     fewer and fewer hits (still a superset of the query's top n: a cell of the global top n has fewer than n better
//     cells in ANY subset of the structures);
//   * pipelining across the barriers: the search for tile t + 1 runs in front of pass A of tile t, the scan of its
//     granule counts beside the threshold search, and the bulk copies (cp.async.bulk -> per-warp stage, completion on
//     the warp's mbarrier) of its first decode step are issued before pass B, so they land while the tile is emitted.
// Posting bytes are decoded one 64-byte granule per lane (fully unrolled LEB128 walk); votes are fire-and-forget
// shared-memory atomics into the tile's planes.
// ------------------------------------------------------------------------------------------------
constexpr int K3V_MAX_THREADS = 512;              // two co-resident CTAs per SM at <= 64 registers per thread
constexpr uint32_t K3V_HIST_BINS = 2048;           // idf bin = float bits >> 20 (sign-less): 8 exponent + 3 mantissa bits
constexpr uint32_t K3V_GRANULE_STAGE = 80;         // 64-byte granule + 16 bytes of look-ahead (a varint may spill 4 bytes)
constexpr uint32_t K3V_WARP_STAGE = 32 * K3V_GRANULE_STAGE; // one granule per lane per warp step
constexpr int K3V_DECODE_WARPS = 8;                // warps of a CTA that decode (each owns a stage); all warps run the epilogue

struct ScanItem {
    uint32_t q, tile_begin, tile_end;
};

struct ScanV3Args {
    const QueryDesc *queries;
    const QHash *qh;
    const uint16_t *edge_of_hash, *edge_node, *edge_group;
    const float *idf_sum;
    const float *pen; // nres^-lp, negative where the structure fails --num-residue / --plddt; padded to a multiple of 4
    const ScanItem *items;
    uint32_t n_items;
    unsigned int *item_counter;
    uint32_t tile_words;                  // u32 words of vote planes per CTA (multiple of 4)
    uint32_t max_hashes, max_nodes, max_ew; // sizes of the fixed shared-memory arrays
    uint32_t stage_lists;                 // list descriptors staged in shared memory (else read from a.qh)
    uint32_t limit_n;                     // > 0: keep only hits that can reach the run's top limit_n
    FilterParams fp;
    const uint64_t *hit_offsets;          // [nq + 1] hit regions
    unsigned int *hit_counts;
    HitRec *hits;
    unsigned int *overflow;               // set when a region was too small (host re-runs unlimited)
    uint32_t use_bulk;                    // posting bytes staged by bulk copies (else: plain vector loads; A/B switch)
    uint32_t n_dwarps;                    // warps of a CTA that decode (each owns a stage); all warps run the epilogue
};

__host__ __device__ __forceinline__ uint32_t k3v_edge_words(uint32_t n_edges) {
    return n_edges <= 32 ? 1u : (n_edges + 31u) >> 5;
}
__host__ __device__ __forceinline__ uint32_t k3v_tile_ids(uint32_t tile_words, uint32_t planes, uint32_t n_structs) {
    uint32_t t = (tile_words / planes) & ~31u;
    const uint32_t n32 = (n_structs + 31u) & ~31u;
    return t > n32 ? n32 : t;
}
__device__ __forceinline__ uint32_t k3v_idf_bin(float idf) {
    const uint32_t u = __float_as_uint(idf);
    return (u & 0x80000000u) ? 0u : (u >> 20);
}
__device__ __forceinline__ uint32_t u4_get(const uint4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
__device__ __forceinline__ float f4_get(const float4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

template <bool NARROW>
__global__ void __launch_bounds__(K3V_MAX_THREADS, 2) k3_scan_v3(IndexView ix, ScanV3Args a) {
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const uint32_t MQ = a.max_hashes;
    // ---- shared-memory carve-up ----
    uint32_t *planes = smem;                                                            // [tile_words]
    uint64_t *l_range = reinterpret_cast<uint64_t *>(planes + a.tile_words);            // [2 MQ] start, end (staged)
    uint32_t *l_vote = reinterpret_cast<uint32_t *>(l_range + (a.stage_lists ? 2 * MQ : 0)); // [2 MQ] vote word, edge
    uint32_t *item_prefix = l_vote + (a.stage_lists ? 2 * MQ : 0);                      // [MQ + 1] granules of the tile, scanned
    uint32_t *seg_lo = item_prefix + (MQ + 1);                                          // [MQ] first granule of the tile
    uint32_t *b_next = seg_lo + MQ;                                                     // [MQ] first granule of the next tile
    uint32_t *node_mask = b_next + MQ;                                                  // [max_nodes * max_ew + max_ew]
    uint32_t *hist = node_mask + a.max_nodes * a.max_ew + a.max_ew;                     // [K3V_HIST_BINS]
    uint32_t *wqueue = hist + K3V_HIST_BINS;                                            // [nwarps * K3_WQ] candidate cells
    uintptr_t sp = reinterpret_cast<uintptr_t>(wqueue + nwarps * K3_WQ);
    sp = (sp + 15) & ~(uintptr_t)15;
    const uint32_t n_dwarps = min(nwarps, a.n_dwarps);
    uint8_t *stage = reinterpret_cast<uint8_t *>(sp);                                   // [n_dwarps][K3V_WARP_STAGE]
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage + (size_t)n_dwarps * K3V_WARP_STAGE); // [n_dwarps]
    __shared__ uint32_t s_item, s_total[2], s_thr, s_maxbin;

    uint8_t *wstage = stage + (size_t)min(warp, n_dwarps - 1) * K3V_WARP_STAGE;
    uint64_t *wbar = bars + min(warp, n_dwarps - 1);
    uint8_t *myslot = wstage + lane * K3V_GRANULE_STAGE;
    if (lane == 0 && warp < n_dwarps) fda::mbar_init(wbar, 1);
    {
        uint4 *z = reinterpret_cast<uint4 *>(planes);
        for (uint32_t i = tid; i < (a.tile_words >> 2); i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    fda::mbar_fence_init();
    __syncthreads();
    uint32_t phase = 0; // parity of the warp's stage barrier

    for (;;) {
        if (tid == 0) s_item = atomicAdd(a.item_counter, 1u);
        __syncthreads();
        const uint32_t item_no = s_item;
        if (item_no >= a.n_items) break;
        const ScanItem item = a.items[item_no];
        const uint32_t q = item.q;
        const QueryDesc qd = a.queries[q];
        const uint32_t Q = qd.n_hashes;
        const uint32_t EW = k3v_edge_words(qd.n_edges);
        const uint32_t PL = (NARROW ? 1u : 2u) + EW;
        const uint32_t tile_ids = k3v_tile_ids(a.tile_words, PL, ix.n_structs);
        uint32_t *w_acc = planes;
        uint32_t *w_match = NARROW ? nullptr : planes + tile_ids;
        uint32_t *w_edge = planes + (NARROW ? 1 : 2) * tile_ids;
        uint32_t *cont_mask = node_mask + qd.n_nodes * EW;
        // 32-bit shared-memory addresses of the planes for the votes (red.shared)
        const uint32_t acc_sa = fda::smem_u32(w_acc), edge_sa0 = fda::smem_u32(w_edge);
        const uint32_t match_sa = NARROW ? 0u : fda::smem_u32(planes + tile_ids);
        const float scale = idf_scale<NARROW>(a.idf_sum[q]);
        const float inv_scale = 1.0f / scale;
        const uint64_t hbase = a.hit_offsets[q];
        const uint32_t hcap = (uint32_t)(a.hit_offsets[q + 1] - hbase);
        for (uint32_t i = tid; i < qd.n_nodes * EW + EW; i += blockDim.x) node_mask[i] = 0; // + cont_mask
        for (uint32_t i = tid; i < K3V_HIST_BINS; i += blockDim.x) hist[i] = 0;
        if (tid == 0) s_maxbin = 0;
        uint32_t run_thr = 0; // threshold bin of the run so far (CTA-uniform)

        // granules of list k that may hold ids of [from, to): [B(from), B(to)], B(v) = (first granule g >= 1 whose
        // running id skip_id[b0 + g] -- the id of the last posting that starts before g -- is >= v) - 1.
        // `c` = B(from) is known; the search for B(to) gallops from it (tiles are walked in ascending order).
        auto list_range = [&](uint32_t k, uint64_t &h_start, uint64_t &h_end) {
            if (a.stage_lists) {
                h_start = l_range[2 * k];
                h_end = l_range[2 * k + 1];
            } else {
                const QHash h = a.qh[qd.hash_begin + k];
                h_start = h.start;
                h_end = h.end;
            }
        };
        auto advance = [&](uint32_t k, uint32_t c, uint32_t to) -> uint32_t {
            uint64_t h_start, h_end;
            list_range(k, h_start, h_end);
            if (h_end <= h_start) return 0;
            const uint64_t b0 = h_start >> SKIP_SHIFT;
            const uint32_t nseg = (uint32_t)(((h_end - 1) >> SKIP_SHIFT) - b0) + 1;
            if (to >= ix.n_structs) return nseg - 1;
            const uint32_t *r = ix.skip_id + b0;
            uint32_t x = c + 1, y = nseg, probe = c + 1, step = 1;
            while (probe < nseg) {
                if (r[probe] >= to) {
                    y = probe;
                    break;
                }
                x = probe + 1;
                probe += step;
                step <<= 1;
            }
            while (x < y) {
                const uint32_t m = (x + y) >> 1;
                if (r[m] < to) x = m + 1;
                else y = m;
            }
            return x - 1;
        };
        // tile `t`: granule counts of the lists (item_prefix[k + 1], not yet scanned), first granules (seg_lo), and
        // the positions the next tile starts from (b_next); `from_pos` = position array of the tile's lower bound
        auto setup_tile = [&](uint32_t t) {
            const uint32_t hi_t = min(ix.n_structs, (t + 1) * tile_ids);
            for (uint32_t k = tid; k < Q; k += blockDim.x) {
                uint64_t h_start, h_end;
                list_range(k, h_start, h_end);
                const uint32_t c = b_next[k];
                uint32_t n_gr = 0, nx = c;
                if (h_end > h_start) {
                    nx = advance(k, c, hi_t);
                    n_gr = nx - c + 1;
                }
                seg_lo[k] = c;
                b_next[k] = nx;
                item_prefix[k + 1] = n_gr;
            }
        };
        auto scan_prefix = [&](uint32_t slot) { // one warp: inclusive scan of item_prefix[1..Q], total -> s_total[slot]
            uint32_t carry = 0;
            for (uint32_t base = 1; base <= Q; base += 32) {
                const uint32_t idx = base + lane;
                uint32_t v = idx <= Q ? item_prefix[idx] : 0;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t tt = __shfl_up_sync(0xffffffffu, v, o);
                    if ((int)lane >= o) v += tt;
                }
                v += carry;
                if (idx <= Q) item_prefix[idx] = v;
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
            if (lane == 0) {
                item_prefix[0] = 0;
                s_total[slot] = carry;
            }
        };

        // ---- item setup: descriptors, position of every list at the run's first id ----
        const uint32_t lo_first = item.tile_begin * tile_ids;
        for (uint32_t k = tid; k < Q; k += blockDim.x) {
            const QHash h = a.qh[qd.hash_begin + k];
            uint32_t c = 0;
            if (h.end > h.start && lo_first > 0) {
                const uint64_t b0 = h.start >> SKIP_SHIFT;
                const uint32_t nseg = (uint32_t)(((h.end - 1) >> SKIP_SHIFT) - b0) + 1;
                const uint32_t *r = ix.skip_id + b0;
                uint32_t x = 1, y = nseg; // first g with r(g) >= lo_first
                while (x < y) {
                    const uint32_t m = (x + y) >> 1;
                    if (r[m] < lo_first) x = m + 1;
                    else y = m;
                }
                c = x - 1;
            }
            b_next[k] = c;
            if (a.stage_lists) {
                l_range[2 * k] = h.start;
                l_range[2 * k + 1] = h.end;
                const uint32_t wgt = (uint32_t)(fmaxf(h.idf, 0.f) * scale + 0.5f);
                l_vote[2 * k] = NARROW ? ((1u << 24) | wgt) : wgt;
                l_vote[2 * k + 1] = a.edge_of_hash[qd.hash_begin + k];
            }
        }
        __syncthreads();
        for (uint32_t e = tid; e < qd.n_edges; e += blockDim.x) {
            atomicOr(&node_mask[a.edge_node[qd.edge_begin + e] * EW + (e >> 5)], 1u << (e & 31));
            if (qd.group_iters && (e & 31) && a.edge_group[qd.edge_begin + e] == a.edge_group[qd.edge_begin + e - 1])
                atomicOr(&cont_mask[e >> 5], 1u << (e & 31));
        }
        setup_tile(item.tile_begin);
        __syncthreads();
        if (warp == 0) scan_prefix(item.tile_begin & 1u);
        __syncthreads();

        // ---- decode state of this lane: the granule it decodes next (prefetched) ----
        uint32_t pk = 0, pseg = 0, psid = 0, psoff = 0, pn = 0;
        uint64_t pgi = 0;
        bool pact = false;
        const uint32_t gpc = n_dwarps << 5; // granules per CTA step
        auto prefetch = [&](uint32_t s, uint32_t total_items) {
            const uint32_t it = s * gpc + (warp << 5) + lane;
            pact = it < total_items;
            pk = pseg = psid = psoff = 0;
            uint64_t gi = 0;
            if (pact) {
                uint32_t x = 0, y = Q; // list index: last k with item_prefix[k] <= it
                while (y - x > 1) {
                    const uint32_t m = (x + y) >> 1;
                    if (item_prefix[m] <= it) x = m;
                    else y = m;
                }
                pk = x;
                pseg = seg_lo[x] + (it - item_prefix[x]);
                uint64_t h_start, h_end;
                list_range(x, h_start, h_end);
                gi = (h_start >> SKIP_SHIFT) + pseg;
                if (pseg) {
                    psoff = ix.skip_off[gi];
                    psid = ix.skip_id[gi];
                }
            }
            pn = __popc(__ballot_sync(0xffffffffu, pact));
            pgi = gi;
            if (pn && a.use_bulk) {
                if (lane == 0) fda::mbar_arrive_expect_tx(wbar, pn * K3V_GRANULE_STAGE);
                __syncwarp();
                if (pact) fda::bulk_g2s(myslot, ix.values + (gi << SKIP_SHIFT), K3V_GRANULE_STAGE, wbar);
            }
        };
        if (warp < n_dwarps) prefetch(0, s_total[item.tile_begin & 1u]);

        for (uint32_t t = item.tile_begin; t < item.tile_end; t++) {
            const uint32_t lo = t * tile_ids;
            const uint32_t hi = min(ix.n_structs, lo + tile_ids);
            const uint32_t T = hi - lo;
            const uint32_t total_items = s_total[t & 1u];
            const bool has_next = t + 1 < item.tile_end;

            // ---- decode + vote: ONE LANE PER GRANULE ----
            if (warp < n_dwarps && total_items) {
                const uint32_t nsteps = (total_items + gpc - 1) / gpc;
                for (uint32_t s = 0; s < nsteps; s++) {
                    const uint32_t ck = pk, cseg = pseg, csid = psid, csoff = psoff, cn = pn;
                    const bool act = pact;
                    if (cn == 0) break; // the granules are handed out in ascending order: no later step has one for this warp
                    if (a.use_bulk) {
                        fda::mbar_wait(wbar, phase);
                        phase ^= 1u;
                    }
                    uint32_t W[20];
                    {
                        const uint4 *src = a.use_bulk ? reinterpret_cast<const uint4 *>(myslot)
                                                      : reinterpret_cast<const uint4 *>(ix.values + (pgi << SKIP_SHIFT));
#pragma unroll
                        for (int j = 0; j < 5; j++) {
                            const uint4 v = src[j];
                            W[4 * j] = v.x;
                            W[4 * j + 1] = v.y;
                            W[4 * j + 2] = v.z;
                            W[4 * j + 3] = v.w;
                        }
                    }
                    __syncwarp(); // every lane holds its bytes in registers: the stage is free for the next step's copies
                    if (s + 1 < nsteps) prefetch(s + 1, total_items);
                    else pn = 0;
                    uint32_t lp0 = 0xff, lp1 = 0, lid = 0, ladd = 0, lebit = 0; // byte range [lp0, lp1), running id, vote
                    uint32_t edge_sa = edge_sa0;
                    if (act) {
                        uint64_t h_start, h_end;
                        uint32_t add, e;
                        list_range(ck, h_start, h_end);
                        if (a.stage_lists) {
                            add = l_vote[2 * ck];
                            e = l_vote[2 * ck + 1];
                        } else {
                            const QHash h = a.qh[qd.hash_begin + ck];
                            e = a.edge_of_hash[qd.hash_begin + ck];
                            const uint32_t wgt = (uint32_t)(fmaxf(h.idf, 0.f) * scale + 0.5f);
                            add = NARROW ? ((1u << 24) | wgt) : wgt;
                        }
                        const uint64_t G0 = ((h_start >> SKIP_SHIFT) + cseg) << SKIP_SHIFT;
                        // varints that START in [p0, p1) of this granule are this lane's; the last may end in the look-ahead
                        lp0 = cseg == 0 ? (uint32_t)(h_start - G0) : csoff;
                        lp1 = (uint32_t)min(h_end - G0, (uint64_t)SKIP_BYTES);
                        lid = (cseg == 0 ? 0u : csid) - lo; // tile-relative running id (wraps below the tile)
                        ladd = add;
                        lebit = 1u << (e & 31);
                        edge_sa = edge_sa0 + (e >> 5) * tile_ids * 4u;
                    }
                    // The warp walks only the words that hold bytes of some lane's range (short lists fill a fraction of
                    // their granule); a varint that starts before byte 64 may end in the look-ahead word.  The byte step is
                    // written without branches (selects + two predicated reductions on 32-bit shared addresses): every
                    // lane runs the same ~14 instructions per byte whatever its list looks like.
                    const uint32_t wlo = __reduce_min_sync(0xffffffffu, act ? lp0 : 0xffu) >> 2;
                    const uint32_t whi = __reduce_max_sync(0xffffffffu, act ? lp1 : 0u);
                    const uint64_t rmask = act && lp1 > lp0 ? ((lp1 >= 64u ? ~0ull : ((1ull << lp1) - 1ull)) & ~((1ull << lp0) - 1ull)) : 0ull;
                    const uint32_t rm_lo = (uint32_t)rmask, rm_hi = (uint32_t)(rmask >> 32);
                    uint32_t cur = 0, shift = 0;
#pragma unroll
                    for (int w = 0; w < (int)(SKIP_BYTES / 4) + 1; w++) {
                        if (w < (int)(SKIP_BYTES / 4) ? ((uint32_t)w < wlo || (uint32_t)(4 * w) >= whi) : whi < SKIP_BYTES) continue;
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const int p = 4 * w + b;
                            const uint32_t byte = (W[w] >> (8 * b)) & 0xffu;
                            const bool inr = p < 32 ? ((rm_lo >> p) & 1u) != 0u : (p < 64 ? ((rm_hi >> (p - 32)) & 1u) != 0u : false);
                            const bool on = inr || shift != 0u;
                            const bool cont = (byte & 0x80u) != 0u;
                            cur |= on ? ((byte & 0x7fu) << shift) : 0u;
                            const uint32_t nid = lid + cur;
                            const bool fin = on && !cont;
                            if (fin && nid < T) {
                                fda::red_add_shared(acc_sa + nid * 4u, ladd);
                                if (!NARROW) fda::red_add_shared(match_sa + nid * 4u, 1u);
                                fda::red_or_shared(edge_sa + nid * 4u, lebit);
                            }
                            lid = fin ? nid : lid;
                            const bool more = on && cont;
                            shift = more ? shift + 7u : 0u;
                            cur = more ? cur : 0u;
                        }
                    }
                }
            }
            __syncthreads();

            // ---- the next tile's granule ranges (latency of the skip-table loads hides under the epilogue of the others) ----
            if (has_next) setup_tile(t + 1);
            if (!total_items) { // nothing of the query falls into this tile: the planes are clean
                __syncthreads();
                if (has_next && warp == 0) scan_prefix((t + 1) & 1u);
                __syncthreads();
                if (has_next && warp < n_dwarps) prefetch(0, s_total[(t + 1) & 1u]);
                continue;
            }

            // ---- epilogue ----
            // The tile holds far more cells than votes (half a posting per cell on the bench), so the pass over the
            // cells must cost a few instructions per cell: four cells per thread from one LDS.128 of the occupancy
            // plane, and a voted cell is first held against an INTEGER bound -- the smallest fixed-point idf that
            // could reach the run's threshold bin under the largest length penalty of the database -- before its
            // penalty is loaded and its idf evaluated.  Cells that reach the threshold go to a CTA queue and are
            // turned into hit records on full lanes (node / edge counts are ~100 instructions per record).
            // First tile of a run: pass A (idf histogram of every passing cell) -> threshold -> pass B1 (queue) -> B2.
            // Later tiles: ONE pass B1 against the threshold of the tiles before (it only rises), feeding the histogram.
            const uint32_t *occ = NARROW ? w_acc : w_match; // the plane whose word is non-zero exactly in the voted cells
            const uint32_t T4 = (T + 3) >> 2;               // four cells per thread (planes and pen are padded)
            const uint4 *occ4 = reinterpret_cast<const uint4 *>(occ);
            const uint4 *acc4 = reinterpret_cast<const uint4 *>(w_acc);
            const float4 *pen4 = reinterpret_cast<const float4 *>(a.pen + lo);
            auto node_count = [&](uint32_t x) -> uint32_t {
                uint32_t nc = 0;
                if (EW == 1) {
                    const uint32_t ebits = w_edge[x];
                    for (uint32_t nd = 0; nd < qd.n_nodes; nd++) nc += (ebits & node_mask[nd]) != 0;
                } else {
                    for (uint32_t nd = 0; nd < qd.n_nodes; nd++) {
                        uint32_t any = 0;
                        for (uint32_t i = 0; i < EW; i++) any |= w_edge[i * tile_ids + x] & node_mask[nd * EW + i];
                        nc += any != 0;
                    }
                }
                return nc;
            };
            auto edge_count = [&](uint32_t x) -> uint32_t {
                uint32_t ec = 0;
                for (uint32_t i = 0; i < EW; i++) {
                    // bits of one group (fd_query.edge_group) count once: smear every set bit down to the first bit
                    // of its group, then count first bits
                    uint32_t g = w_edge[i * tile_ids + x];
                    const uint32_t cm = cont_mask[i];
                    for (uint32_t itg = 0; itg < qd.group_iters; itg++) g |= (g & cm) >> 1;
                    ec += __popc(g & ~cm);
                }
                return ec;
            };
            const bool node_filters = a.fp.covered_node_count > 0 || a.fp.covered_node_ratio > 0.f;
            // one voted cell: match count, idf, filter_before_matching (filter.rs:76-100)
            auto eval = [&](uint32_t x, uint32_t o, uint32_t v, float p, uint32_t &mc, float &idf) -> bool {
                mc = NARROW ? (v >> 24) : o;
                const uint32_t fixed = NARROW ? (v & 0xffffffu) : v;
                idf = ((float)fixed * inv_scale) * fabsf(p);
                bool pass = (__float_as_uint(p) >> 31) == 0u; // --num-residue, --plddt
                if (a.fp.total_match_count > 0) pass = pass && mc >= a.fp.total_match_count;
                if (a.fp.idf_score_cutoff > 0.f) pass = pass && idf >= a.fp.idf_score_cutoff;
                if (pass && node_filters) {
                    const uint32_t nc = node_count(x);
                    if (a.fp.covered_node_count > 0) pass = pass && nc >= a.fp.covered_node_count;
                    if (a.fp.covered_node_ratio > 0.f)
                        pass = pass && (float)nc / (float)qd.expected_node_count >= a.fp.covered_node_ratio;
                }
                return pass;
            };
            auto clear_cell = [&](uint32_t x) {
                w_acc[x] = 0;
                if (!NARROW) w_match[x] = 0;
                for (uint32_t i = 0; i < EW; i++) w_edge[i * tile_ids + x] = 0;
            };
            auto find_thr = [&]() { // one warp: highest bin such that the bins at or above it hold at least limit_n cells
                uint32_t above = 0, tb = 0;
                for (int c = (int)(s_maxbin >> 5); c >= 0; c--) { // from the highest occupied bin down
                    const uint32_t hv = hist[c * 32 + lane];
                    uint32_t suf = hv; // suffix sum over lanes >= lane
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t u = __shfl_down_sync(0xffffffffu, suf, o);
                        if ((int)lane + o < 32) suf += u;
                    }
                    const uint32_t reach = __ballot_sync(0xffffffffu, above + suf >= a.limit_n);
                    if (reach) {
                        tb = c * 32 + (31 - __clz(reach));
                        break;
                    }
                    above += __shfl_sync(0xffffffffu, suf, 0);
                }
                if (lane == 0) s_thr = tb;
            };
            const bool two_pass = a.limit_n && t == item.tile_begin;
            if (two_pass) {
                // pass A: idf histogram of the passing cells
                uint32_t mybin = 0;
                for (uint32_t x4 = tid; x4 < T4; x4 += blockDim.x) {
                    const uint4 o = occ4[x4];
                    if (!(o.x | o.y | o.z | o.w)) continue;
                    const float4 p = pen4[x4];
                    const uint4 v = NARROW ? o : acc4[x4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t oj = u4_get(o, j);
                        if (!oj) continue;
                        uint32_t mc;
                        float idf;
                        if (eval(4 * x4 + j, oj, u4_get(v, j), f4_get(p, j), mc, idf)) {
                            const uint32_t bin = k3v_idf_bin(idf);
                            atomicAdd(&hist[bin], 1u);
                            mybin = max(mybin, bin);
                        }
                    }
                }
                mybin = __reduce_max_sync(0xffffffffu, mybin);
                if (lane == 0 && mybin) atomicMax(&s_maxbin, mybin);
                __syncthreads();
                if (warp == 0) find_thr();
                else if (warp == 1 && has_next) scan_prefix((t + 1) & 1u);
                __syncthreads();
                run_thr = s_thr;
            }
            // pass B: every voted cell is held against the run's threshold with three instructions (fixed-point idf x
            // penalty >= lower edge of the threshold bin, all scaled by the query's power-of-two factor: the same
            // comparison as idf_bin(idf) >= thr); the few cells that reach it are compacted through the warp's queue
            // and turned into hit records on full lanes; the planes are zeroed on the way.
            {
                const bool feed_hist = a.limit_n && !two_pass;
                const float thr_s = run_thr ? __uint_as_float(run_thr << 20) * scale : 0.f;
                uint32_t *wq = wqueue + warp * K3_WQ;
                uint32_t wq_n = 0; // warp-uniform
                uint32_t mybin = 0;
                auto drain = [&](uint32_t first, uint32_t n_take) { // queue entries [first, first + n_take), one per lane
                    bool emit = false;
                    HitRec rec{0, 0, 0, 0.f};
                    if (lane < n_take) {
                        const uint32_t x = wq[first + lane];
                        uint32_t mc;
                        float idf;
                        if (eval(x, occ[x], w_acc[x], a.pen[lo + x], mc, idf)) {
                            const uint32_t bin = k3v_idf_bin(idf);
                            if (feed_hist) {
                                atomicAdd(&hist[bin], 1u);
                                mybin = max(mybin, bin);
                            }
                            if (bin >= run_thr) {
                                rec = HitRec{lo + x, mc, (node_count(x) << 16) | (edge_count(x) & 0xffffu), idf};
                                emit = true;
                            }
                        }
                        clear_cell(x);
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, emit);
                    if (m) {
                        uint32_t pos = 0;
                        if (lane == 0) pos = atomicAdd(&a.hit_counts[q], (unsigned int)__popc(m));
                        pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1));
                        if (emit) {
                            if (pos < hcap) a.hits[hbase + pos] = rec;
                            else atomicOr(a.overflow, 1u);
                        }
                    }
                };
                const uint32_t n_iter = (T4 + blockDim.x - 1) / blockDim.x;
                for (uint32_t itn = 0; itn < n_iter; itn++) {
                    const uint32_t x4 = (itn * nwarps + warp) * 32 + lane; // a warp reads 512 contiguous bytes
                    uint4 o = make_uint4(0, 0, 0, 0);
                    if (x4 < T4) o = occ4[x4];
                    const bool any = (o.x | o.y | o.z | o.w) != 0;
                    if (!__any_sync(0xffffffffu, any)) continue;
                    uint32_t cmask = 0;
                    if (any) {
                        const float4 p = pen4[x4];
                        const uint4 v = NARROW ? o : acc4[x4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t vj = u4_get(v, j);
                            const float sj = (float)(NARROW ? (vj & 0xffffffu) : vj) * fabsf(f4_get(p, j));
                            cmask |= (u4_get(o, j) != 0u && sj >= thr_s) ? (1u << j) : 0u;
                        }
                        if (!cmask) {
                            const uint4 z = make_uint4(0, 0, 0, 0);
                            reinterpret_cast<uint4 *>(w_acc)[x4] = z;
                            if (!NARROW) reinterpret_cast<uint4 *>(w_match)[x4] = z;
                            for (uint32_t i = 0; i < EW; i++) reinterpret_cast<uint4 *>(w_edge + i * tile_ids)[x4] = z;
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                if (!((cmask >> j) & 1u) && u4_get(o, j)) clear_cell(4 * x4 + j);
                        }
                    }
                    const uint32_t cnt = __popc(cmask);
                    uint32_t incl = cnt;
#pragma unroll
                    for (int ofs = 1; ofs < 32; ofs <<= 1) {
                        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, ofs);
                        if ((int)lane >= ofs) incl += u;
                    }
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total == 0) continue;
                    uint32_t pos = wq_n + incl - cnt;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if ((cmask >> j) & 1u) wq[pos++] = 4 * x4 + j;
                    wq_n += total;
                    __syncwarp();
                    uint32_t done = 0;
                    while (wq_n - done >= 32) {
                        drain(done, 32);
                        done += 32;
                    }
                    if (done) { // move the remainder to the front
                        const uint32_t rem = wq_n - done;
                        uint32_t keepv = 0;
                        if (lane < rem) keepv = wq[done + lane];
                        __syncwarp();
                        if (lane < rem) wq[lane] = keepv;
                        wq_n = rem;
                    }
                    __syncwarp();
                }
                if (wq_n) drain(0, wq_n);
                if (feed_hist) {
                    mybin = __reduce_max_sync(0xffffffffu, mybin);
                    if (lane == 0 && mybin) atomicMax(&s_maxbin, mybin);
                }
            }
            __syncthreads(); // the planes are clean and the next tile's votes may start
            if (!two_pass) { // the threshold for the next tile and the scan of its granule counts
                if (warp == 0 && a.limit_n) find_thr();
                else if (warp == 1 && has_next) scan_prefix((t + 1) & 1u);
                __syncthreads();
                if (a.limit_n) run_thr = s_thr;
            }
            // the first decode step of the next tile
            if (has_next && warp < n_dwarps) prefetch(0, s_total[(t + 1) & 1u]);
        }
    }
}

// The epilogue of k3_scan over merged dense votes: grid (id tiles, queries of the slice).
template <bool NARROW, int EW>
__global__ void __launch_bounds__(K3_THREADS)
    k3_select_dense(IndexView ix, const QueryDesc *queries, const uint16_t *edge_node, const uint16_t *edge_group,
                    const float *idf_sum_per_query, const float *pen, const uint32_t *dense, uint32_t n_queries,
                    uint32_t first_query, uint32_t q_begin, uint32_t tile_ids, FilterParams fp,
                    const uint64_t *hit_offsets, unsigned int *hit_counts, HitRec *hits) {
    __shared__ uint32_t node_mask[K3_MAX_NODES * EW];
    __shared__ uint32_t cont_mask[EW];
    __shared__ uint32_t wqueue[K3_WARPS * K3_WQ];
    const uint32_t qs = blockIdx.y, q = q_begin + qs;
    const QueryDesc qd = queries[q];
    const uint32_t lo = blockIdx.x * tile_ids;
    const uint32_t hi = min(ix.n_structs, lo + tile_ids);
    for (uint32_t i = threadIdx.x; i < qd.n_nodes * EW; i += blockDim.x) node_mask[i] = 0;
    if (threadIdx.x < EW) cont_mask[threadIdx.x] = 0;
    __syncthreads();
    build_edge_masks<EW>(qd, edge_node, edge_group, node_mask, cont_mask);
    __syncthreads();
    const float inv_scale = 1.0f / idf_scale<NARROW>(idf_sum_per_query[q]);
    const size_t plane_stride = (size_t)n_queries * ix.n_structs;
    const uint32_t *acc = dense + (size_t)(q - first_query) * ix.n_structs + lo;
    const uint32_t *match = NARROW ? nullptr : acc + plane_stride;
    const uint32_t *edge = acc + (NARROW ? 1 : 2) * plane_stride;
    emit_cells<NARROW, EW>(acc, match, edge, plane_stride, hi - lo, lo, qd, node_mask, cont_mask, wqueue, inv_scale, ix,
                           pen, fp, &hit_counts[qs], hits + hit_offsets[qs]);
}

// Sparse records {key, planes...} -> dense planes [planes][n_queries][N] of the receiving rank's slice.
__global__ void k3_apply_records(const uint32_t *records, uint64_t n_records, uint32_t planes, uint32_t acc_planes,
                                 uint32_t n_queries, uint32_t n_structs, uint32_t *dense) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_records) return;
    const uint32_t *r = records + k * (1 + planes);
    const uint32_t key = r[0];
    const size_t plane_stride = (size_t)n_queries * n_structs;
    uint32_t *cell = dense + key; // key = local query * N + nid
    for (uint32_t p = 0; p < planes; p++) {
        const uint32_t v = r[1 + p];
        if (v == 0) continue;
        if (p < acc_planes) atomicAdd(cell + p * plane_stride, v);
        else atomicOr(cell + p * plane_stride, v);
    }
}

// sort key: idf descending, nid ascending  (query_pdb.rs:404 stable sort over ascending nid)
__device__ __forceinline__ uint32_t float_desc_key(float f) {
    uint32_t u = __float_as_uint(f);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u); // ascending-order transform
    return ~u;                                      // descending
}
__global__ void k3_make_sort_keys(const HitRec *hits, const uint64_t *hit_offsets, const unsigned int *hit_counts,
                                  uint32_t n_queries, uint64_t *keys, uint32_t *vals) {
    const uint32_t q = blockIdx.y;
    const uint64_t base = hit_offsets[q];
    const uint32_t n = hit_counts[q];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const HitRec r = hits[base + k];
        keys[base + k] = ((uint64_t)float_desc_key(r.idf) << 32) | r.nid;
        vals[base + k] = k;
    }
}
// ---- top-n pre-selection: only hits that can reach the top n of their query are sorted ----
constexpr uint32_t K3_HIST_BINS = 4096; // bin = sign-less float bits >> 19: 8 exponent + 4 mantissa bits, monotone in idf
__device__ __forceinline__ uint32_t idf_bin(float idf) {
    const uint32_t u = __float_as_uint(idf);
    return (u & 0x80000000u) ? 0u : (u >> 19);
}
__global__ void k3_idf_hist(const HitRec *hits, const uint64_t *hit_offsets, const unsigned int *hit_counts,
                            uint32_t *hist) {
    const uint32_t q = blockIdx.y;
    const uint64_t base = hit_offsets[q];
    const uint32_t n = hit_counts[q];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        atomicAdd(&hist[(size_t)q * K3_HIST_BINS + idf_bin(hits[base + k].idf)], 1u);
}
// one warp per query: the lowest bin such that the bins at or above it hold at least top_n hits
__global__ void k3_idf_threshold(const uint32_t *hist, uint32_t n_queries, uint32_t top_n, uint32_t *thr_bin) {
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= n_queries) return;
    const uint32_t *h = hist + (size_t)q * K3_HIST_BINS;
    uint32_t above = 0, thr = 0;
    for (int c = K3_HIST_BINS / 32 - 1; c >= 0; c--) { // chunks of 32 bins from the top; lane 31 = highest bin
        const uint32_t v = h[c * 32 + lane];
        uint32_t suf = v; // suffix sum over lanes >= lane
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_down_sync(0xffffffffu, suf, o);
            if ((int)lane + o < 32) suf += t;
        }
        const uint32_t reach = __ballot_sync(0xffffffffu, above + suf >= top_n);
        if (reach) {
            thr = c * 32 + (31 - __clz(reach)); // highest lane whose suffix already reaches top_n
            break;
        }
        above += __shfl_sync(0xffffffffu, suf, 0);
    }
    if (lane == 0) thr_bin[q] = thr;
}
__global__ void k3_make_sort_keys_top(const HitRec *hits, const uint64_t *hit_offsets, const unsigned int *hit_counts,
                                      const uint32_t *thr_bin, unsigned int *kept, uint64_t *keys, uint32_t *vals) {
    const uint32_t q = blockIdx.y, lane = threadIdx.x & 31;
    const uint64_t base = hit_offsets[q];
    const uint32_t n = hit_counts[q], thr = thr_bin[q];
    const uint32_t n_round = (n + 31) & ~31u;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_round; k += gridDim.x * blockDim.x) {
        HitRec r{0, 0, 0, 0.f};
        bool keep = false;
        if (k < n) {
            r = hits[base + k];
            keep = idf_bin(r.idf) >= thr;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(&kept[q], (unsigned int)__popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1));
            if (keep) {
                keys[base + pos] = ((uint64_t)float_desc_key(r.idf) << 32) | r.nid;
                vals[base + pos] = k;
            }
        }
    }
}

__global__ void k3_segment_ends(const uint64_t *hit_offsets, const unsigned int *hit_counts, uint32_t n_queries,
                                uint64_t *ends) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_queries) ends[q] = hit_offsets[q] + hit_counts[q];
}
__global__ void k3_gather(const HitRec *hits, const uint32_t *sorted_vals, const uint64_t *hit_offsets,
                          const uint64_t *out_offsets, fd_struct_hit *out) {
    const uint32_t q = blockIdx.y;
    const uint64_t base = hit_offsets[q], obase = out_offsets[q];
    const uint32_t n = (uint32_t)(out_offsets[q + 1] - obase);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const HitRec r = hits[base + sorted_vals[base + k]];
        out[obase + k] = fd_struct_hit{r.nid, r.match_count, r.node_edge >> 16, r.node_edge & 0xffffu, r.idf};
    }
}

// ---- per-query top-n in one kernel (pools of at most K3_TOPN_SORT_MAX hits per query) ----
// One CTA per query: the hits of the query's region go to shared memory as 64-bit keys (idf descending, nid ascending:
// query_pdb.rs:404 is a stable sort over ascending nid) with their position, a bitonic sort orders them, and the first
// top_n become fd_struct_hit rows of the PADDED output out[q][top_n] (nid + id_offset: id-range shards report global
// ids), out_count[q] = rows written.  Replaces hist / keys / segmented sort / gather (about ten launches and a host
// round trip) when the tile-level pre-selection has already cut the pools to a few hundred hits.  A larger pool sets
// *too_big and the host falls back to the segmented sort.
constexpr uint32_t K3_TOPN_SORT_MAX = 4096;
__global__ void __launch_bounds__(256) k3_topn_sort(const HitRec *hits, const uint64_t *hit_offsets,
                                                     const unsigned int *hit_counts, uint32_t top_n, uint32_t id_offset,
                                                     fd_struct_hit *out, uint32_t *out_count, unsigned int *too_big) {
    __shared__ uint64_t key[K3_TOPN_SORT_MAX];
    __shared__ uint16_t pos[K3_TOPN_SORT_MAX];
    const uint32_t q = blockIdx.x;
    const uint64_t base = hit_offsets[q];
    const uint32_t cap = (uint32_t)(hit_offsets[q + 1] - base);
    const uint32_t n = min(hit_counts[q], cap);
    if (n > K3_TOPN_SORT_MAX) {
        if (threadIdx.x == 0) {
            atomicOr(too_big, 1u);
            out_count[q] = 0;
        }
        return;
    }
    uint32_t m = 32;
    while (m < n) m <<= 1;
    for (uint32_t k = threadIdx.x; k < m; k += blockDim.x) {
        if (k < n) {
            const HitRec r = hits[base + k];
            key[k] = ((uint64_t)float_desc_key(r.idf) << 32) | r.nid;
        } else {
            key[k] = ~0ull;
        }
        pos[k] = (uint16_t)k;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= m; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t k = threadIdx.x; k < (m >> 1); k += blockDim.x) {
                const uint32_t lo_i = 2 * k - (k & (stride - 1));
                const uint32_t hi_i = lo_i + stride;
                const bool up = (lo_i & size) == 0;
                const uint64_t x = key[lo_i], y = key[hi_i];
                if ((x > y) == up) {
                    key[lo_i] = y;
                    key[hi_i] = x;
                    const uint16_t t = pos[lo_i];
                    pos[lo_i] = pos[hi_i];
                    pos[hi_i] = t;
                }
            }
            __syncthreads();
        }
    const uint32_t n_out = min(n, top_n);
    for (uint32_t k = threadIdx.x; k < n_out; k += blockDim.x) {
        const HitRec r = hits[base + pos[k]];
        out[(size_t)q * top_n + k] = fd_struct_hit{r.nid + id_offset, r.match_count, r.node_edge >> 16, r.node_edge & 0xffffu, r.idf};
    }
    if (threadIdx.x == 0) out_count[q] = n_out;
}

// Owner-side merge of the id-range shards' per-query top-n lists (count_query.rs:172-217 is a per-structure merge, so
// disjoint id ranges merge by concatenation; query_pdb.rs:404-411 sort + truncate then keeps the global top n).
// One CTA per own query: rows lists[r][q][0 .. counts[r][q]) of the world ranks -> out[q][0 .. min(total, top_n)).
__global__ void __launch_bounds__(256) k3_merge_topn(const fd_struct_hit *lists, const uint32_t *counts, uint32_t world,
                                                      uint32_t n_own, uint32_t top_n, fd_struct_hit *out,
                                                      uint32_t *out_count) {
    __shared__ uint64_t key[K3_TOPN_SORT_MAX];
    __shared__ uint16_t pos[K3_TOPN_SORT_MAX]; // r * top_n + k
    __shared__ uint32_t s_off[65];
    const uint32_t q = blockIdx.x;
    if (threadIdx.x == 0) {
        uint32_t o = 0;
        for (uint32_t r = 0; r < world; r++) {
            s_off[r] = o;
            o += min(counts[(size_t)r * n_own + q], top_n);
        }
        s_off[world] = o;
    }
    __syncthreads();
    const uint32_t n = s_off[world];
    uint32_t m = 32;
    while (m < n) m <<= 1;
    for (uint32_t k = threadIdx.x; k < m; k += blockDim.x) {
        key[k] = ~0ull;
        pos[k] = 0;
    }
    __syncthreads();
    for (uint32_t r = 0; r < world; r++) {
        const uint32_t c = s_off[r + 1] - s_off[r];
        for (uint32_t k = threadIdx.x; k < c; k += blockDim.x) {
            const fd_struct_hit h = lists[((size_t)r * n_own + q) * top_n + k];
            key[s_off[r] + k] = ((uint64_t)float_desc_key(h.idf) << 32) | h.nid;
            pos[s_off[r] + k] = (uint16_t)(r * top_n + k);
        }
    }
    __syncthreads();
    for (uint32_t size = 2; size <= m; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t k = threadIdx.x; k < (m >> 1); k += blockDim.x) {
                const uint32_t lo_i = 2 * k - (k & (stride - 1));
                const uint32_t hi_i = lo_i + stride;
                const bool up = (lo_i & size) == 0;
                const uint64_t x = key[lo_i], y = key[hi_i];
                if ((x > y) == up) {
                    key[lo_i] = y;
                    key[hi_i] = x;
                    const uint16_t t = pos[lo_i];
                    pos[lo_i] = pos[hi_i];
                    pos[hi_i] = t;
                }
            }
            __syncthreads();
        }
    const uint32_t n_out = min(n, top_n);
    for (uint32_t k = threadIdx.x; k < n_out; k += blockDim.x) {
        const uint32_t r = pos[k] / top_n, j = pos[k] % top_n;
        out[(size_t)q * top_n + k] = lists[((size_t)r * n_own + q) * top_n + j];
    }
    if (threadIdx.x == 0) out_count[q] = n_out;
}

IndexView make_view(const fd_ctx *ctx) {
    const FdDeviceIndex &d = ctx->idx;
    return IndexView{d.hashes, d.offsets, d.values, d.counts, d.dir, d.skip_off, d.skip_id, d.nres, d.plddt,
                     d.count, (uint32_t)d.n_structs};
}

// ------------------------------------------------------------------------------------------------
// host side of a batch
// ------------------------------------------------------------------------------------------------

FilterParams make_filter(const fd_prefilter_params *params) {
    return FilterParams{params->length_penalty,
                        (uint32_t)std::min<uint64_t>(params->total_match_count, 0xffffffffu),
                        (uint32_t)std::min<uint64_t>(params->covered_node_count, 0xffffffffu),
                        params->covered_node_ratio,
                        params->idf_score_cutoff,
                        (uint32_t)std::min<uint64_t>(params->num_res_cutoff, 0xffffffffu),
                        params->plddt_cutoff};
}

// The batch flattened on the host and uploaded: per-hash arrays, per-edge node ids, query descriptors.
struct Batch {
    std::vector<uint32_t> f_hash, f_gcount;
    std::vector<uint16_t> f_edge, f_edge_node, f_edge_group;
    std::vector<QueryDesc> descs;
    uint32_t max_hashes = 0, max_edges = 0, max_nodes = 0;
    bool narrow = true;
    int ew = 1;
    DevBuf<uint32_t> d_hash, d_gcount;
    DevBuf<uint16_t> d_edge, d_edge_node, d_edge_group;
    DevBuf<QueryDesc> d_desc;
    DevBuf<QHash> d_qh;
    DevBuf<unsigned long long> d_postings, d_bytes;
    DevBuf<float> d_idfsum, d_pen;
    std::vector<unsigned long long> h_postings;
};

// flatten + sample_query (count_query.rs:222-253) + upload + lookup + per-query sums + length-penalty table.
// need_hashes = false: only descriptors and edge nodes are needed (fd_votes_select).
int prepare_batch(fd_ctx *ctx, const fd_query *queries, uint32_t nq, const fd_prefilter_params *params,
                  bool need_hashes, int bound_mode, Batch &B, const uint32_t *gcounts = nullptr, uint64_t g_structs = 0) {
    cudaStream_t s = ctx->stream;
    const uint32_t N = (uint32_t)ctx->idx.n_structs;
    const bool has_r = params->sampling_ratio >= 0.f, has_c = params->sampling_count >= 0;
    const bool sampling = has_r != has_c;
    B.descs.resize(nq);
    std::vector<uint32_t> sample_counts;
    if (sampling) {
        std::vector<uint32_t> all;
        for (uint32_t q = 0; q < nq; q++) all.insert(all.end(), queries[q].hashes, queries[q].hashes + queries[q].n_hashes);
        sample_counts.resize(all.size());
        if (gcounts) std::copy(gcounts, gcounts + all.size(), sample_counts.begin()); // id-range shards: global list lengths
        else FD_TRY(fd_posting_counts(ctx, all.data(), all.size(), sample_counts.data()));
    }
    size_t gbase = 0; // position of the query's first hash in gcounts
    size_t sample_base = 0;
    std::vector<uint32_t> order;
    for (uint32_t q = 0; q < nq; q++) {
        const fd_query &Q = queries[q];
        if (Q.n_hashes && (!Q.hashes || !Q.edge_of_hash)) return fd_fail(ctx, FD_ERR_ARG, "fd_query: NULL array");
        if (Q.n_edges && !Q.edge_node) return fd_fail(ctx, FD_ERR_ARG, "fd_query: NULL edge_node");
        if (sampling) {
            order.resize(Q.n_hashes);
            for (uint32_t k = 0; k < Q.n_hashes; k++) order[k] = k;
            const uint32_t *cnt = sample_counts.data() + sample_base;
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cnt[a] < cnt[b]; });
            size_t keep = has_r ? (size_t)std::ceil(params->sampling_ratio * (float)Q.n_hashes)
                                : (size_t)params->sampling_count;
            if (keep < order.size()) order.resize(keep);
            sample_base += Q.n_hashes;
        }
        uint32_t group_iters = 0;
        if (Q.edge_group) { // runs of equal values = one query edge; a run must stay inside one 32-bit mask word
            uint32_t run = 1;
            for (uint32_t e = 1; e < Q.n_edges; e++) {
                if (Q.edge_group[e] == Q.edge_group[e - 1]) {
                    if ((e & 31) == 0) return fd_fail(ctx, FD_ERR_ARG, "fd_query: an edge_group run crosses a multiple of 32");
                    run++;
                } else {
                    run = 1;
                }
                group_iters = std::max(group_iters, run - 1);
            }
        }
        const uint32_t n_kept = sampling ? (uint32_t)order.size() : Q.n_hashes;
        B.descs[q] = QueryDesc{(uint32_t)B.f_hash.size(), n_kept, (uint32_t)B.f_edge_node.size(),
                               Q.n_edges, Q.n_nodes, Q.expected_node_count, group_iters};
        if (!sampling) { // the common case: the arrays are taken as they are (bulk copies; a sharded rank flattens
            // the whole batch of every rank, millions of hashes per call)
            uint32_t worst = 0;
            for (uint32_t k = 0; k < Q.n_hashes; k++) worst = std::max<uint32_t>(worst, Q.edge_of_hash[k]);
            if (Q.n_hashes && worst >= Q.n_edges) return fd_fail(ctx, FD_ERR_ARG, "fd_query: edge_of_hash out of range");
            B.f_hash.insert(B.f_hash.end(), Q.hashes, Q.hashes + Q.n_hashes);
            B.f_edge.insert(B.f_edge.end(), Q.edge_of_hash, Q.edge_of_hash + Q.n_hashes);
            if (gcounts) B.f_gcount.insert(B.f_gcount.end(), gcounts + gbase, gcounts + gbase + Q.n_hashes);
        } else {
            for (uint32_t k : order) {
                if (Q.edge_of_hash[k] >= Q.n_edges) return fd_fail(ctx, FD_ERR_ARG, "fd_query: edge_of_hash out of range");
                B.f_hash.push_back(Q.hashes[k]);
                B.f_edge.push_back(Q.edge_of_hash[k]);
                if (gcounts) B.f_gcount.push_back(gcounts[gbase + k]);
            }
        }
        gbase += Q.n_hashes;
        for (uint32_t e = 0; e < Q.n_edges; e++)
            if (Q.edge_node[e] >= Q.n_nodes) return fd_fail(ctx, FD_ERR_ARG, "fd_query: edge_node out of range");
        B.f_edge_node.insert(B.f_edge_node.end(), Q.edge_node, Q.edge_node + Q.n_edges);
        if (Q.edge_group) B.f_edge_group.insert(B.f_edge_group.end(), Q.edge_group, Q.edge_group + Q.n_edges);
        else
            for (uint32_t e = 0; e < Q.n_edges; e++) B.f_edge_group.push_back((uint16_t)e);
        B.max_hashes = std::max<uint32_t>(B.max_hashes, n_kept);
        B.max_edges = std::max(B.max_edges, Q.n_edges);
        B.max_nodes = std::max(B.max_nodes, Q.n_nodes);
    }
    if (B.max_hashes > K3_MAX_HASHES || B.max_edges > 32 * K3_MAX_EDGE_WORDS || B.max_nodes > K3_MAX_NODES)
        return fd_fail(ctx, FD_ERR_LIMIT,
                       "query too large for the shared-memory vote kernel (limits: 4095 hashes, 256 edges, 256 "
                       "nodes per query); the dense / sparse vote exchange of hash-range shards has no wide path");
    B.narrow = B.max_hashes <= K3_MAX_HASHES_NARROW;
    B.ew = B.max_edges <= 32 ? 1 : (B.max_edges <= 64 ? 2 : (B.max_edges <= 128 ? 4 : 8));
    if (nq == 0 || N == 0) return FD_OK;
    const uint32_t nqh = (uint32_t)B.f_hash.size();
    FD_CUDA(ctx, B.d_edge_node.alloc(B.f_edge_node.size()));
    FD_CUDA(ctx, B.d_edge_group.alloc(B.f_edge_group.size()));
    FD_CUDA(ctx, B.d_desc.alloc(nq));
    FD_CUDA(ctx, B.d_idfsum.alloc(nq));
    FD_CUDA(ctx, B.d_pen.alloc(N));
    FD_CUDA(ctx, B.d_postings.alloc(nq));
    FD_CUDA(ctx, B.d_bytes.alloc(nq));
    FD_CUDA(ctx, B.d_hash.alloc(nqh));
    FD_CUDA(ctx, B.d_edge.alloc(nqh));
    FD_CUDA(ctx, B.d_qh.alloc(nqh));
    FD_CUDA(ctx, cudaMemcpyAsync(B.d_edge_node.p, B.f_edge_node.data(), B.f_edge_node.size() * 2, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(B.d_edge_group.p, B.f_edge_group.data(), B.f_edge_group.size() * 2, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(B.d_desc.p, B.descs.data(), nq * sizeof(QueryDesc), cudaMemcpyHostToDevice, s));
    IndexView ix = make_view(ctx);
    B.h_postings.assign(nq, 0);
    StageTimer st(ctx, "lookup");
    FD_LAUNCH(ctx, k3_length_penalty, fd_div_up(N, 256), 256, 0, ix.nres, N, params->length_penalty, B.d_pen.p);
    if (need_hashes && nqh) {
        FD_CUDA(ctx, cudaMemcpyAsync(B.d_hash.p, B.f_hash.data(), nqh * 4, cudaMemcpyHostToDevice, s));
        FD_CUDA(ctx, cudaMemcpyAsync(B.d_edge.p, B.f_edge.data(), nqh * 2, cudaMemcpyHostToDevice, s));
        if (gcounts) {
            FD_CUDA(ctx, B.d_gcount.alloc(nqh));
            FD_CUDA(ctx, cudaMemcpyAsync(B.d_gcount.p, B.f_gcount.data(), nqh * 4, cudaMemcpyHostToDevice, s));
        }
        FD_LAUNCH(ctx, k3_lookup, fd_div_up(nqh, 256), 256, 0, ix, B.d_hash.p, nqh, params->freq_filter,
                  gcounts ? B.d_gcount.p : (const uint32_t *)nullptr, (uint32_t)g_structs, B.d_qh.p);
    } else if (nqh) {
        FD_CUDA(ctx, cudaMemsetAsync(B.d_qh.p, 0, nqh * sizeof(QHash), s));
    }
    FD_LAUNCH(ctx, k3_query_sums, fd_div_up(nq, 128), 128, 0, B.d_desc.p, B.d_qh.p, nq,
              gcounts ? (uint32_t)g_structs : N, bound_mode, B.d_postings.p, B.d_bytes.p, B.d_idfsum.p);
    if (need_hashes) {
        std::vector<unsigned long long> h_bytes(nq);
        FD_CUDA(ctx, cudaMemcpyAsync(B.h_postings.data(), B.d_postings.p, nq * 8, cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, cudaMemcpyAsync(h_bytes.data(), B.d_bytes.p, nq * 8, cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, st.finish());
        ctx->last_posting_bytes = 0;
        for (unsigned long long b : h_bytes) ctx->last_posting_bytes += b;
    } else {
        FD_CUDA(ctx, st.finish());
    }
    return FD_OK;
}

// motif-sized queries keep their per-list descriptors in shared memory (24 B per hash)
static inline bool k3_stage_lists(uint32_t max_hashes) { return max_hashes <= 512; }

struct TilePlan {
    uint32_t tile_ids, n_tiles, threads;
    size_t smem;
};

int plan_tiles(fd_ctx *ctx, const Batch &B0, uint32_t N, TilePlan &tp, int ew = 0) {
    struct {
        bool narrow;
        int ew;
        uint32_t max_hashes, max_nodes;
    } B{B0.narrow, ew > 0 ? ew : B0.ew, B0.max_hashes, B0.max_nodes};
    const uint32_t bytes_per_id = (B.narrow ? 4 : 8) + 4 * B.ew;
    // Every tile's CTA walks all of the query's short lists, so fewer, larger tiles decode less: when the id range
    // fits the shared memory of one or two SMs a query gets 1024-thread CTAs with 220 KB tiles; otherwise 256-thread
    // CTAs with 72 KB tiles, three per SM.
    auto fixed_for = [&](uint32_t threads) {
        return (size_t)((k3_stage_lists(B.max_hashes) ? 8 : 2) * B.max_hashes + 2 + B.max_nodes * B.ew + B.ew +
                        (threads / 32) * K3_WQ) * 4 + 64;
    };
    uint32_t threads = K3_THREADS;
    size_t budget = 72 * 1024;
    if ((size_t)((N + 31) & ~31u) * bytes_per_id <= 2 * (220 * 1024 - fixed_for(K3_MAX_THREADS))) {
        threads = K3_MAX_THREADS; // one or two tiles cover every structure (measured: 0.187 vs 0.216 ms, profiles/)
        budget = 220 * 1024;
    }
    if (const char *e = getenv("FD_K3_TILE_KB")) budget = (size_t)std::max(8, atoi(e)) * 1024;
    if (const char *e = getenv("FD_K3_THREADS")) threads = (uint32_t)std::min(1024, std::max(64, atoi(e) & ~31));
    budget = std::min<size_t>(budget, 220 * 1024);
    const size_t fixed_smem = fixed_for(threads);
    tp.threads = threads;
    const size_t avail = budget > fixed_smem + 256 * bytes_per_id ? budget - fixed_smem : 256 * bytes_per_id;
    uint32_t tile_ids = (uint32_t)(avail / bytes_per_id) & ~31u;
    tile_ids = std::max<uint32_t>(256, tile_ids);
    if (tile_ids > N) tile_ids = (N + 31) & ~31u;
    tp.tile_ids = tile_ids;
    tp.n_tiles = fd_div_up(N, tile_ids);
    tp.smem = (size_t)tile_ids * bytes_per_id + fixed_smem;
    if (tp.smem > 220 * 1024) return fd_fail(ctx, FD_ERR_LIMIT, "query needs more shared memory than one SM has");
    return FD_OK;
}

template <bool NARROW, int EW, int MODE>
int launch_scan_t(fd_ctx *ctx, dim3 grid, const TilePlan &tp, IndexView ix, const Batch &B, FilterParams fp,
                  const uint64_t *hit_offsets, unsigned int *hit_counts, HitRec *hits, uint32_t *dense, SparseOut sp,
                  const uint32_t *qlist) {
    auto kern = k3_scan<NARROW, EW, MODE>;
    FD_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp.smem));
    kern<<<grid, tp.threads, tp.smem, ctx->stream>>>(ix, B.d_desc.p, B.d_qh.p, B.d_edge.p, B.d_edge_node.p,
                                                     B.d_edge_group.p, B.d_idfsum.p, B.d_pen.p, tp.tile_ids, fp, hit_offsets, hit_counts,
                                                     hits, dense, sp, k3_stage_lists(B.max_hashes) ? 1 : 0, qlist);
    ctx->launches++;
    return FD_OK;
}

template <int MODE>
int launch_scan(fd_ctx *ctx, dim3 grid, const TilePlan &tp, IndexView ix, const Batch &B, FilterParams fp,
                const uint64_t *hit_offsets, unsigned int *hit_counts, HitRec *hits, uint32_t *dense,
                SparseOut sp = SparseOut{nullptr, nullptr, nullptr, 0, 0}, const uint32_t *qlist = nullptr, int ew = 0) {
    if (ew <= 0) ew = B.ew;
#define FD_SCAN_CASE(NARROW, EW) \
    return launch_scan_t<NARROW, EW, MODE>(ctx, grid, tp, ix, B, fp, hit_offsets, hit_counts, hits, dense, sp, qlist)
    if (B.narrow) {
        if (ew == 1) FD_SCAN_CASE(true, 1);
        else if (ew == 2) FD_SCAN_CASE(true, 2);
        else if (ew == 4) FD_SCAN_CASE(true, 4);
        else FD_SCAN_CASE(true, 8);
    } else {
        if (ew == 1) FD_SCAN_CASE(false, 1);
        else if (ew == 2) FD_SCAN_CASE(false, 2);
        else if (ew == 4) FD_SCAN_CASE(false, 4);
        else FD_SCAN_CASE(false, 8);
    }
#undef FD_SCAN_CASE
    return FD_ERR_ARG;
}

// Queries grouped by the number of edge-mask words they need (1, 2, 4, 8): qorder = query numbers class by class.
// One scan launch per class keeps the tiles of narrow queries large (8 B of shared memory per structure for <= 32
// vote bits); sizing every tile for the widest query of the batch would split all of them.
static void edge_word_classes(const Batch &B, uint32_t nq, std::vector<uint32_t> &qorder, uint32_t class_begin[5]) {
    qorder.clear();
    class_begin[0] = 0;
    for (int c = 0; c < 4; c++) {
        for (uint32_t q = 0; q < nq; q++) {
            const uint32_t need = std::max(1u, (B.descs[q].n_edges + 31) / 32);
            const int cls = need <= 1 ? 0 : need <= 2 ? 1 : need <= 4 ? 2 : 3;
            if (cls == c) qorder.push_back(q);
        }
        class_begin[c + 1] = (uint32_t)qorder.size();
    }
}

template <bool NARROW, int EW>
int launch_select_t(fd_ctx *ctx, dim3 grid, IndexView ix, const Batch &B, const uint32_t *dense,
                    uint32_t n_dense_queries, uint32_t first_query, uint32_t q_begin, uint32_t tile_ids, FilterParams fp, const uint64_t *hit_offsets, unsigned int *hit_counts,
                    HitRec *hits) {
    k3_select_dense<NARROW, EW><<<grid, K3_THREADS, 0, ctx->stream>>>(
        ix, B.d_desc.p, B.d_edge_node.p, B.d_edge_group.p, B.d_idfsum.p, B.d_pen.p, dense, n_dense_queries,
        first_query, q_begin, tile_ids, fp, hit_offsets, hit_counts, hits);
    ctx->launches++;
    return FD_OK;
}

// Orders the hit regions by (idf desc, nid asc), keeps top_n per query and copies them to the host.
// hit_off (host) / d_hit_off: region offsets of nsel queries; d_hit_cnt: survivors per query.
int select_and_copy(fd_ctx *ctx, uint32_t nsel, const std::vector<uint64_t> &hit_off, const DevBuf<uint64_t> &d_hit_off,
                    const DevBuf<unsigned int> &d_hit_cnt, const DevBuf<HitRec> &d_hits, uint64_t top_n, uint32_t N,
                    fd_struct_hit **out_hits, uint64_t *h_off) {
    cudaStream_t s = ctx->stream;
    const uint64_t pool = hit_off[nsel];
    std::vector<unsigned int> h_cnt(nsel);
    DevBuf<uint64_t> d_seg_end, d_keys, d_keys2, d_out_off;
    DevBuf<uint32_t> d_vals, d_vals2;
    StageTimer st(ctx, "select");
    FD_CUDA(ctx, cudaMemcpyAsync(h_cnt.data(), d_hit_cnt.p, nsel * 4, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, d_keys.alloc(pool));
    FD_CUDA(ctx, d_keys2.alloc(pool));
    FD_CUDA(ctx, d_vals.alloc(pool));
    FD_CUDA(ctx, d_vals2.alloc(pool));
    FD_CUDA(ctx, d_seg_end.alloc(nsel));
    dim3 g2(std::max<uint32_t>(1, std::min<uint32_t>(64, fd_div_up(N, 256))), nsel);
    DevBuf<uint32_t> d_hist, d_thr;
    DevBuf<unsigned int> d_kept;
    uint64_t topsel_ratio = 128; // the pre-selection pays once the sort would be much larger than the histogram
    if (const char *e = getenv("FD_K3_TOPSEL_RATIO")) topsel_ratio = strtoull(e, nullptr, 10);
    if (top_n < 0x7fffffffull && pool > topsel_ratio * top_n * nsel) {
        // only the hits that can reach the top n of their query take part in the sort
        FD_CUDA(ctx, d_hist.alloc((size_t)nsel * K3_HIST_BINS));
        FD_CUDA(ctx, d_thr.alloc(nsel));
        FD_CUDA(ctx, d_kept.alloc(nsel));
        FD_CUDA(ctx, cudaMemsetAsync(d_hist.p, 0, (size_t)nsel * K3_HIST_BINS * 4, s));
        FD_CUDA(ctx, cudaMemsetAsync(d_kept.p, 0, nsel * 4, s));
        FD_LAUNCH(ctx, k3_idf_hist, g2, 256, 0, d_hits.p, d_hit_off.p, d_hit_cnt.p, d_hist.p);
        FD_LAUNCH(ctx, k3_idf_threshold, fd_div_up((uint64_t)nsel * 32, 256), 256, 0, d_hist.p, nsel, (uint32_t)top_n, d_thr.p);
        FD_LAUNCH(ctx, k3_make_sort_keys_top, g2, 256, 0, d_hits.p, d_hit_off.p, d_hit_cnt.p, d_thr.p, d_kept.p, d_keys.p,
                  d_vals.p);
        FD_LAUNCH(ctx, k3_segment_ends, fd_div_up(nsel, 256), 256, 0, d_hit_off.p, d_kept.p, nsel, d_seg_end.p);
    } else {
        FD_LAUNCH(ctx, k3_make_sort_keys, g2, 256, 0, d_hits.p, d_hit_off.p, d_hit_cnt.p, nsel, d_keys.p, d_vals.p);
        FD_LAUNCH(ctx, k3_segment_ends, fd_div_up(nsel, 256), 256, 0, d_hit_off.p, d_hit_cnt.p, nsel, d_seg_end.p);
    }
    size_t tb = 0;
    cub::DeviceSegmentedSort::SortPairs(nullptr, tb, d_keys.p, d_keys2.p, d_vals.p, d_vals2.p, (int64_t)pool,
                                        (int64_t)nsel, d_hit_off.p, d_seg_end.p, s);
    DevBuf<uint8_t> tmp;
    FD_CUDA(ctx, tmp.alloc(tb));
    if (pool)
        FD_CUDA(ctx, cub::DeviceSegmentedSort::SortPairs(tmp.p, tb, d_keys.p, d_keys2.p, d_vals.p, d_vals2.p,
                                                         (int64_t)pool, (int64_t)nsel, d_hit_off.p, d_seg_end.p, s));
    ctx->launches += 3;
    FD_CUDA(ctx, cudaStreamSynchronize(s));
    h_off[0] = 0;
    for (uint32_t q = 0; q < nsel; q++) h_off[q + 1] = h_off[q] + std::min<uint64_t>(h_cnt[q], top_n);
    const uint64_t n_out = h_off[nsel];
    DevBuf<fd_struct_hit> d_out;
    FD_CUDA(ctx, d_out.alloc(n_out));
    FD_CUDA(ctx, d_out_off.alloc(nsel + 1));
    FD_CUDA(ctx, cudaMemcpyAsync(d_out_off.p, h_off, (nsel + 1) * 8, cudaMemcpyHostToDevice, s));
    if (n_out) FD_LAUNCH(ctx, k3_gather, g2, 256, 0, d_hits.p, d_vals2.p, d_hit_off.p, d_out_off.p, d_out.p);
    fd_struct_hit *h_hits = (fd_struct_hit *)malloc(std::max<uint64_t>(n_out, 1) * sizeof(fd_struct_hit));
    if (!h_hits) return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
    cudaError_t e = cudaMemcpyAsync(h_hits, d_out.p, n_out * sizeof(fd_struct_hit), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = st.finish();
    if (e != cudaSuccess) {
        free(h_hits);
        return fd_fail(ctx, FD_ERR_CUDA, std::string("select: ") + cudaGetErrorString(e));
    }
    *out_hits = h_hits;
    return FD_OK;
}

} // namespace

extern "C" {

int fd_index_attach(fd_ctx *ctx, const uint32_t *hashes, const uint64_t *offsets, uint64_t count,
                    const uint8_t *values, uint64_t value_bytes, uint64_t n_structs, const uint32_t *nres,
                    const float *plddt) {
    if (!ctx) return FD_ERR_ARG;
    if (!offsets || (count && (!hashes || !values)) || (n_structs && !nres))
        return fd_fail(ctx, FD_ERR_ARG, "fd_index_attach: NULL argument");
    if (count > 0xfffffff0ull) return fd_fail(ctx, FD_ERR_LIMIT, "more than 2^32 distinct hashes");
    if (n_structs > 0xfffffff0ull) return fd_fail(ctx, FD_ERR_LIMIT, "structure ids must fit in 32 bits");
    if (offsets[0] != 0 || offsets[count] != value_bytes)
        return fd_fail(ctx, FD_ERR_ARG, "fd_index_attach: offsets[0] must be 0 and offsets[count] == value_bytes");
    if (ctx->borrowed) return fd_fail(ctx, FD_ERR_STATE, "fd_index_attach: a forked context shares its parent's index");
    FD_ENTER(ctx);
    fd_ctx_release_index(ctx);
    FdDeviceIndex &d = ctx->idx;
    cudaStream_t s = ctx->stream;
    const uint64_t nblocks = (value_bytes >> SKIP_SHIFT) + 2;
    FD_CUDA(ctx, cudaMalloc(&d.hashes, std::max<uint64_t>(count, 1) * 4));
    FD_CUDA(ctx, cudaMalloc(&d.offsets, (count + 1) * 8));
    FD_CUDA(ctx, cudaMalloc(&d.values, value_bytes + VALUES_PAD));
    FD_CUDA(ctx, cudaMalloc(&d.counts, std::max<uint64_t>(count, 1) * 4));
    FD_CUDA(ctx, cudaMalloc(&d.dir, ((uint64_t)DIR_SIZE + 2) * 4));
    FD_CUDA(ctx, cudaMalloc(&d.skip_off, nblocks));
    FD_CUDA(ctx, cudaMalloc(&d.skip_id, nblocks * 4));
    FD_CUDA(ctx, cudaMalloc(&d.nres, std::max<uint64_t>(n_structs, 1) * 4));
    FD_CUDA(ctx, cudaMalloc(&d.plddt, std::max<uint64_t>(n_structs, 1) * 4));
    d.count = count;
    d.value_bytes = value_bytes;
    d.n_structs = n_structs;
    d.n_skip = nblocks;
    FD_CUDA(ctx, cudaMemcpyAsync(d.hashes, hashes, count * 4, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d.offsets, offsets, (count + 1) * 8, cudaMemcpyHostToDevice, s));
    FD_TRY(fd_copy_to_device_staged(ctx, d.values, values, value_bytes));
    FD_CUDA(ctx, cudaMemsetAsync(d.values + value_bytes, 0, VALUES_PAD, s));
    FD_CUDA(ctx, cudaMemsetAsync(d.skip_off, 0, nblocks, s));
    FD_CUDA(ctx, cudaMemsetAsync(d.skip_id, 0, nblocks * 4, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d.nres, nres, n_structs * 4, cudaMemcpyHostToDevice, s));
    if (plddt) FD_CUDA(ctx, cudaMemcpyAsync(d.plddt, plddt, n_structs * 4, cudaMemcpyHostToDevice, s));
    else FD_CUDA(ctx, cudaMemsetAsync(d.plddt, 0, std::max<uint64_t>(n_structs, 1) * 4, s));

    StageTimer st(ctx, "attach");
    FD_LAUNCH(ctx, k3_build_dir, fd_div_up((uint64_t)DIR_SIZE + 1, 256), 256, 0, d.hashes, count, d.dir);
    // posting counts per list
    DevBuf<uint32_t> terms;
    DevBuf<uint64_t> terms64, block_prefix, partial, scanned;
    DevBuf<uint8_t> tmp;
    FD_CUDA(ctx, terms.alloc(nblocks));
    FD_CUDA(ctx, terms64.alloc(nblocks));
    FD_CUDA(ctx, block_prefix.alloc(nblocks));
    FD_CUDA(ctx, partial.alloc(nblocks));
    FD_CUDA(ctx, scanned.alloc(nblocks));
    FD_CUDA(ctx, cudaMemsetAsync(terms.p, 0, nblocks * 4, s));
    FD_LAUNCH(ctx, k3_block_terms, fd_div_up(nblocks, 256), 256, 0, d.values, value_bytes, nblocks, terms.p);
    FD_LAUNCH(ctx, k3_widen32, fd_div_up(nblocks, 256), 256, 0, terms.p, nblocks, terms64.p);
    size_t tb1 = 0, tb2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb1, terms64.p, block_prefix.p, nblocks, s);
    cub::DeviceScan::InclusiveScan(nullptr, tb2, partial.p, scanned.p, SegScanOp(), nblocks, s);
    FD_CUDA(ctx, tmp.alloc(std::max(tb1, tb2)));
    size_t tb = tb1;
    FD_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb, terms64.p, block_prefix.p, nblocks, s));
    ctx->launches += 2;
    if (count) FD_LAUNCH(ctx, k3_list_counts, fd_div_up(count, 256), 256, 0, d.values, d.offsets, block_prefix.p, count, d.counts);
    // skip table
    if (count) {
        FD_LAUNCH(ctx, k3_skip_partials, fd_div_up(nblocks, 256), 256, 0, d.values, d.offsets, count, value_bytes,
                  nblocks, d.skip_off, partial.p);
        tb = tb2;
        FD_CUDA(ctx, cub::DeviceScan::InclusiveScan(tmp.p, tb, partial.p, scanned.p, SegScanOp(), nblocks, s));
        ctx->launches += 2;
        FD_LAUNCH(ctx, k3_low32, fd_div_up(nblocks, 256), 256, 0, scanned.p, nblocks, d.skip_id);
    }
    FD_CUDA(ctx, st.finish());
    d.attached = true;
    ctx->idx_generation++;
    return FD_OK;
}

int fd_posting_counts(fd_ctx *ctx, const uint32_t *hashes, uint64_t n, uint32_t *out_counts) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->idx.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_posting_counts: no index attached");
    if (n && (!hashes || !out_counts)) return fd_fail(ctx, FD_ERR_ARG, "fd_posting_counts: NULL argument");
    FD_ENTER(ctx);
    if (n == 0) return FD_OK;
    DevBuf<uint32_t> dh, dc;
    FD_CUDA(ctx, dh.alloc(n));
    FD_CUDA(ctx, dc.alloc(n));
    FD_CUDA(ctx, cudaMemcpyAsync(dh.p, hashes, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    StageTimer st(ctx, "lookup");
    FD_LAUNCH(ctx, k3_counts_only, fd_div_up(n, 256), 256, 0, make_view(ctx), dh.p, n, dc.p);
    FD_CUDA(ctx, cudaMemcpyAsync(out_counts, dc.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, st.finish());
    return FD_OK;
}

int fd_get_entries(fd_ctx *ctx, uint32_t hash, uint64_t **out_ids, uint64_t *out_n) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->idx.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_get_entries: no index attached");
    if (!out_ids || !out_n) return fd_fail(ctx, FD_ERR_ARG, "fd_get_entries: NULL argument");
    FD_ENTER(ctx);
    uint32_t cnt = 0;
    FD_TRY(fd_posting_counts(ctx, &hash, 1, &cnt));
    DevBuf<uint64_t> d_ids;
    DevBuf<unsigned long long> d_n;
    FD_CUDA(ctx, d_ids.alloc(cnt));
    FD_CUDA(ctx, d_n.alloc(1));
    FD_LAUNCH(ctx, k3_decode_list, 1, 32, 0, make_view(ctx), hash, d_ids.p, d_n.p, (uint64_t)cnt);
    unsigned long long n = 0;
    FD_CUDA(ctx, cudaMemcpyAsync(&n, d_n.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n != cnt) return fd_fail(ctx, FD_ERR_STATE, "fd_get_entries: posting count table disagrees with decode");
    uint64_t *h = (uint64_t *)malloc(std::max<uint64_t>(n, 1) * 8);
    if (!h) return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
    FD_CUDA(ctx, cudaMemcpy(h, d_ids.p, n * 8, cudaMemcpyDeviceToHost));
    *out_ids = h;
    *out_n = n;
    return FD_OK;
}

// the first scan kernel (one CTA per (tile, query), one launch per edge-word class); FD_K3_V1=1 selects it
static int scan_v1(fd_ctx *ctx, const Batch &B, uint32_t nq, uint32_t N, const fd_prefilter_params *params,
                   const DevBuf<uint64_t> &d_hit_off, DevBuf<unsigned int> &d_hit_cnt, DevBuf<HitRec> &d_hits) {
    cudaStream_t s = ctx->stream;
    std::vector<uint32_t> qorder;
    uint32_t class_begin[5];
    const int class_ew[4] = {1, 2, 4, 8};
    edge_word_classes(B, nq, qorder, class_begin);
    DevBuf<uint32_t> d_qorder;
    FD_CUDA(ctx, d_qorder.alloc(nq));
    FD_CUDA(ctx, cudaMemcpyAsync(d_qorder.p, qorder.data(), nq * 4, cudaMemcpyHostToDevice, s));
    StageTimer st(ctx, "scan");
    for (int c = 0; c < 4; c++) {
        const uint32_t nc = class_begin[c + 1] - class_begin[c];
        if (!nc) continue;
        const int ew = std::min(class_ew[c], B.ew);
        TilePlan tp;
        FD_TRY(plan_tiles(ctx, B, N, tp, ew));
        FD_TRY(launch_scan<0>(ctx, dim3(tp.n_tiles, nc), tp, make_view(ctx), B, make_filter(params), d_hit_off.p,
                              d_hit_cnt.p, d_hits.p, nullptr, SparseOut{nullptr, nullptr, nullptr, 0, 0},
                              d_qorder.p + class_begin[c], ew));
    }
    FD_CUDA(ctx, st.finish());
    return FD_OK;
}

// Shared-memory plan of k3_scan_v3: CTAs per SM, threads, bytes per CTA, words of vote planes.
struct ScanV3Plan {
    uint32_t ctas_per_sm, threads, tile_words, n_dwarps;
    size_t smem;
    bool stage_lists;
};
static int plan_scan_v3(fd_ctx *ctx, const Batch &B, ScanV3Plan &pl) {
    uint32_t ctas = 2, threads = 512;
    if (const char *e = getenv("FD_K3_CTAS")) ctas = (uint32_t)std::min(4, std::max(1, atoi(e)));
    if (const char *e = getenv("FD_K3_THREADS")) threads = (uint32_t)std::min(K3V_MAX_THREADS, std::max(64, atoi(e) & ~31));
    if (ctas * threads > 1024) threads = (1024 / ctas) & ~31u; // the kernel is compiled for 64 registers per thread
    const uint32_t max_ew = k3v_edge_words(B.max_edges);
    pl.n_dwarps = K3V_DECODE_WARPS;
    if (const char *e = getenv("FD_K3_DWARPS")) pl.n_dwarps = (uint32_t)std::min(32, std::max(1, atoi(e)));
    pl.n_dwarps = std::min(pl.n_dwarps, threads / 32);
    pl.stage_lists = k3_stage_lists(B.max_hashes);
    // 228 KB per SM, 1 KB reserved per resident CTA, at most 227 KB per CTA
    const size_t per_cta = std::min<size_t>(227 * 1024, (228 * 1024) / ctas - 1024 - 128); // 1 KB reserved per CTA + the kernel's static shared memory
    const size_t fixed = ((size_t)(pl.stage_lists ? 9 : 3) * B.max_hashes + 2 + (size_t)B.max_nodes * max_ew + max_ew +
                          K3V_HIST_BINS + (size_t)(threads / 32) * K3_WQ) * 4 + 16 + (size_t)pl.n_dwarps * (K3V_WARP_STAGE + 8) + 64;
    const uint32_t planes_max = (B.narrow ? 1u : 2u) + max_ew;
    if (per_cta < fixed + (size_t)256 * planes_max * 4) {
        if (ctas > 1) { // very wide queries: one CTA per SM
            ctas = 1;
            threads = std::min<uint32_t>(threads, K3V_MAX_THREADS);
            const size_t one = 227 * 1024;
            if (one < fixed + (size_t)256 * planes_max * 4)
                return fd_fail(ctx, FD_ERR_LIMIT, "query needs more shared memory than one SM has");
            pl.ctas_per_sm = 1;
            pl.threads = threads;
            pl.tile_words = (uint32_t)((one - fixed) / 4) & ~3u;
            pl.smem = (size_t)pl.tile_words * 4 + fixed;
            return FD_OK;
        }
        return fd_fail(ctx, FD_ERR_LIMIT, "query needs more shared memory than one SM has");
    }
    pl.ctas_per_sm = ctas;
    pl.threads = threads;
    pl.tile_words = (uint32_t)((per_cta - fixed) / 4) & ~3u;
    pl.smem = (size_t)pl.tile_words * 4 + fixed;
    return FD_OK;
}

// ------------------------------------------------------------------------------------------------
// Wide queries (whole-structure queries: an empty -q makes every residue a query residue, query.rs:226-233; a
// 300-residue chain has ~3 10^4 query edges and ~10^5 hashes).  Their vote state does not fit shared memory, so
// count_query (count_query.rs:82-220) runs in global memory, organised like the reference itself -- by NODE GROUP
// (the hashes whose edge starts at one query residue): all four fields are sums over the node groups, so the groups
// are processed in chunks whose decoded postings fit a budget:
//   k3w_count / k3w_decode   one thread per 64-byte granule of the chunk's posting lists (skip table -> any granule
//                            decodes on its own): postings -> keys  nid << kbits | k   (k = position of the hash in
//                            the query's (node, edge)-sorted hash order)
//   radix sort of the keys over the bits in use, then ONE segmented reduction by nid whose values are computed on
//   the fly from adjacent keys: {1, idf weight of k, first key of its (nid, edge), first key of its (nid, node)}
//   k3w_accumulate           the chunk's per-structure sums into the dense accumulators (a structure appears once
//                            per chunk: plain adds)
//   k3w_emit                 length penalty, filter_before_matching, hit records; the usual top-n selection follows.
// The idf sum is accumulated in 2^-32 fixed point (64 bits): independent of the summation order.
// ------------------------------------------------------------------------------------------------
struct WideAgg {
    uint32_t match, edge, node, pad;
    unsigned long long idf;
};
struct WideAggSum {
    __host__ __device__ WideAgg operator()(const WideAgg &a, const WideAgg &b) const {
        return WideAgg{a.match + b.match, a.edge + b.edge, a.node + b.node, 0u, a.idf + b.idf};
    }
};
struct WideKeyNid { // sorted key -> structure id
    const uint64_t *keys;
    uint32_t kbits;
    __host__ __device__ uint32_t operator()(uint64_t i) const { return (uint32_t)(keys[i] >> kbits); }
};
struct WideKeyValue { // sorted key -> the posting's contribution
    const uint64_t *keys;
    const uint32_t *edge_of_k, *node_of_k;
    const unsigned long long *w_of_k;
    uint32_t kbits;
    __host__ __device__ WideAgg operator()(uint64_t i) const {
        const uint64_t key = keys[i], kmask = (1ull << kbits) - 1ull;
        const uint32_t k = (uint32_t)(key & kmask);
        bool new_edge = true, new_node = true;
        if (i > 0) {
            const uint64_t prev = keys[i - 1];
            if ((prev >> kbits) == (key >> kbits)) {
                const uint32_t pk = (uint32_t)(prev & kmask);
                new_edge = edge_of_k[pk] != edge_of_k[k];
                new_node = node_of_k[pk] != node_of_k[k];
            }
        }
        return WideAgg{1u, new_edge ? 1u : 0u, new_node ? 1u : 0u, 0u, w_of_k[k]};
    }
};

__global__ void k3w_list_granules(const QHash *qh, uint32_t k0, uint32_t n, uint64_t *nseg) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const QHash h = qh[k0 + k];
    nseg[k] = h.end > h.start ? ((h.end - 1) >> SKIP_SHIFT) - (h.start >> SKIP_SHIFT) + 1 : 0;
}

// the granule `item` of the chunk: its list and byte range; false if no varint starts in it
__device__ __forceinline__ bool k3w_item(const IndexView &ix, const QHash *qh, uint32_t k0, const uint64_t *gprefix,
                                         uint32_t n_lists, uint64_t item, uint32_t &k, uint64_t &gi, uint32_t &p0,
                                         uint32_t &p1, uint32_t &id0) {
    uint32_t x = 0, y = n_lists; // last list with gprefix[list] <= item
    while (y - x > 1) {
        const uint32_t m = (x + y) >> 1;
        if (gprefix[m] <= item) x = m;
        else y = m;
    }
    k = k0 + x;
    const QHash h = qh[k];
    const uint64_t g = item - gprefix[x];
    gi = (h.start >> SKIP_SHIFT) + g;
    const uint64_t G0 = gi << SKIP_SHIFT;
    p0 = g == 0 ? (uint32_t)(h.start - G0) : ix.skip_off[gi];
    id0 = g == 0 ? 0u : ix.skip_id[gi];
    p1 = (uint32_t)min(h.end - G0, (uint64_t)SKIP_BYTES);
    return p0 < p1;
}

// varints that start in [p0, p1) = 1 + terminators in [p0, p1 - 1)
__global__ void k3w_count(IndexView ix, const QHash *qh, uint32_t k0, const uint64_t *gprefix, uint32_t n_lists,
                          uint64_t n_items, uint64_t *counts) {
    const uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_items) return;
    uint32_t k, p0, p1, id0;
    uint64_t gi;
    uint64_t c = 0;
    if (k3w_item(ix, qh, k0, gprefix, n_lists, item, k, gi, p0, p1, id0)) {
        const uint4 *src = reinterpret_cast<const uint4 *>(ix.values + (gi << SKIP_SHIFT));
        uint64_t term = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint4 v = src[j];
            term |= (uint64_t)(term_mask4(v.x) | (term_mask4(v.y) << 4) | (term_mask4(v.z) << 8) | (term_mask4(v.w) << 12)) << (16 * j);
        }
        const uint64_t lo_mask = ~((1ull << p0) - 1ull);
        const uint64_t hi_mask = p1 - 1 >= 64 ? ~0ull : ((1ull << (p1 - 1)) - 1ull);
        c = 1 + __popcll(term & lo_mask & hi_mask);
    }
    counts[item] = c;
}

__global__ void k3w_decode(IndexView ix, const QHash *qh, uint32_t k0, const uint64_t *gprefix, uint32_t n_lists,
                           uint64_t n_items, const uint64_t *out_pos, uint32_t kbits, uint32_t id_cap, uint64_t *keys) {
    const uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_items) return;
    uint32_t k, p0, p1, id;
    uint64_t gi;
    if (!k3w_item(ix, qh, k0, gprefix, n_lists, item, k, gi, p0, p1, id)) return;
    const uint8_t *src = ix.values + (gi << SKIP_SHIFT);
    uint64_t pos = out_pos[item];
    uint32_t p = p0;
    while (p < p1) { // a varint that starts before p1 may end in the bytes after it (the value array is padded)
        uint32_t v = 0, shift = 0, byte;
        do {
            byte = src[p++];
            v |= (byte & 0x7fu) << shift;
            shift += 7;
        } while (byte & 0x80u);
        id += v;
        keys[pos++] = ((uint64_t)min(id, id_cap) << kbits) | k; // ids beyond the lookup share one key prefix, skipped later
    }
}

__global__ void k3w_accumulate(const uint32_t *nids, const WideAgg *agg, const uint32_t *n_runs, uint32_t n_structs,
                               uint32_t *match, uint32_t *edges, uint32_t *nodes, unsigned long long *idf) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *n_runs) return;
    const uint32_t nid = nids[r];
    if (nid >= n_structs) return; // count_query.rs:133: ids beyond the lookup are skipped
    const WideAgg a = agg[r];
    match[nid] += a.match;
    edges[nid] += a.edge;
    nodes[nid] += a.node;
    idf[nid] += a.idf;
}

__global__ void k3w_emit(const uint32_t *match, const uint32_t *edges, const uint32_t *nodes, const unsigned long long *idf,
                         const float *pen_signed, uint32_t n_structs, FilterParams fp, uint32_t expected_node_count,
                         HitRec *hits, unsigned int *n_hits) {
    const uint32_t nid = blockIdx.x * blockDim.x + threadIdx.x;
    bool pass = false;
    HitRec rec{0, 0, 0, 0.f};
    if (nid < n_structs && match[nid] > 0) {
        const float p = pen_signed[nid];
        const float sum = (float)((double)idf[nid] * (1.0 / 4294967296.0));
        const float v = sum * fabsf(p);
        pass = (__float_as_uint(p) >> 31) == 0u;
        if (fp.total_match_count > 0) pass = pass && match[nid] >= fp.total_match_count;
        if (fp.idf_score_cutoff > 0.f) pass = pass && v >= fp.idf_score_cutoff;
        if (fp.covered_node_count > 0) pass = pass && nodes[nid] >= fp.covered_node_count;
        if (fp.covered_node_ratio > 0.f) pass = pass && (float)nodes[nid] / (float)expected_node_count >= fp.covered_node_ratio;
        // HitRec packs node and edge counts in 16 bits each; wide queries report them through the side arrays
        rec = HitRec{nid, match[nid], 0u, v};
    }
    const uint32_t m = __ballot_sync(0xffffffffu, pass);
    if (m) {
        uint32_t pos = 0;
        if ((threadIdx.x & 31) == 0) pos = atomicAdd(n_hits, (unsigned int)__popc(m));
        pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << (threadIdx.x & 31)) - 1));
        if (pass) hits[pos] = rec;
    }
}

// count_query + filter + sort + top for ONE wide query; hits appended to `out` (idf descending, nid ascending)
static int count_query_wide(fd_ctx *ctx, const fd_query &Q, const fd_prefilter_params *params, const uint32_t *gcounts,
                            uint64_t g_structs, uint32_t id_offset, std::vector<fd_struct_hit> &out) {
    cudaStream_t s = ctx->stream;
    const uint32_t N = (uint32_t)ctx->idx.n_structs;
    if (Q.n_hashes == 0 || N == 0) return FD_OK;
    if (!Q.hashes || !Q.edge_of_hash || (Q.n_edges && !Q.edge_node)) return fd_fail(ctx, FD_ERR_ARG, "fd_query: NULL array");
    if (Q.edge_group) return fd_fail(ctx, FD_ERR_LIMIT, "edge groups (hash-range shards) are not supported for whole-structure queries");
    for (uint32_t k = 0; k < Q.n_hashes; k++)
        if (Q.edge_of_hash[k] >= Q.n_edges) return fd_fail(ctx, FD_ERR_ARG, "fd_query: edge_of_hash out of range");
    for (uint32_t e = 0; e < Q.n_edges; e++)
        if (Q.edge_node[e] >= Q.n_nodes) return fd_fail(ctx, FD_ERR_ARG, "fd_query: edge_node out of range");
    IndexView ix = make_view(ctx);
    // hash order: by (node, edge), so that a structure's sorted keys list its edges and nodes in runs
    std::vector<uint32_t> order(Q.n_hashes);
    for (uint32_t k = 0; k < Q.n_hashes; k++) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        const uint32_t ea = Q.edge_of_hash[a], eb = Q.edge_of_hash[b];
        const uint32_t na = Q.edge_node[ea], nb = Q.edge_node[eb];
        return na != nb ? na < nb : ea < eb;
    });
    std::vector<uint32_t> f_hash(Q.n_hashes), f_edge(Q.n_hashes), f_node(Q.n_hashes), f_gcount;
    for (uint32_t k = 0; k < Q.n_hashes; k++) {
        f_hash[k] = Q.hashes[order[k]];
        f_edge[k] = Q.edge_of_hash[order[k]];
        f_node[k] = Q.edge_node[f_edge[k]];
    }
    if (gcounts) {
        f_gcount.resize(Q.n_hashes);
        for (uint32_t k = 0; k < Q.n_hashes; k++) f_gcount[k] = gcounts[order[k]];
    }
    const uint32_t nh = Q.n_hashes;
    DevBuf<uint32_t> d_hash, d_edge, d_node, d_gcount;
    DevBuf<QHash> d_qh;
    FD_CUDA(ctx, d_hash.alloc(nh));
    FD_CUDA(ctx, d_edge.alloc(nh));
    FD_CUDA(ctx, d_node.alloc(nh));
    FD_CUDA(ctx, d_qh.alloc(nh));
    FD_CUDA(ctx, cudaMemcpyAsync(d_hash.p, f_hash.data(), nh * 4ull, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_edge.p, f_edge.data(), nh * 4ull, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_node.p, f_node.data(), nh * 4ull, cudaMemcpyHostToDevice, s));
    if (gcounts) {
        FD_CUDA(ctx, d_gcount.alloc(nh));
        FD_CUDA(ctx, cudaMemcpyAsync(d_gcount.p, f_gcount.data(), nh * 4ull, cudaMemcpyHostToDevice, s));
    }
    std::vector<QHash> h_qh(nh);
    {
        StageTimer st(ctx, "lookup");
        FD_LAUNCH(ctx, k3_lookup, fd_div_up(nh, 256), 256, 0, ix, d_hash.p, nh, params->freq_filter,
                  gcounts ? d_gcount.p : (const uint32_t *)nullptr, (uint32_t)g_structs, d_qh.p);
        FD_CUDA(ctx, cudaMemcpyAsync(h_qh.data(), d_qh.p, nh * sizeof(QHash), cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, st.finish());
    }
    // sample_query (count_query.rs:222-253): keep the rarest hashes (ties in the caller's hash order)
    const bool has_r = params->sampling_ratio >= 0.f, has_c = params->sampling_count >= 0;
    if (has_r != has_c) {
        std::vector<uint32_t> cnt(nh);
        if (gcounts) cnt = f_gcount;
        else {
            // list lengths irrespective of the frequency filter
            DevBuf<uint32_t> d_cnt;
            FD_CUDA(ctx, d_cnt.alloc(nh));
            FD_LAUNCH(ctx, k3_counts_only, fd_div_up(nh, 256), 256, 0, ix, d_hash.p, (uint64_t)nh, d_cnt.p);
            FD_CUDA(ctx, cudaMemcpyAsync(cnt.data(), d_cnt.p, nh * 4ull, cudaMemcpyDeviceToHost, s));
            FD_CUDA(ctx, cudaStreamSynchronize(s));
        }
        std::vector<uint32_t> by_caller(nh); // position in the caller's order -> position in the (node, edge) order
        for (uint32_t k = 0; k < nh; k++) by_caller[order[k]] = k;
        std::vector<uint32_t> rank(nh);
        for (uint32_t k = 0; k < nh; k++) rank[k] = by_caller[k];
        std::stable_sort(rank.begin(), rank.end(), [&](uint32_t a, uint32_t b) { return cnt[a] < cnt[b]; });
        const size_t keep = has_r ? (size_t)std::ceil(params->sampling_ratio * (float)nh) : (size_t)params->sampling_count;
        std::vector<uint8_t> kept(nh, 0);
        for (size_t k = 0; k < std::min<size_t>(keep, nh); k++) kept[rank[k]] = 1;
        for (uint32_t k = 0; k < nh; k++)
            if (!kept[k]) h_qh[k] = QHash{0, 0, 0, 0.f};
        FD_CUDA(ctx, cudaMemcpyAsync(d_qh.p, h_qh.data(), nh * sizeof(QHash), cudaMemcpyHostToDevice, s));
    }
    // per-hash idf weights in 2^-32 fixed point; chunks of whole node groups within the key budget
    std::vector<unsigned long long> f_w(nh);
    uint64_t total_postings = 0, total_bytes = 0;
    for (uint32_t k = 0; k < nh; k++) {
        f_w[k] = h_qh[k].end > h_qh[k].start ? (unsigned long long)((double)std::max(h_qh[k].idf, 0.f) * 4294967296.0 + 0.5) : 0ull;
        total_postings += h_qh[k].count;
        total_bytes += h_qh[k].end - h_qh[k].start;
    }
    ctx->last_posting_bytes += total_bytes;
    DevBuf<unsigned long long> d_w;
    FD_CUDA(ctx, d_w.alloc(nh));
    FD_CUDA(ctx, cudaMemcpyAsync(d_w.p, f_w.data(), nh * 8ull, cudaMemcpyHostToDevice, s));
    DevBuf<uint32_t> a_match, a_edge, a_node;
    DevBuf<unsigned long long> a_idf;
    FD_CUDA(ctx, a_match.alloc(N));
    FD_CUDA(ctx, a_edge.alloc(N));
    FD_CUDA(ctx, a_node.alloc(N));
    FD_CUDA(ctx, a_idf.alloc(N));
    FD_CUDA(ctx, cudaMemsetAsync(a_match.p, 0, N * 4ull, s));
    FD_CUDA(ctx, cudaMemsetAsync(a_edge.p, 0, N * 4ull, s));
    FD_CUDA(ctx, cudaMemsetAsync(a_node.p, 0, N * 4ull, s));
    FD_CUDA(ctx, cudaMemsetAsync(a_idf.p, 0, N * 8ull, s));
    uint64_t budget = 256ull << 20; // keys per chunk (2 x 8 B each)
    if (const char *e = getenv("FD_K3W_CHUNK_KEYS")) budget = std::max<uint64_t>(1024, strtoull(e, nullptr, 10));
    uint32_t kbits = 1, nbits = 1;
    while ((1ull << kbits) < nh) kbits++;
    while ((1ull << nbits) < (uint64_t)N + 1) nbits++;
    StageTimer st(ctx, "scan");
    for (uint32_t k0 = 0; k0 < nh;) {
        uint32_t k1 = k0;
        uint64_t chunk_postings = 0;
        while (k1 < nh) { // whole node groups
            uint32_t k2 = k1;
            uint64_t g = 0;
            while (k2 < nh && f_node[k2] == f_node[k1]) g += h_qh[k2++].count;
            if (k1 > k0 && chunk_postings + g > budget) break;
            chunk_postings += g;
            k1 = k2;
        }
        const uint32_t nl = k1 - k0;
        if (chunk_postings) {
            DevBuf<uint64_t> d_nseg, d_gprefix, d_counts, d_pos, d_keys, d_keys2;
            DevBuf<uint8_t> d_tmp;
            FD_CUDA(ctx, d_nseg.alloc(nl + 1));
            FD_CUDA(ctx, d_gprefix.alloc(nl + 1));
            FD_CUDA(ctx, cudaMemsetAsync(d_nseg.p + nl, 0, 8, s));
            FD_LAUNCH(ctx, k3w_list_granules, fd_div_up(nl, 256), 256, 0, d_qh.p, k0, nl, d_nseg.p);
            size_t tb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, d_nseg.p, d_gprefix.p, nl + 1, s);
            FD_CUDA(ctx, d_tmp.alloc(tb));
            FD_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp.p, tb, d_nseg.p, d_gprefix.p, nl + 1, s));
            uint64_t n_items = 0;
            FD_CUDA(ctx, cudaMemcpyAsync(&n_items, d_gprefix.p + nl, 8, cudaMemcpyDeviceToHost, s));
            FD_CUDA(ctx, cudaStreamSynchronize(s));
            FD_CUDA(ctx, d_counts.alloc(n_items + 1));
            FD_CUDA(ctx, d_pos.alloc(n_items + 1));
            FD_CUDA(ctx, cudaMemsetAsync(d_counts.p + n_items, 0, 8, s));
            FD_LAUNCH(ctx, k3w_count, fd_div_up(n_items, 256), 256, 0, ix, d_qh.p, k0, d_gprefix.p, nl, n_items, d_counts.p);
            size_t tb2 = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb2, d_counts.p, d_pos.p, n_items + 1, s);
            if (tb2 > d_tmp.n) FD_CUDA(ctx, d_tmp.alloc(tb2));
            FD_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp.p, tb2, d_counts.p, d_pos.p, n_items + 1, s));
            uint64_t n_keys = 0;
            FD_CUDA(ctx, cudaMemcpyAsync(&n_keys, d_pos.p + n_items, 8, cudaMemcpyDeviceToHost, s));
            FD_CUDA(ctx, cudaStreamSynchronize(s));
            if (n_keys != chunk_postings) return fd_fail(ctx, FD_ERR_STATE, "wide count_query: posting count table disagrees with decode");
            FD_CUDA(ctx, d_keys.alloc(n_keys));
            FD_CUDA(ctx, d_keys2.alloc(n_keys));
            FD_LAUNCH(ctx, k3w_decode, fd_div_up(n_items, 256), 256, 0, ix, d_qh.p, k0, d_gprefix.p, nl, n_items, d_pos.p,
                      kbits, N, d_keys.p);
            size_t tb3 = 0;
            cub::DeviceRadixSort::SortKeys(nullptr, tb3, d_keys.p, d_keys2.p, n_keys, 0, (int)(kbits + nbits), s);
            if (tb3 > d_tmp.n) FD_CUDA(ctx, d_tmp.alloc(tb3));
            FD_CUDA(ctx, cub::DeviceRadixSort::SortKeys(d_tmp.p, tb3, d_keys.p, d_keys2.p, n_keys, 0, (int)(kbits + nbits), s));
            ctx->launches += 10;
            // segmented reduction by structure id
            const uint64_t max_runs = std::min<uint64_t>(n_keys, (uint64_t)N + 1);
            DevBuf<uint32_t> d_nids, d_nruns;
            DevBuf<WideAgg> d_agg;
            FD_CUDA(ctx, d_nids.alloc(max_runs));
            FD_CUDA(ctx, d_agg.alloc(max_runs));
            FD_CUDA(ctx, d_nruns.alloc(1));
            cub::CountingInputIterator<uint64_t> idx(0);
            cub::TransformInputIterator<uint32_t, WideKeyNid, cub::CountingInputIterator<uint64_t>> key_it(idx, WideKeyNid{d_keys2.p, kbits});
            cub::TransformInputIterator<WideAgg, WideKeyValue, cub::CountingInputIterator<uint64_t>> val_it(
                idx, WideKeyValue{d_keys2.p, d_edge.p, d_node.p, d_w.p, kbits});
            size_t tb4 = 0;
            cub::DeviceReduce::ReduceByKey(nullptr, tb4, key_it, d_nids.p, val_it, d_agg.p, d_nruns.p, WideAggSum(), n_keys, s);
            if (tb4 > d_tmp.n) FD_CUDA(ctx, d_tmp.alloc(tb4));
            FD_CUDA(ctx, cub::DeviceReduce::ReduceByKey(d_tmp.p, tb4, key_it, d_nids.p, val_it, d_agg.p, d_nruns.p, WideAggSum(), n_keys, s));
            ctx->launches += 2;
            FD_LAUNCH(ctx, k3w_accumulate, fd_div_up(max_runs, 256), 256, 0, d_nids.p, d_agg.p, d_nruns.p, N, a_match.p,
                      a_edge.p, a_node.p, a_idf.p);
            FD_CUDA(ctx, cudaStreamSynchronize(s)); // the chunk's buffers are released here
        }
        k0 = k1;
    }
    // length penalty + filters + hit records, then sort + top
    const FilterParams fp = make_filter(params);
    DevBuf<float> d_pen;
    DevBuf<HitRec> d_hits;
    DevBuf<unsigned int> d_nh;
    FD_CUDA(ctx, d_pen.alloc(N));
    FD_CUDA(ctx, d_hits.alloc(N));
    FD_CUDA(ctx, d_nh.alloc(1));
    FD_CUDA(ctx, cudaMemsetAsync(d_nh.p, 0, 4, s));
    FD_LAUNCH(ctx, k3_length_penalty_signed, fd_div_up(N, 256), 256, 0, ix.nres, ix.plddt, N, fp.length_penalty,
              fp.num_res_cutoff, fp.plddt_cutoff, d_pen.p);
    FD_LAUNCH(ctx, k3w_emit, fd_div_up(N, 256), 256, 0, a_match.p, a_edge.p, a_node.p, a_idf.p, d_pen.p, N, fp,
              Q.expected_node_count, d_hits.p, d_nh.p);
    unsigned int n_hits = 0;
    FD_CUDA(ctx, cudaMemcpyAsync(&n_hits, d_nh.p, 4, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, st.finish());
    if (n_hits == 0) return FD_OK;
    // sort (idf descending, nid ascending) and keep top n; node / edge counts come from the dense accumulators
    std::vector<HitRec> h_hits(n_hits);
    std::vector<uint32_t> h_edge(N), h_node(N);
    {
        StageTimer st2(ctx, "select");
        FD_CUDA(ctx, cudaMemcpyAsync(h_hits.data(), d_hits.p, (size_t)n_hits * sizeof(HitRec), cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, cudaMemcpyAsync(h_edge.data(), a_edge.p, N * 4ull, cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, cudaMemcpyAsync(h_node.data(), a_node.p, N * 4ull, cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, st2.finish());
    }
    const size_t keep_n = (size_t)std::min<uint64_t>(params->top_n, n_hits);
    auto better = [](const HitRec &a, const HitRec &b) { return a.idf != b.idf ? a.idf > b.idf : a.nid < b.nid; };
    if (keep_n < h_hits.size()) {
        std::partial_sort(h_hits.begin(), h_hits.begin() + keep_n, h_hits.end(), better);
        h_hits.resize(keep_n);
    } else {
        std::sort(h_hits.begin(), h_hits.end(), better);
    }
    for (const HitRec &r : h_hits)
        out.push_back(fd_struct_hit{r.nid + id_offset, r.match_count, h_node[r.nid], h_edge[r.nid], r.idf});
    return FD_OK;
}

static inline bool query_is_wide(const fd_query &Q) {
    return Q.n_hashes > (uint32_t)K3_MAX_HASHES || Q.n_edges > 32u * K3_MAX_EDGE_WORDS || Q.n_nodes > (uint32_t)K3_MAX_NODES;
}

struct CountOpts {
    const uint32_t *gcounts = nullptr; // id-range shards: global posting count of every query hash (flattened, batch order)
    uint64_t g_structs = 0;            // and the structure count of the whole database
    uint32_t id_offset = 0;            // first structure id of the shard: added to the reported nid
    const uint32_t *slice_begin = nullptr; // [world + 1] owner rank of every query: exchange + merge (fd_count_query_sharded)
    uint64_t batch_id = 0;             // != 0: the caller's promise that this is the batch of the previous call with this id
};

// The flattened batch with its lookup results, kept on the context between calls that carry the same batch id: a
// serving loop that searches one prepared batch again (and every rank of a sharded search, which flattens the WHOLE
// batch of all ranks) then skips flatten + upload + lookup (0.6 ms per 1 024 queries, mostly host time).
struct CqCache {
    uint64_t id = 0, idx_gen = 0, posting_bytes = 0;
    fd_prefilter_params params;
    bool sharded_counts = false;
    uint64_t g_structs = 0;
    Batch B;
};
static void cq_cache_free(void *p) { delete (CqCache *)p; }

static int count_query_impl(fd_ctx *ctx, const fd_query *queries, uint32_t nq, const fd_prefilter_params *params,
                            const CountOpts &opts, fd_struct_hit **out_hits, uint64_t **out_offsets) {
    *out_hits = nullptr;
    *out_offsets = nullptr;
    cudaStream_t s = ctx->stream;
    const uint32_t N = (uint32_t)ctx->idx.n_structs;
    ctx->last_posting_bytes = 0;
    {
        // whole-structure queries do not fit the shared-memory vote kernel: they take the global-memory path one by
        // one, the rest of the batch the usual one; the rows are put back in the caller's order
        std::vector<uint32_t> wide, small;
        for (uint32_t q = 0; q < nq; q++) (query_is_wide(queries[q]) ? wide : small).push_back(q);
        if (!wide.empty()) {
            if (opts.slice_begin)
                return fd_fail(ctx, FD_ERR_LIMIT, "whole-structure queries are not supported by the sharded search (use fd_count_query_batch_ex per shard)");
            std::vector<size_t> gbase(nq + 1, 0);
            for (uint32_t q = 0; q < nq; q++) gbase[q + 1] = gbase[q] + queries[q].n_hashes;
            std::vector<std::vector<fd_struct_hit>> rows(nq);
            uint64_t bytes = 0;
            if (!small.empty()) {
                std::vector<fd_query> sq;
                std::vector<uint32_t> sg;
                for (uint32_t q : small) {
                    sq.push_back(queries[q]);
                    if (opts.gcounts) sg.insert(sg.end(), opts.gcounts + gbase[q], opts.gcounts + gbase[q + 1]);
                }
                CountOpts so = opts;
                so.gcounts = opts.gcounts ? sg.data() : nullptr;
                fd_struct_hit *sh = nullptr;
                uint64_t *soff = nullptr;
                FD_TRY(count_query_impl(ctx, sq.data(), (uint32_t)sq.size(), params, so, &sh, &soff));
                for (size_t k = 0; k < small.size(); k++) rows[small[k]].assign(sh + soff[k], sh + soff[k + 1]);
                free(sh);
                free(soff);
                bytes = ctx->last_posting_bytes;
            }
            ctx->last_posting_bytes = bytes;
            for (uint32_t q : wide)
                FD_TRY(count_query_wide(ctx, queries[q], params, opts.gcounts ? opts.gcounts + gbase[q] : nullptr,
                                        opts.g_structs, opts.id_offset, rows[q]));
            uint64_t *h_off = (uint64_t *)calloc((size_t)nq + 1, 8);
            if (!h_off) return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
            for (uint32_t q = 0; q < nq; q++) h_off[q + 1] = h_off[q] + rows[q].size();
            fd_struct_hit *h_hits = (fd_struct_hit *)malloc(std::max<uint64_t>(h_off[nq], 1) * sizeof(fd_struct_hit));
            if (!h_hits) {
                free(h_off);
                return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
            }
            for (uint32_t q = 0; q < nq; q++)
                if (!rows[q].empty()) memcpy(h_hits + h_off[q], rows[q].data(), rows[q].size() * sizeof(fd_struct_hit));
            *out_hits = h_hits;
            *out_offsets = h_off;
            return FD_OK;
        }
    }
    Batch B_local;
    Batch *Bp = &B_local;
    // id-range shards: the fixed-point scale of the idf sum must be the same on every rank -> bound from the global
    // list lengths (k3_query_sums sums the idf of all query hashes, present locally or not)
    {
        HostTimer ht(ctx, "cq_host_prepare"); // flatten + upload + lookup (includes the lookup kernels' wait)
        CqCache *cache = (CqCache *)ctx->cq_cache;
        const bool hit = opts.batch_id != 0 && cache && cache->id == opts.batch_id && cache->idx_gen == ctx->idx_generation &&
                         memcmp(&cache->params, params, sizeof(*params)) == 0 &&
                         cache->sharded_counts == (opts.gcounts != nullptr) && cache->g_structs == opts.g_structs &&
                         cache->B.descs.size() == nq;
        if (hit) {
            Bp = &cache->B;
            ctx->last_posting_bytes = cache->posting_bytes;
        } else if (opts.batch_id != 0) {
            if (ctx->cq_cache) cq_cache_free(ctx->cq_cache);
            ctx->cq_cache = nullptr;
            cache = new CqCache();
            ctx->cq_cache = cache;
            ctx->cq_cache_free = cq_cache_free;
            const int rc = prepare_batch(ctx, queries, nq, params, true, 0, cache->B, opts.gcounts, opts.g_structs);
            if (rc != FD_OK) return rc; // (the cache keeps id 0: never matched)
            cache->id = opts.batch_id;
            cache->idx_gen = ctx->idx_generation;
            cache->params = *params;
            cache->sharded_counts = opts.gcounts != nullptr;
            cache->g_structs = opts.g_structs;
            cache->posting_bytes = ctx->last_posting_bytes;
            Bp = &cache->B;
        } else {
            FD_TRY(prepare_batch(ctx, queries, nq, params, true, 0, B_local, opts.gcounts, opts.g_structs));
        }
    }
    const Batch &B = *Bp;
    HostTimer ht_rest(ctx, "cq_host_rest"); // work items, pools, scan, select, exchange, copies
    uint64_t *h_off = (uint64_t *)calloc((size_t)nq + 1, 8);
    if (!h_off) return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
    if (opts.slice_begin && (nq == 0 || N == 0 || B.f_hash.empty())) {
        free(h_off);
        return fd_fail(ctx, FD_ERR_ARG, "fd_count_query_sharded: empty batch or empty shard (every rank takes part in the exchange)");
    }
    if (nq == 0 || N == 0 || B.f_hash.empty()) {
        *out_offsets = h_off;
        *out_hits = (fd_struct_hit *)malloc(sizeof(fd_struct_hit));
        return FD_OK;
    }
    const bool use_v1 = getenv("FD_K3_V1") && atoi(getenv("FD_K3_V1")) != 0;
    ScanV3Plan pl{};
    if (!use_v1) {
        const int rc = plan_scan_v3(ctx, B, pl);
        if (rc != FD_OK) {
            free(h_off);
            return rc;
        }
    }
    // work items (query, run of consecutive id tiles), costliest queries first; a query is cut into as many runs as
    // it takes to give every resident CTA a few items (a single query still spreads over the whole GPU)
    std::vector<ScanItem> items;
    std::vector<uint32_t> n_tiles(nq, 1);
    if (!use_v1) {
        std::vector<uint32_t> qorder(nq);
        uint32_t n_live = 0;
        for (uint32_t q = 0; q < nq; q++) {
            qorder[q] = q;
            n_live += B.h_postings[q] != 0;
        }
        std::stable_sort(qorder.begin(), qorder.end(),
                         [&](uint32_t x, uint32_t y) { return B.h_postings[x] > B.h_postings[y]; });
        const uint32_t want_items = 2u * (uint32_t)ctx->num_sms * pl.ctas_per_sm;
        uint32_t runs_per_query = n_live ? fd_div_up(want_items, n_live) : 1;
        if (const char *e = getenv("FD_K3_RUNS")) runs_per_query = (uint32_t)std::max(1, atoi(e));
        for (uint32_t q : qorder) {
            if (B.h_postings[q] == 0) continue;
            const uint32_t planes = (B.narrow ? 1u : 2u) + k3v_edge_words(B.descs[q].n_edges);
            const uint32_t tile_ids = k3v_tile_ids(pl.tile_words, planes, N);
            n_tiles[q] = fd_div_up(N, tile_ids);
            const uint32_t runs = std::min(runs_per_query, n_tiles[q]);
            for (uint32_t r = 0; r < runs; r++) {
                const uint32_t t0 = (uint32_t)((uint64_t)n_tiles[q] * r / runs), t1 = (uint32_t)((uint64_t)n_tiles[q] * (r + 1) / runs);
                if (t1 > t0) items.push_back(ScanItem{q, t0, t1});
            }
        }
    }
    const uint64_t top_n = params->top_n;
    auto attempt = [&](bool limit) -> int {
        // hit regions: a query can hit at most min(N, its postings) structures; with the tile-level pre-selection a
        // tile emits its top n plus the rest of the threshold bin
        std::vector<uint64_t> hit_off(nq + 1, 0);
        for (uint32_t q = 0; q < nq; q++) {
            uint64_t cap = std::min<uint64_t>(N, B.h_postings[q]);
            if (limit) cap = std::min<uint64_t>(cap, (uint64_t)n_tiles[q] * (top_n + std::max<uint64_t>(1024, top_n)));
            hit_off[q + 1] = hit_off[q] + cap;
        }
        DevBuf<uint64_t> d_hit_off;
        DevBuf<unsigned int> d_hit_cnt, d_flags;
        DevBuf<HitRec> d_hits;
        DevBuf<ScanItem> d_items;
        DevBuf<float> d_pen2;
        FD_CUDA(ctx, d_hit_off.alloc(nq + 1));
        FD_CUDA(ctx, d_hit_cnt.alloc(nq));
        FD_CUDA(ctx, d_hits.alloc(hit_off[nq]));
        FD_CUDA(ctx, d_flags.alloc(2)); // [0] work counter, [1] overflow
        FD_CUDA(ctx, cudaMemcpyAsync(d_hit_off.p, hit_off.data(), (nq + 1) * 8, cudaMemcpyHostToDevice, s));
        FD_CUDA(ctx, cudaMemsetAsync(d_hit_cnt.p, 0, nq * 4, s));
        FD_CUDA(ctx, cudaMemsetAsync(d_flags.p, 0, 8, s));
        unsigned int h_flags[2] = {0, 0};
        if (use_v1) {
            FD_TRY(scan_v1(ctx, B, nq, N, params, d_hit_off, d_hit_cnt, d_hits));
        } else if (!items.empty()) {
            FD_CUDA(ctx, d_items.alloc(items.size()));
            FD_CUDA(ctx, d_pen2.alloc((size_t)N + 4)); // the epilogue reads it four cells at a time
            FD_CUDA(ctx, cudaMemsetAsync(d_pen2.p + N, 0, 16, s));
            FD_CUDA(ctx, cudaMemcpyAsync(d_items.p, items.data(), items.size() * sizeof(ScanItem), cudaMemcpyHostToDevice, s));
            const FilterParams fp = make_filter(params);
            IndexView ix = make_view(ctx);
            StageTimer st(ctx, "scan");
            FD_LAUNCH(ctx, k3_length_penalty_signed, fd_div_up(N, 256), 256, 0, ix.nres, ix.plddt, N, fp.length_penalty,
                      fp.num_res_cutoff, fp.plddt_cutoff, d_pen2.p);
            const uint32_t grid = (uint32_t)std::min<uint64_t>(items.size(), (uint64_t)ctx->num_sms * pl.ctas_per_sm);
            ScanV3Args a{B.d_desc.p, B.d_qh.p, B.d_edge.p, B.d_edge_node.p, B.d_edge_group.p, B.d_idfsum.p, d_pen2.p,
                         d_items.p, (uint32_t)items.size(), d_flags.p, pl.tile_words, B.max_hashes, B.max_nodes,
                         k3v_edge_words(B.max_edges), pl.stage_lists ? 1u : 0u, limit ? (uint32_t)top_n : 0u, fp,
                         d_hit_off.p, d_hit_cnt.p, d_hits.p, d_flags.p + 1,
                         (getenv("FD_K3_BULK") && atoi(getenv("FD_K3_BULK")) == 0) ? 0u : 1u, pl.n_dwarps};
            if (B.narrow) {
                FD_CUDA(ctx, cudaFuncSetAttribute(k3_scan_v3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
                FD_LAUNCH(ctx, k3_scan_v3<true>, grid, pl.threads, pl.smem, ix, a);
            } else {
                FD_CUDA(ctx, cudaFuncSetAttribute(k3_scan_v3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
                FD_LAUNCH(ctx, k3_scan_v3<false>, grid, pl.threads, pl.smem, ix, a);
            }
            FD_CUDA(ctx, cudaMemcpyAsync(h_flags, d_flags.p, 8, cudaMemcpyDeviceToHost, s));
            FD_CUDA(ctx, st.finish());
        }
        if (h_flags[1]) return 1; // a hit region overflowed
        const bool sharded = opts.slice_begin != nullptr;
        if ((limit && top_n <= 4096 && !(getenv("FD_K3_TOPN_SORT") && atoi(getenv("FD_K3_TOPN_SORT")) == 0)) || sharded) {
            // pools of a few hundred hits per query: one sort kernel, padded rows, one copy
            DevBuf<fd_struct_hit> d_pad;
            DevBuf<uint32_t> d_cnt;
            DevBuf<unsigned int> d_big;
            FD_CUDA(ctx, d_pad.alloc((size_t)nq * top_n));
            FD_CUDA(ctx, d_cnt.alloc(nq + 1));
            FD_CUDA(ctx, d_big.alloc(1));
            FD_CUDA(ctx, cudaMemsetAsync(d_big.p, 0, 4, s));
            unsigned int h_big = 0;
            {
                StageTimer st(ctx, "select");
                FD_LAUNCH(ctx, k3_topn_sort, nq, 256, 0, d_hits.p, d_hit_off.p, d_hit_cnt.p, (uint32_t)top_n,
                          opts.id_offset, d_pad.p, d_cnt.p, d_big.p);
                FD_CUDA(ctx, cudaMemcpyAsync(&h_big, d_big.p, 4, cudaMemcpyDeviceToHost, s));
                FD_CUDA(ctx, st.finish());
            }
            if (h_big && sharded) {
                // a pool too large for the sort kernel: the segmented-sort path on the host side, then back to the
                // padded device layout for the exchange
                fd_struct_hit *hh = nullptr;
                std::vector<uint64_t> ho(nq + 1, 0);
                FD_TRY(select_and_copy(ctx, nq, hit_off, d_hit_off, d_hit_cnt, d_hits, top_n, N, &hh, ho.data()));
                std::vector<fd_struct_hit> pad((size_t)nq * top_n);
                std::vector<uint32_t> cnt(nq);
                for (uint32_t q = 0; q < nq; q++) {
                    cnt[q] = (uint32_t)(ho[q + 1] - ho[q]);
                    for (uint32_t k = 0; k < cnt[q]; k++) {
                        pad[(size_t)q * top_n + k] = hh[ho[q] + k];
                        pad[(size_t)q * top_n + k].nid += opts.id_offset;
                    }
                }
                free(hh);
                FD_CUDA(ctx, cudaMemcpyAsync(d_pad.p, pad.data(), pad.size() * sizeof(fd_struct_hit), cudaMemcpyHostToDevice, s));
                FD_CUDA(ctx, cudaMemcpyAsync(d_cnt.p, cnt.data(), (size_t)nq * 4, cudaMemcpyHostToDevice, s));
                FD_CUDA(ctx, cudaStreamSynchronize(s));
                h_big = 0;
            }
            if (!h_big) {
                const fd_struct_hit *d_rows = d_pad.p;
                const uint32_t *d_rows_cnt = d_cnt.p;
                uint32_t n_out_q = nq;
                DevBuf<fd_struct_hit> d_recv, d_merged;
                DevBuf<uint32_t> d_rcnt, d_mcnt;
                if (sharded) {
                    // all-to-all of the fixed-size per-query blocks to the rank that owns the query, then the merge
                    const int world = fd_comm_world_of(ctx), rank = fd_comm_rank_of(ctx);
                    const uint32_t own0 = opts.slice_begin[rank], n_own = opts.slice_begin[rank + 1] - own0;
                    FD_CUDA(ctx, d_recv.alloc((size_t)world * n_own * top_n));
                    FD_CUDA(ctx, d_rcnt.alloc((size_t)world * n_own));
                    FD_CUDA(ctx, d_merged.alloc((size_t)n_own * top_n));
                    FD_CUDA(ctx, d_mcnt.alloc(n_own + 1));
                    std::vector<uint64_t> so(world), sb(world), ro(world), rb(world), so2(world), sb2(world), ro2(world), rb2(world);
                    for (int r = 0; r < world; r++) {
                        const uint64_t nr = opts.slice_begin[r + 1] - opts.slice_begin[r];
                        so[r] = (uint64_t)opts.slice_begin[r] * top_n * sizeof(fd_struct_hit);
                        sb[r] = nr * top_n * sizeof(fd_struct_hit);
                        ro[r] = (uint64_t)r * n_own * top_n * sizeof(fd_struct_hit);
                        rb[r] = (uint64_t)n_own * top_n * sizeof(fd_struct_hit);
                        so2[r] = (uint64_t)opts.slice_begin[r] * 4;
                        sb2[r] = nr * 4;
                        ro2[r] = (uint64_t)r * n_own * 4;
                        rb2[r] = (uint64_t)n_own * 4;
                    }
                    {
                        StageTimer st(ctx, "exchange");
                        FD_TRY(fd_comm_alltoallv_dev(ctx, (const uint8_t *)d_pad.p, so.data(), sb.data(), (uint8_t *)d_recv.p,
                                                     ro.data(), rb.data()));
                        FD_TRY(fd_comm_alltoallv_dev(ctx, (const uint8_t *)d_cnt.p, so2.data(), sb2.data(), (uint8_t *)d_rcnt.p,
                                                     ro2.data(), rb2.data()));
                        FD_CUDA(ctx, st.finish());
                    }
                    ctx->last_exchange_bytes = 0;
                    for (int r = 0; r < world; r++)
                        if (r != rank) ctx->last_exchange_bytes += sb[r] + sb2[r];
                    if (n_own) {
                        StageTimer st(ctx, "merge");
                        FD_LAUNCH(ctx, k3_merge_topn, n_own, 256, 0, d_recv.p, d_rcnt.p, (uint32_t)world, n_own, (uint32_t)top_n,
                                  d_merged.p, d_mcnt.p);
                        FD_CUDA(ctx, st.finish());
                    }
                    d_rows = d_merged.p;
                    d_rows_cnt = d_mcnt.p;
                    n_out_q = n_own;
                }
                void *h_pad_v = nullptr, *h_cnt_v = nullptr;
                FD_TRY(fd_pinned(ctx, 6, (size_t)n_out_q * top_n * sizeof(fd_struct_hit) + 16, &h_pad_v));
                FD_TRY(fd_pinned(ctx, 7, (size_t)n_out_q * 4 + 4, &h_cnt_v));
                fd_struct_hit *h_pad = (fd_struct_hit *)h_pad_v;
                uint32_t *h_cnt = (uint32_t *)h_cnt_v;
                {
                    StageTimer st(ctx, "select");
                    FD_CUDA(ctx, cudaMemcpyAsync(h_cnt, d_rows_cnt, (size_t)n_out_q * 4, cudaMemcpyDeviceToHost, s));
                    FD_CUDA(ctx, cudaMemcpyAsync(h_pad, d_rows, (size_t)n_out_q * top_n * sizeof(fd_struct_hit), cudaMemcpyDeviceToHost, s));
                    FD_CUDA(ctx, st.finish());
                }
                h_off[0] = 0;
                for (uint32_t q = 0; q < n_out_q; q++) h_off[q + 1] = h_off[q] + h_cnt[q];
                fd_struct_hit *h_hits = (fd_struct_hit *)malloc(std::max<uint64_t>(h_off[n_out_q], 1) * sizeof(fd_struct_hit));
                if (!h_hits) return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
                for (uint32_t q = 0; q < n_out_q; q++)
                    memcpy(h_hits + h_off[q], h_pad + (size_t)q * top_n, (size_t)h_cnt[q] * sizeof(fd_struct_hit));
                *out_hits = h_hits;
                return FD_OK;
            }
        }
        {
            const int rc2 = select_and_copy(ctx, nq, hit_off, d_hit_off, d_hit_cnt, d_hits, top_n, N, out_hits, h_off);
            if (rc2 == FD_OK && opts.id_offset)
                for (uint64_t k = 0; k < h_off[nq]; k++) (*out_hits)[k].nid += opts.id_offset;
            return rc2;
        }
    };
    // the pre-selection pays when the top n is a small part of a tile
    bool limit = !use_v1 && top_n > 0 && top_n <= 0xffffu && top_n * 8 <= N;
    if (const char *e = getenv("FD_K3_LIMIT")) limit = limit && atoi(e) != 0;
    int rc = attempt(limit);
    if (rc == 1) rc = attempt(false);
    if (rc != FD_OK) {
        free(h_off);
        return rc;
    }
    *out_offsets = h_off;
    return FD_OK;
}

int fd_count_query_batch(fd_ctx *ctx, const fd_query *queries, uint32_t nq, const fd_prefilter_params *params,
                         fd_struct_hit **out_hits, uint64_t **out_offsets) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->idx.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_count_query_batch: no index attached");
    if ((nq && !queries) || !params || !out_hits || !out_offsets)
        return fd_fail(ctx, FD_ERR_ARG, "fd_count_query_batch: NULL argument");
    FD_ENTER(ctx);
    return count_query_impl(ctx, queries, nq, params, CountOpts(), out_hits, out_offsets);
}

int fd_count_query_batch_id(fd_ctx *ctx, const fd_query *queries, uint32_t nq, const fd_prefilter_params *params,
                            uint64_t batch_id, fd_struct_hit **out_hits, uint64_t **out_offsets) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->idx.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_count_query_batch: no index attached");
    if ((nq && !queries) || !params || !out_hits || !out_offsets)
        return fd_fail(ctx, FD_ERR_ARG, "fd_count_query_batch: NULL argument");
    FD_ENTER(ctx);
    CountOpts o;
    o.batch_id = batch_id;
    return count_query_impl(ctx, queries, nq, params, o, out_hits, out_offsets);
}

int fd_count_query_batch_ex(fd_ctx *ctx, const fd_query *queries, uint32_t nq, const fd_prefilter_params *params,
                            const uint32_t *global_counts, uint64_t global_n_structs, fd_struct_hit **out_hits,
                            uint64_t **out_offsets) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->idx.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_count_query_batch_ex: no index attached");
    if ((nq && !queries) || !params || !out_hits || !out_offsets)
        return fd_fail(ctx, FD_ERR_ARG, "fd_count_query_batch_ex: NULL argument");
    if (global_counts && (global_n_structs == 0 || global_n_structs > 0xfffffff0ull))
        return fd_fail(ctx, FD_ERR_ARG, "fd_count_query_batch_ex: global_n_structs out of range");
    FD_ENTER(ctx);
    CountOpts o;
    o.gcounts = global_counts;
    o.g_structs = global_n_structs;
    return count_query_impl(ctx, queries, nq, params, o, out_hits, out_offsets);
}

int fd_count_query_sharded(fd_ctx *ctx, const fd_query *queries, uint32_t nq, const fd_prefilter_params *params,
                           const uint32_t *global_counts, uint64_t global_n_structs, uint64_t first_id,
                           const uint32_t *slice_begin, fd_struct_hit **out_hits, uint64_t **out_offsets) {
    return fd_count_query_sharded_id(ctx, queries, nq, params, global_counts, global_n_structs, first_id, slice_begin, 0,
                                     out_hits, out_offsets);
}

int fd_count_query_sharded_id(fd_ctx *ctx, const fd_query *queries, uint32_t nq, const fd_prefilter_params *params,
                              const uint32_t *global_counts, uint64_t global_n_structs, uint64_t first_id,
                              const uint32_t *slice_begin, uint64_t batch_id, fd_struct_hit **out_hits,
                              uint64_t **out_offsets) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->idx.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_count_query_sharded: no index attached");
    if (!ctx->comm) return fd_fail(ctx, FD_ERR_STATE, "fd_count_query_sharded: call fd_comm_init first");
    if ((nq && !queries) || !params || !out_hits || !out_offsets || !global_counts || !slice_begin)
        return fd_fail(ctx, FD_ERR_ARG, "fd_count_query_sharded: NULL argument");
    const int world = fd_comm_world_of(ctx);
    if (slice_begin[0] != 0 || slice_begin[world] != nq)
        return fd_fail(ctx, FD_ERR_ARG, "fd_count_query_sharded: slice_begin must run from 0 to n_queries");
    for (int r = 0; r < world; r++)
        if (slice_begin[r] > slice_begin[r + 1]) return fd_fail(ctx, FD_ERR_ARG, "fd_count_query_sharded: slices must be ascending");
    if (global_n_structs == 0 || global_n_structs > 0xfffffff0ull || first_id + ctx->idx.n_structs > global_n_structs)
        return fd_fail(ctx, FD_ERR_ARG, "fd_count_query_sharded: id range outside the database");
    if (params->top_n == 0 || params->top_n > 4096 / (uint64_t)world)
        return fd_fail(ctx, FD_ERR_LIMIT,
                       "the sharded search exchanges fixed-size per-query blocks: --top must be between 1 and 4096 / ranks");
    FD_ENTER(ctx);
    CountOpts o;
    o.gcounts = global_counts;
    o.g_structs = global_n_structs;
    o.id_offset = (uint32_t)first_id;
    o.slice_begin = slice_begin;
    o.batch_id = batch_id;
    return count_query_impl(ctx, queries, nq, params, o, out_hits, out_offsets);
}

uint64_t fd_last_exchange_bytes(const fd_ctx *ctx) { return ctx ? ctx->last_exchange_bytes : 0; }

int fd_votes_scan(fd_ctx *ctx, const fd_query *queries, uint32_t nq, const fd_prefilter_params *params,
                  fd_votes_layout *layout, uint32_t **d_votes) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->idx.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_votes_scan: no index attached");
    if ((nq && !queries) || !params || !layout || !d_votes) return fd_fail(ctx, FD_ERR_ARG, "fd_votes_scan: NULL argument");
    FD_ENTER(ctx);
    const uint32_t N = (uint32_t)ctx->idx.n_structs;
    Batch B;
    FD_TRY(prepare_batch(ctx, queries, nq, params, true, 1, B));
    const uint32_t planes = (B.narrow ? 1 : 2) + B.ew;
    *layout = fd_votes_layout{nq, N, B.narrow ? 1u : 0u, (uint32_t)B.ew, planes, 0, (uint64_t)planes * nq * N};
    if (layout->words > ctx->votes_cap) {
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->votes);
        ctx->votes = nullptr;
        ctx->votes_cap = 0;
        FD_CUDA(ctx, cudaMalloc(&ctx->votes, layout->words * 4));
        ctx->votes_cap = layout->words;
    }
    *d_votes = ctx->votes;
    if (nq == 0 || N == 0) return FD_OK;
    TilePlan tp;
    FD_TRY(plan_tiles(ctx, B, N, tp));
    StageTimer st(ctx, "scan");
    // every (query, tile) CTA writes its planes, so the buffer needs no clearing
    FD_TRY(launch_scan<1>(ctx, dim3(tp.n_tiles, nq), tp, make_view(ctx), B, make_filter(params), nullptr, nullptr,
                             nullptr, ctx->votes));
    FD_CUDA(ctx, st.finish());
    return FD_OK;
}

int fd_votes_scan_sparse(fd_ctx *ctx, const fd_query *queries, uint32_t nq, const fd_prefilter_params *params,
                         const uint32_t *slice_begin, uint32_t world, fd_votes_layout *layout, uint32_t **d_records,
                         uint64_t *region_offset, uint64_t *region_count) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->idx.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_votes_scan_sparse: no index attached");
    if ((nq && !queries) || !params || !layout || !d_records || !slice_begin || !region_offset || !region_count ||
        world == 0 || world > 64)
        return fd_fail(ctx, FD_ERR_ARG, "fd_votes_scan_sparse: bad argument");
    if (slice_begin[0] != 0 || slice_begin[world] != nq)
        return fd_fail(ctx, FD_ERR_ARG, "fd_votes_scan_sparse: slice_begin must run from 0 to n_queries");
    FD_ENTER(ctx);
    const uint32_t N = (uint32_t)ctx->idx.n_structs;
    for (uint32_t r = 0; r < world; r++) {
        if (slice_begin[r] > slice_begin[r + 1]) return fd_fail(ctx, FD_ERR_ARG, "fd_votes_scan_sparse: slices must be ascending");
        if ((uint64_t)(slice_begin[r + 1] - slice_begin[r]) * N > 0xffffffffull)
            return fd_fail(ctx, FD_ERR_LIMIT, "a rank's slice of the batch times the number of structures must stay below 2^32");
    }
    Batch B;
    FD_TRY(prepare_batch(ctx, queries, nq, params, true, 1, B));
    const uint32_t planes = (B.narrow ? 1 : 2) + B.ew;
    *layout = fd_votes_layout{nq, N, B.narrow ? 1u : 0u, (uint32_t)B.ew, planes, 0, 0};
    // a (query, structure) cell is non-empty only if a posting of the query names the structure
    region_offset[0] = 0;
    for (uint32_t r = 0; r < world; r++) {
        uint64_t cap = 0;
        for (uint32_t q = slice_begin[r]; q < slice_begin[r + 1]; q++) cap += std::min<uint64_t>(N, B.h_postings.empty() ? 0 : B.h_postings[q]);
        region_offset[r + 1] = region_offset[r] + cap;
        region_count[r] = 0;
    }
    const uint64_t words = region_offset[world] * (1 + planes);
    layout->words = words;
    if (words > ctx->votes_cap || !ctx->votes) { // never hand out NULL: a rank whose shard holds none of the batch's lists has 0 words
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->votes);
        ctx->votes = nullptr;
        ctx->votes_cap = 0;
        FD_CUDA(ctx, cudaMalloc(&ctx->votes, std::max<uint64_t>(words, 1) * 4));
        ctx->votes_cap = std::max<uint64_t>(words, 1);
    }
    *d_records = ctx->votes;
    if (nq == 0 || N == 0) return FD_OK;
    cudaStream_t s = ctx->stream;
    DevBuf<uint32_t> d_slice;
    DevBuf<uint64_t> d_off;
    DevBuf<unsigned long long> d_cnt;
    FD_CUDA(ctx, d_slice.alloc(world + 1));
    FD_CUDA(ctx, d_off.alloc(world + 1));
    FD_CUDA(ctx, d_cnt.alloc(world));
    FD_CUDA(ctx, cudaMemcpyAsync(d_slice.p, slice_begin, (world + 1) * 4, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_off.p, region_offset, (world + 1) * 8, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemsetAsync(d_cnt.p, 0, world * 8, s));
    std::vector<uint32_t> qorder;
    uint32_t class_begin[5];
    const int class_ew[4] = {1, 2, 4, 8};
    edge_word_classes(B, nq, qorder, class_begin);
    DevBuf<uint32_t> d_qorder;
    FD_CUDA(ctx, d_qorder.alloc(nq));
    FD_CUDA(ctx, cudaMemcpyAsync(d_qorder.p, qorder.data(), nq * 4, cudaMemcpyHostToDevice, s));
    StageTimer st(ctx, "scan");
    for (int c = 0; c < 4; c++) { // one launch per edge-word class; the records keep the batch's plane count
        const uint32_t nc = class_begin[c + 1] - class_begin[c];
        if (!nc) continue;
        const int ew = std::min(class_ew[c], B.ew);
        TilePlan tp;
        FD_TRY(plan_tiles(ctx, B, N, tp, ew));
        FD_TRY(launch_scan<2>(ctx, dim3(tp.n_tiles, nc), tp, make_view(ctx), B, make_filter(params), nullptr, nullptr,
                              nullptr, ctx->votes, SparseOut{d_slice.p, d_off.p, d_cnt.p, world, planes},
                              d_qorder.p + class_begin[c], ew));
    }
    FD_CUDA(ctx, cudaMemcpyAsync(region_count, d_cnt.p, world * 8, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, st.finish());
    return FD_OK;
}

int fd_votes_merge_begin(fd_ctx *ctx, const fd_votes_layout *layout, uint32_t first_query, uint32_t n_queries,
                         fd_votes_layout *slice_layout, uint32_t **d_dense) {
    if (!ctx || !layout || !slice_layout || !d_dense) return fd_fail(ctx, FD_ERR_ARG, "fd_votes_merge_begin: NULL argument");
    FD_ENTER(ctx);
    *slice_layout = *layout;
    slice_layout->n_queries = n_queries;
    slice_layout->first_query = first_query;
    slice_layout->words = (uint64_t)layout->planes * n_queries * layout->n_structs;
    const uint64_t words = std::max<uint64_t>(slice_layout->words, 1);
    if (words > ctx->merge_cap) {
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->merge);
        ctx->merge = nullptr;
        ctx->merge_cap = 0;
        FD_CUDA(ctx, cudaMalloc(&ctx->merge, words * 4));
        ctx->merge_cap = words;
    }
    StageTimer st(ctx, "merge");
    FD_CUDA(ctx, cudaMemsetAsync(ctx->merge, 0, words * 4, ctx->stream));
    FD_CUDA(ctx, st.finish());
    *d_dense = ctx->merge;
    return FD_OK;
}

int fd_votes_apply(fd_ctx *ctx, const fd_votes_layout *slice_layout, uint32_t *d_dense, const uint32_t *d_records,
                   uint64_t n_records) {
    if (!ctx || !slice_layout || (n_records && (!d_dense || !d_records)))
        return fd_fail(ctx, FD_ERR_ARG, "fd_votes_apply: NULL argument");
    FD_ENTER(ctx);
    if (n_records == 0) return FD_OK;
    StageTimer st(ctx, "merge");
    FD_LAUNCH(ctx, k3_apply_records, fd_div_up(n_records, 256), 256, 0, d_records, n_records, slice_layout->planes,
              slice_layout->narrow ? 1u : 2u, slice_layout->n_queries, slice_layout->n_structs, d_dense);
    FD_CUDA(ctx, st.finish());
    return FD_OK;
}

int fd_votes_select(fd_ctx *ctx, const fd_query *queries, uint32_t nq, const fd_prefilter_params *params,
                    const fd_votes_layout *layout, const uint32_t *d_votes, uint32_t q_begin, uint32_t q_end,
                    fd_struct_hit **out_hits, uint64_t **out_offsets) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->idx.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_votes_select: no index attached");
    if ((nq && !queries) || !params || !layout || !out_hits || !out_offsets || (layout->words && !d_votes))
        return fd_fail(ctx, FD_ERR_ARG, "fd_votes_select: NULL argument");
    if (q_begin > q_end || q_end > nq || q_begin < layout->first_query ||
        q_end > (uint64_t)layout->first_query + layout->n_queries || layout->n_structs != ctx->idx.n_structs)
        return fd_fail(ctx, FD_ERR_ARG, "fd_votes_select: layout / query range mismatch");
    FD_ENTER(ctx);
    *out_hits = nullptr;
    *out_offsets = nullptr;
    cudaStream_t s = ctx->stream;
    const uint32_t N = (uint32_t)ctx->idx.n_structs;
    const uint32_t nsel = q_end - q_begin;
    Batch B;
    FD_TRY(prepare_batch(ctx, queries, nq, params, false, 1, B));
    // the layout was fixed by the scan over the WHOLE batch; the queries given here may be a slice of it
    if ((layout->narrow && !B.narrow) || (uint32_t)B.ew > layout->edge_words ||
        layout->planes != (layout->narrow ? 1u : 2u) + layout->edge_words)
        return fd_fail(ctx, FD_ERR_ARG, "fd_votes_select: layout does not match the batch");
    B.narrow = layout->narrow != 0;
    B.ew = (int)layout->edge_words;
    uint64_t *h_off = (uint64_t *)calloc((size_t)nsel + 1, 8);
    if (!h_off) return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
    if (nsel == 0 || N == 0) {
        *out_offsets = h_off;
        *out_hits = (fd_struct_hit *)malloc(sizeof(fd_struct_hit));
        return FD_OK;
    }
    std::vector<uint64_t> hit_off(nsel + 1, 0);
    for (uint32_t q = 0; q < nsel; q++) hit_off[q + 1] = hit_off[q] + N;
    DevBuf<uint64_t> d_hit_off;
    DevBuf<unsigned int> d_hit_cnt;
    DevBuf<HitRec> d_hits;
    auto body = [&]() -> int {
        FD_CUDA(ctx, d_hit_off.alloc(nsel + 1));
        FD_CUDA(ctx, d_hit_cnt.alloc(nsel));
        FD_CUDA(ctx, d_hits.alloc(hit_off[nsel]));
        FD_CUDA(ctx, cudaMemcpyAsync(d_hit_off.p, hit_off.data(), (nsel + 1) * 8, cudaMemcpyHostToDevice, s));
        FD_CUDA(ctx, cudaMemsetAsync(d_hit_cnt.p, 0, nsel * 4, s));
        const uint32_t tile_ids = 8192;
        dim3 grid(fd_div_up(N, tile_ids), nsel);
        IndexView ix = make_view(ctx);
        FilterParams fp = make_filter(params);
        {
            StageTimer st(ctx, "select");
#define FD_SEL_CASE(NARROW, EW) \
    FD_TRY((launch_select_t<NARROW, EW>(ctx, grid, ix, B, d_votes, layout->n_queries, layout->first_query, q_begin, \
                                        tile_ids, fp, d_hit_off.p, d_hit_cnt.p, d_hits.p)))
            if (B.narrow) {
                if (B.ew == 1) FD_SEL_CASE(true, 1);
                else if (B.ew == 2) FD_SEL_CASE(true, 2);
                else if (B.ew == 4) FD_SEL_CASE(true, 4);
                else FD_SEL_CASE(true, 8);
            } else {
                if (B.ew == 1) FD_SEL_CASE(false, 1);
                else if (B.ew == 2) FD_SEL_CASE(false, 2);
                else if (B.ew == 4) FD_SEL_CASE(false, 4);
                else FD_SEL_CASE(false, 8);
            }
#undef FD_SEL_CASE
            FD_CUDA(ctx, st.finish());
        }
        return select_and_copy(ctx, nsel, hit_off, d_hit_off, d_hit_cnt, d_hits, params->top_n, N, out_hits, h_off);
    };
    const int rc = body();
    if (rc != FD_OK) {
        free(h_off);
        return rc;
    }
    *out_offsets = h_off;
    return FD_OK;
}

} // extern "C"
