// fd_query.cu -- K3: posting-list scan + per-structure vote (query path (ii), prefilter stage).
//
// Replaces count_query (reference src/controller/count_query.rs:82-220), FolddiscoIndex::get_entries
// (src/index/indextable.rs:44-86, 439-463), StructureFilter::filter_before_matching (src/controller/filter.rs:
// 76-100) and the idf sort + --top truncation of src/cli/workflows/query_pdb.rs:395-411, for a batch of queries.
//
// HBM layout of the attached index (fd_index_attach):
//   hashes[count] u32 ascending | offsets[count+1] u64 | values[value_bytes] raw delta+LEB128 bytes (the file,
//   unmodified) | counts[count] u32 postings per list | dir[2^20+1] u32 bucket directory over hash>>12 |
//   skip_pos/skip_id[value_bytes/256+1]: for every 256-byte boundary of values[] the first varint start at or
//   after it and the running structure id of its list at that point, so a list can be entered at any 256 B
//   segment without decoding its prefix (the LEB128 stream is otherwise strictly sequential).
//
// Kernels per batch:
//   k3_lookup : one thread per query hash: directory + binary search -> byte range, posting count, idf weight
//   k3_scan   : one CTA per (structure-id tile, query).  The tile's votes live in shared memory
//               (packed match_count|idf fixed point, plus an edge bitmask per structure).  Work items are the
//               256-byte segments of the query's lists that intersect the tile; each lane decodes one segment
//               and votes with shared-memory atomics.  The epilogue derives node/edge counts, applies the
//               length penalty and the structure filter, and appends survivors to the query's hit region.
//   (cub segmented sort) + k3_gather: order hits by (idf desc, nid asc) and keep top_n.
// Only posting bytes, skip entries and survivors touch HBM; the N-sized per-node-group arrays and O(N * edges)
// bit sweeps of the reference never exist.
#include <cub/cub.cuh>

#include <algorithm>

#include "fd_common.cuh"

void fd_ctx_release_index(fd_ctx *ctx);

namespace {

constexpr int DIR_SHIFT = 12;
constexpr uint32_t DIR_SIZE = 1u << (32 - DIR_SHIFT);
constexpr int SKIP_SHIFT = 8; // 256-byte segments
constexpr uint32_t SKIP_BYTES = 1u << SKIP_SHIFT;

constexpr int K3_THREADS = 256;
constexpr int K3_MAX_HASHES_NARROW = 255; // match_count fits 8 bits
constexpr int K3_MAX_HASHES = 4095;       // smem prefix array
constexpr int K3_MAX_EDGE_WORDS = 8;      // 256 edges
constexpr int K3_MAX_NODES = 256;

// ------------------------------------------------------------------------------------------------
// attach-time kernels
// ------------------------------------------------------------------------------------------------

__global__ void k3_build_dir(const uint32_t *hashes, uint64_t count, uint32_t *dir) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > count) return;
    // dir[b] = first position whose bucket >= b
    const uint32_t lo = k == 0 ? 0u : (hashes[k - 1] >> DIR_SHIFT) + 1u;
    const uint32_t hi = k == count ? DIR_SIZE : (hashes[k] >> DIR_SHIFT);
    for (uint32_t b = lo; b <= hi && b <= DIR_SIZE; b++) dir[b] = (uint32_t)k;
}

// number of varint terminators (bytes with the top bit clear) per 256-byte block
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

// first varint start >= p inside the list that begins at lstart (p > lstart)
__device__ __forceinline__ uint64_t varint_start_at_or_after(const uint8_t *values, uint64_t p) {
    while (values[p - 1] & 0x80u) p++;
    return p;
}

// For every 256-byte boundary b: skip_pos[b] and a (reset flag, partial id sum) pair whose segmented inclusive
// scan is the running structure id at skip_pos[b].
__global__ void k3_skip_partials(const uint8_t *values, const uint64_t *offsets, uint64_t count, uint64_t nbytes,
                                 uint64_t nblocks, uint32_t *skip_pos, uint64_t *partial) {
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const uint64_t p0 = b << SKIP_SHIFT;
    if (p0 >= nbytes) { // pad entry
        skip_pos[b] = (uint32_t)min(nbytes, (uint64_t)0xffffffffu);
        partial[b] = 1ull << 32;
        return;
    }
    const uint64_t l = list_containing(offsets, count, p0);
    const uint64_t lstart = offsets[l];
    if (lstart == p0) {
        skip_pos[b] = (uint32_t)p0;
        partial[b] = 1ull << 32; // reset, running id 0
        return;
    }
    const uint64_t pos = varint_start_at_or_after(values, p0);
    skip_pos[b] = (uint32_t)pos;
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
    const uint32_t *skip_pos;
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
};

__global__ void k3_lookup(IndexView ix, const uint32_t *qhashes, uint32_t n_qhashes, float freq_filter, QHash *out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_qhashes) return;
    QHash r{0, 0, 0, 0.f};
    const int64_t l = find_list(ix, qhashes[k]);
    if (l >= 0) {
        const uint32_t c = ix.counts[l];
        bool keep = c > 0;
        // count_query.rs:124-128: skip hashes more frequent than freq_filter
        if (keep && freq_filter >= 0.f && (float)c / (float)ix.n_structs > freq_filter) keep = false;
        if (keep) {
            r.start = ix.offsets[l];
            r.end = ix.offsets[l + 1];
            r.count = c;
            r.idf = log2f((float)ix.n_structs / (float)c);
        }
    }
    out[k] = r;
}

// per-query totals in a fixed order (deterministic): postings, posting bytes, sum of idf weights
__global__ void k3_query_sums(const QueryDesc *queries, const QHash *qh, uint32_t n_queries,
                              unsigned long long *postings, unsigned long long *bytes, float *idf_sum) {
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
    idf_sum[q] = s;
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

struct HitRec { // 16 bytes; key/value for the segmented sort are derived from it
    uint32_t nid;
    uint32_t match_count;
    uint32_t node_edge; // node_count << 16 | edge_count
    float idf;
};

struct FilterParams {
    float length_penalty;
    uint32_t total_match_count, covered_node_count;
    float covered_node_ratio, idf_score_cutoff;
    uint32_t num_res_cutoff;
    float plddt_cutoff;
};

// Shared-memory vote tile.  NARROW: one word = match_count (8 bits) | idf fixed point (24 bits);
// wide: separate match and idf words.  EW edge-bitmask words per structure.
template <bool NARROW, int EW>
__global__ void __launch_bounds__(K3_THREADS)
    k3_scan(IndexView ix, const QueryDesc *queries, const QHash *qh, const uint16_t *edge_of_hash,
            const uint16_t *edge_node, const float *idf_sum_per_query, uint32_t tile_ids, FilterParams fp,
            const uint64_t *hit_offsets, unsigned int *hit_counts, HitRec *hits) {
    extern __shared__ uint32_t smem[];
    const uint32_t q = blockIdx.y;
    const QueryDesc qd = queries[q];
    const uint32_t lo = blockIdx.x * tile_ids;
    const uint32_t hi = min(ix.n_structs, lo + tile_ids);
    const uint32_t T = hi - lo;
    const bool single_tile = gridDim.x == 1;
    const uint32_t Q = qd.n_hashes;

    // smem carve-up
    uint32_t *w_acc = smem;                                  // [tile_ids] narrow: packed; wide: idf fixed
    uint32_t *w_match = NARROW ? nullptr : w_acc + tile_ids; // [tile_ids] wide only
    uint32_t *w_edge = w_acc + (NARROW ? 1 : 2) * tile_ids;  // [tile_ids * EW]
    uint32_t *item_prefix = w_edge + (size_t)tile_ids * EW;  // [Q + 1]
    uint32_t *seg_lo = item_prefix + (Q + 1);                // [Q] first relevant segment of each list
    uint32_t *node_mask = seg_lo + Q;                        // [n_nodes * EW]
    __shared__ uint32_t s_total_items;

    for (uint32_t i = threadIdx.x; i < T * (NARROW ? 1u : 2u); i += K3_THREADS) w_acc[i] = 0;
    if (NARROW) {
        for (uint32_t i = threadIdx.x; i < T * EW; i += K3_THREADS) w_edge[i] = 0;
    } else {
        for (uint32_t i = threadIdx.x; i < T * EW; i += K3_THREADS) w_edge[i] = 0;
    }
    for (uint32_t i = threadIdx.x; i < qd.n_nodes * EW; i += K3_THREADS) node_mask[i] = 0;
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < qd.n_edges; e += K3_THREADS)
        atomicOr(&node_mask[edge_node[qd.edge_begin + e] * EW + (e >> 5)], 1u << (e & 31));

    // fixed-point scale for idf: the sum over all of the query's hashes must fit the accumulator
    const float idf_total = idf_sum_per_query[q] + 1.0f;
    const float scale = exp2f(floorf(log2f((NARROW ? 16777215.0f : 4294967040.0f) / idf_total)));
    const float inv_scale = 1.0f / scale;

    // ---- which 256-byte segments of each list intersect this tile ----
    for (uint32_t k = threadIdx.x; k < Q; k += K3_THREADS) {
        const QHash h = qh[qd.hash_begin + k];
        uint32_t n_items = 0, first = 0;
        if (h.end > h.start) {
            const uint64_t b0 = h.start >> SKIP_SHIFT, b1 = (h.end - 1) >> SKIP_SHIFT;
            const uint32_t nseg = (uint32_t)(b1 - b0) + 1;
            if (single_tile || nseg == 1) {
                first = 0;
                n_items = nseg;
            } else {
                // r(k') = skip_id[b0 + k'] for k' in [1, nseg-1], non-decreasing
                const uint32_t *r = ix.skip_id + b0;
                uint32_t a = 1, b = nseg; // count of k' with r(k') < lo
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
                first = below_lo;                 // segment below_lo may still hold ids >= lo
                n_items = below_hi + 1 - below_lo; // segments [below_lo, below_hi]
            }
        }
        seg_lo[k] = first;
        item_prefix[k + 1] = n_items;
    }
    if (threadIdx.x == 0) item_prefix[0] = 0;
    __syncthreads();
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

    // ---- decode + vote ----
    for (uint32_t it = threadIdx.x; it < total_items; it += K3_THREADS) {
        // list index: last k with item_prefix[k] <= it
        uint32_t a = 0, b = Q;
        while (b - a > 1) {
            uint32_t m = (a + b) >> 1;
            if (item_prefix[m] <= it) a = m;
            else b = m;
        }
        const uint32_t k = a;
        const QHash h = qh[qd.hash_begin + k];
        const uint32_t seg = seg_lo[k] + (it - item_prefix[k]);
        const uint64_t b0 = h.start >> SKIP_SHIFT;
        const uint64_t last_block = (h.end - 1) >> SKIP_SHIFT;
        uint64_t p = seg == 0 ? h.start : (uint64_t)ix.skip_pos[b0 + seg];
        uint32_t id = seg == 0 ? 0u : ix.skip_id[b0 + seg];
        const uint64_t pend = (b0 + seg) < last_block ? (uint64_t)ix.skip_pos[b0 + seg + 1] : h.end;
        const uint32_t e = edge_of_hash[qd.hash_begin + k];
        const uint32_t ebit = 1u << (e & 31), eword = e >> 5;
        const uint32_t w = (uint32_t)(fmaxf(h.idf, 0.f) * scale + 0.5f);
        const uint32_t add = NARROW ? ((1u << 24) | w) : w;
        uint32_t cur = 0;
        int shift = 0;
        for (; p < pend; p++) {
            const uint32_t v = ix.values[p];
            cur |= (v & 0x7Fu) << shift;
            if (v & 0x80u) {
                shift += 7;
                continue;
            }
            id += cur;
            cur = 0;
            shift = 0;
            if (id >= hi) break;
            if (id >= lo) {
                const uint32_t x = id - lo;
                atomicAdd(&w_acc[x], add);
                if (!NARROW) atomicAdd(&w_match[x], 1u);
                atomicOr(&w_edge[x * EW + eword], ebit);
            }
        }
    }
    __syncthreads();

    // ---- epilogue: counts, length penalty, filter, append ----
    const uint64_t out_base = hit_offsets[q];
    for (uint32_t x0 = 0; x0 < T; x0 += K3_THREADS) {
        const uint32_t x = x0 + threadIdx.x;
        bool pass = false;
        HitRec rec{0, 0, 0, 0.f};
        if (x < T) {
            const uint32_t acc = w_acc[x];
            const uint32_t mc = NARROW ? (acc >> 24) : w_match[x];
            if (mc > 0) {
                const uint32_t fixed = NARROW ? (acc & 0xffffffu) : acc;
                uint32_t ew[EW];
                uint32_t ec = 0;
#pragma unroll
                for (int i = 0; i < EW; i++) {
                    ew[i] = w_edge[x * EW + i];
                    ec += __popc(ew[i]);
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
                // count_query.rs:199: idf_sum *= (nres as f32).powf(-lp)
                const float idf = ((float)fixed * inv_scale) * powf((float)nr, -fp.length_penalty);
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
        }
        const uint32_t m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            const int lane = threadIdx.x & 31;
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(&hit_counts[q], (unsigned int)__popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1));
            if (pass) hits[out_base + pos] = rec;
        }
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

IndexView make_view(const fd_ctx *ctx) {
    const FdDeviceIndex &d = ctx->idx;
    return IndexView{d.hashes, d.offsets, d.values, d.counts, d.dir, d.skip_pos, d.skip_id, d.nres, d.plddt,
                     d.count, (uint32_t)d.n_structs};
}

template <bool NARROW, int EW>
int launch_scan(fd_ctx *ctx, dim3 grid, size_t smem, IndexView ix, const QueryDesc *queries, const QHash *qh,
                const uint16_t *edge_of_hash, const uint16_t *edge_node, const float *idf_sum, uint32_t tile_ids,
                FilterParams fp, const uint64_t *hit_offsets, unsigned int *hit_counts, HitRec *hits) {
    auto kern = k3_scan<NARROW, EW>;
    FD_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, K3_THREADS, smem, ctx->stream>>>(ix, queries, qh, edge_of_hash, edge_node, idf_sum, tile_ids, fp,
                                                  hit_offsets, hit_counts, hits);
    ctx->launches++;
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
    if (value_bytes > 0xfffffff0ull)
        return fd_fail(ctx, FD_ERR_LIMIT, "posting bytes per device must stay below 4 GiB in this version; shard the index");
    if (n_structs > 0xfffffff0ull) return fd_fail(ctx, FD_ERR_LIMIT, "structure ids must fit in 32 bits");
    if (offsets[0] != 0 || offsets[count] != value_bytes)
        return fd_fail(ctx, FD_ERR_ARG, "fd_index_attach: offsets[0] must be 0 and offsets[count] == value_bytes");
    FD_ENTER(ctx);
    fd_ctx_release_index(ctx);
    FdDeviceIndex &d = ctx->idx;
    cudaStream_t s = ctx->stream;
    const uint64_t nblocks = (value_bytes >> SKIP_SHIFT) + 2;
    FD_CUDA(ctx, cudaMalloc(&d.hashes, std::max<uint64_t>(count, 1) * 4));
    FD_CUDA(ctx, cudaMalloc(&d.offsets, (count + 1) * 8));
    FD_CUDA(ctx, cudaMalloc(&d.values, value_bytes + 16));
    FD_CUDA(ctx, cudaMalloc(&d.counts, std::max<uint64_t>(count, 1) * 4));
    FD_CUDA(ctx, cudaMalloc(&d.dir, ((uint64_t)DIR_SIZE + 2) * 4));
    FD_CUDA(ctx, cudaMalloc(&d.skip_pos, nblocks * 4));
    FD_CUDA(ctx, cudaMalloc(&d.skip_id, nblocks * 4));
    FD_CUDA(ctx, cudaMalloc(&d.nres, std::max<uint64_t>(n_structs, 1) * 4));
    FD_CUDA(ctx, cudaMalloc(&d.plddt, std::max<uint64_t>(n_structs, 1) * 4));
    d.count = count;
    d.value_bytes = value_bytes;
    d.n_structs = n_structs;
    d.n_skip = nblocks;
    FD_CUDA(ctx, cudaMemcpyAsync(d.hashes, hashes, count * 4, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d.offsets, offsets, (count + 1) * 8, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d.values, values, value_bytes, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemsetAsync(d.values + value_bytes, 0, 16, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d.nres, nres, n_structs * 4, cudaMemcpyHostToDevice, s));
    if (plddt) FD_CUDA(ctx, cudaMemcpyAsync(d.plddt, plddt, n_structs * 4, cudaMemcpyHostToDevice, s));
    else FD_CUDA(ctx, cudaMemsetAsync(d.plddt, 0, std::max<uint64_t>(n_structs, 1) * 4, s));

    StageTimer st(ctx, "attach");
    FD_LAUNCH(ctx, k3_build_dir, fd_div_up(count + 1, 256), 256, 0, d.hashes, count, d.dir);
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
                  nblocks, d.skip_pos, partial.p);
        tb = tb2;
        FD_CUDA(ctx, cub::DeviceScan::InclusiveScan(tmp.p, tb, partial.p, scanned.p, SegScanOp(), nblocks, s));
        ctx->launches += 2;
        FD_LAUNCH(ctx, k3_low32, fd_div_up(nblocks, 256), 256, 0, scanned.p, nblocks, d.skip_id);
    }
    FD_CUDA(ctx, st.finish());
    d.attached = true;
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

int fd_count_query_batch(fd_ctx *ctx, const fd_query *queries, uint32_t nq, const fd_prefilter_params *params,
                         fd_struct_hit **out_hits, uint64_t **out_offsets) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->idx.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_count_query_batch: no index attached");
    if ((nq && !queries) || !params || !out_hits || !out_offsets)
        return fd_fail(ctx, FD_ERR_ARG, "fd_count_query_batch: NULL argument");
    FD_ENTER(ctx);
    *out_hits = nullptr;
    *out_offsets = nullptr;
    cudaStream_t s = ctx->stream;
    const uint32_t N = (uint32_t)ctx->idx.n_structs;

    // ---- flatten the batch on the host, applying sample_query (count_query.rs:222-253) if requested ----
    const bool has_r = params->sampling_ratio >= 0.f, has_c = params->sampling_count >= 0;
    const bool sampling = has_r != has_c;
    std::vector<uint32_t> f_hash, f_query;
    std::vector<uint16_t> f_edge, f_edge_node;
    std::vector<QueryDesc> descs(nq);
    uint32_t max_hashes = 0, max_edges = 0, max_nodes = 0;
    std::vector<uint32_t> sample_counts;
    if (sampling) {
        std::vector<uint32_t> all;
        for (uint32_t q = 0; q < nq; q++) all.insert(all.end(), queries[q].hashes, queries[q].hashes + queries[q].n_hashes);
        sample_counts.resize(all.size());
        FD_TRY(fd_posting_counts(ctx, all.data(), all.size(), sample_counts.data()));
    }
    size_t sample_base = 0;
    for (uint32_t q = 0; q < nq; q++) {
        const fd_query &Q = queries[q];
        if (Q.n_hashes && (!Q.hashes || !Q.edge_of_hash)) return fd_fail(ctx, FD_ERR_ARG, "fd_query: NULL array");
        if (Q.n_edges && !Q.edge_node) return fd_fail(ctx, FD_ERR_ARG, "fd_query: NULL edge_node");
        std::vector<uint32_t> order(Q.n_hashes);
        for (uint32_t k = 0; k < Q.n_hashes; k++) order[k] = k;
        if (sampling) {
            const uint32_t *cnt = sample_counts.data() + sample_base;
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cnt[a] < cnt[b]; });
            size_t keep = has_r ? (size_t)std::ceil(params->sampling_ratio * (float)Q.n_hashes)
                                : (size_t)params->sampling_count;
            if (keep < order.size()) order.resize(keep);
            sample_base += Q.n_hashes;
        }
        descs[q] = QueryDesc{(uint32_t)f_hash.size(), (uint32_t)order.size(), (uint32_t)f_edge_node.size(),
                             Q.n_edges, Q.n_nodes, Q.expected_node_count};
        for (uint32_t k : order) {
            if (Q.edge_of_hash[k] >= Q.n_edges) return fd_fail(ctx, FD_ERR_ARG, "fd_query: edge_of_hash out of range");
            f_hash.push_back(Q.hashes[k]);
            f_query.push_back(q);
            f_edge.push_back(Q.edge_of_hash[k]);
        }
        for (uint32_t e = 0; e < Q.n_edges; e++) {
            if (Q.edge_node[e] >= Q.n_nodes) return fd_fail(ctx, FD_ERR_ARG, "fd_query: edge_node out of range");
            f_edge_node.push_back(Q.edge_node[e]);
        }
        max_hashes = std::max<uint32_t>(max_hashes, (uint32_t)order.size());
        max_edges = std::max(max_edges, Q.n_edges);
        max_nodes = std::max(max_nodes, Q.n_nodes);
    }
    if (max_hashes > K3_MAX_HASHES || max_edges > 32 * K3_MAX_EDGE_WORDS || max_nodes > K3_MAX_NODES)
        return fd_fail(ctx, FD_ERR_LIMIT,
                       "query too large for the shared-memory vote kernel (limits: 4095 hashes, 256 edges, 256 "
                       "nodes per query); whole-structure queries are not supported in this version");
    uint64_t *h_off = (uint64_t *)calloc((size_t)nq + 1, 8);
    if (!h_off) return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
    ctx->last_posting_bytes = 0;
    if (nq == 0 || N == 0 || f_hash.empty()) {
        *out_offsets = h_off;
        *out_hits = (fd_struct_hit *)malloc(sizeof(fd_struct_hit));
        return FD_OK;
    }
    const uint32_t nqh = (uint32_t)f_hash.size();

    DevBuf<uint32_t> d_hash;
    DevBuf<uint16_t> d_edge, d_edge_node;
    DevBuf<QueryDesc> d_desc;
    DevBuf<QHash> d_qh;
    DevBuf<unsigned long long> d_postings, d_bytes;
    DevBuf<float> d_idfsum;
    FD_CUDA(ctx, d_hash.alloc(nqh));
    FD_CUDA(ctx, d_edge.alloc(nqh));
    FD_CUDA(ctx, d_edge_node.alloc(f_edge_node.size()));
    FD_CUDA(ctx, d_desc.alloc(nq));
    FD_CUDA(ctx, d_qh.alloc(nqh));
    FD_CUDA(ctx, d_postings.alloc(nq));
    FD_CUDA(ctx, d_bytes.alloc(nq));
    FD_CUDA(ctx, d_idfsum.alloc(nq));
    FD_CUDA(ctx, cudaMemcpyAsync(d_hash.p, f_hash.data(), nqh * 4, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_edge.p, f_edge.data(), nqh * 2, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_edge_node.p, f_edge_node.data(), f_edge_node.size() * 2, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_desc.p, descs.data(), nq * sizeof(QueryDesc), cudaMemcpyHostToDevice, s));

    IndexView ix = make_view(ctx);
    std::vector<unsigned long long> h_postings(nq);
    {
        StageTimer st(ctx, "lookup");
        FD_LAUNCH(ctx, k3_lookup, fd_div_up(nqh, 256), 256, 0, ix, d_hash.p, nqh, params->freq_filter, d_qh.p);
        FD_LAUNCH(ctx, k3_query_sums, fd_div_up(nq, 128), 128, 0, d_desc.p, d_qh.p, nq, d_postings.p, d_bytes.p,
                  d_idfsum.p);
        std::vector<unsigned long long> h_bytes(nq);
        FD_CUDA(ctx, cudaMemcpyAsync(h_postings.data(), d_postings.p, nq * 8, cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, cudaMemcpyAsync(h_bytes.data(), d_bytes.p, nq * 8, cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, st.finish());
        for (unsigned long long b : h_bytes) ctx->last_posting_bytes += b;
    }
    // hit regions: a query can hit at most min(N, its postings) structures
    std::vector<uint64_t> hit_off(nq + 1, 0);
    for (uint32_t q = 0; q < nq; q++) hit_off[q + 1] = hit_off[q] + std::min<uint64_t>(N, h_postings[q]);
    const uint64_t pool = hit_off[nq];
    DevBuf<uint64_t> d_hit_off, d_seg_end, d_keys, d_keys2, d_out_off;
    DevBuf<unsigned int> d_hit_cnt;
    DevBuf<uint32_t> d_vals, d_vals2;
    DevBuf<HitRec> d_hits;
    FD_CUDA(ctx, d_hit_off.alloc(nq + 1));
    FD_CUDA(ctx, d_hit_cnt.alloc(nq));
    FD_CUDA(ctx, d_hits.alloc(pool));
    FD_CUDA(ctx, cudaMemcpyAsync(d_hit_off.p, hit_off.data(), (nq + 1) * 8, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemsetAsync(d_hit_cnt.p, 0, nq * 4, s));

    // ---- scan + vote ----
    const bool narrow = max_hashes <= K3_MAX_HASHES_NARROW;
    const int ew = max_edges <= 32 ? 1 : (max_edges <= 64 ? 2 : (max_edges <= 128 ? 4 : 8));
    const uint32_t bytes_per_id = (narrow ? 4 : 8) + 4 * ew;
    const size_t fixed_smem = (size_t)(2 * max_hashes + 2 + max_nodes * ew) * 4;
    const size_t budget = 64 * 1024; // ~3 CTAs per SM
    uint32_t tile_ids = (uint32_t)((budget - std::min(budget / 2, fixed_smem)) / bytes_per_id);
    tile_ids = std::max<uint32_t>(256, tile_ids & ~31u);
    if (tile_ids > N) tile_ids = (N + 31) & ~31u;
    const uint32_t n_tiles = fd_div_up(N, tile_ids);
    const size_t smem = (size_t)tile_ids * bytes_per_id + fixed_smem + 64;
    if (smem > 220 * 1024) return fd_fail(ctx, FD_ERR_LIMIT, "query needs more shared memory than one SM has");
    FilterParams fp{params->length_penalty,
                    (uint32_t)std::min<uint64_t>(params->total_match_count, 0xffffffffu),
                    (uint32_t)std::min<uint64_t>(params->covered_node_count, 0xffffffffu),
                    params->covered_node_ratio,
                    params->idf_score_cutoff,
                    (uint32_t)std::min<uint64_t>(params->num_res_cutoff, 0xffffffffu),
                    params->plddt_cutoff};
    {
        StageTimer st(ctx, "scan");
        dim3 grid(n_tiles, nq);
        int rc;
#define FD_SCAN_CASE(NARROW, EW)                                                                              \
    rc = launch_scan<NARROW, EW>(ctx, grid, smem, ix, d_desc.p, d_qh.p, d_edge.p, d_edge_node.p, d_idfsum.p, \
                                 tile_ids, fp, d_hit_off.p, d_hit_cnt.p, d_hits.p)
        if (narrow) {
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
        FD_TRY(rc);
        FD_CUDA(ctx, st.finish());
    }

    // ---- order by (idf desc, nid asc), keep top_n ----
    std::vector<unsigned int> h_cnt(nq);
    fd_struct_hit *h_hits = nullptr;
    {
        StageTimer st(ctx, "select");
        FD_CUDA(ctx, cudaMemcpyAsync(h_cnt.data(), d_hit_cnt.p, nq * 4, cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, d_keys.alloc(pool));
        FD_CUDA(ctx, d_keys2.alloc(pool));
        FD_CUDA(ctx, d_vals.alloc(pool));
        FD_CUDA(ctx, d_vals2.alloc(pool));
        FD_CUDA(ctx, d_seg_end.alloc(nq));
        dim3 g2(std::max<uint32_t>(1, std::min<uint32_t>(64, fd_div_up(N, 256))), nq);
        FD_LAUNCH(ctx, k3_make_sort_keys, g2, 256, 0, d_hits.p, d_hit_off.p, d_hit_cnt.p, nq, d_keys.p, d_vals.p);
        FD_LAUNCH(ctx, k3_segment_ends, fd_div_up(nq, 256), 256, 0, d_hit_off.p, d_hit_cnt.p, nq, d_seg_end.p);
        size_t tb = 0;
        cub::DeviceSegmentedSort::SortPairs(nullptr, tb, d_keys.p, d_keys2.p, d_vals.p, d_vals2.p, (int64_t)pool,
                                            (int64_t)nq, d_hit_off.p, d_seg_end.p, s);
        DevBuf<uint8_t> tmp;
        FD_CUDA(ctx, tmp.alloc(tb));
        if (pool)
            FD_CUDA(ctx, cub::DeviceSegmentedSort::SortPairs(tmp.p, tb, d_keys.p, d_keys2.p, d_vals.p, d_vals2.p,
                                                             (int64_t)pool, (int64_t)nq, d_hit_off.p, d_seg_end.p, s));
        ctx->launches += 3;
        FD_CUDA(ctx, cudaStreamSynchronize(s));
        for (uint32_t q = 0; q < nq; q++)
            h_off[q + 1] = h_off[q] + std::min<uint64_t>(h_cnt[q], params->top_n);
        const uint64_t n_out = h_off[nq];
        DevBuf<fd_struct_hit> d_out;
        FD_CUDA(ctx, d_out.alloc(n_out));
        FD_CUDA(ctx, d_out_off.alloc(nq + 1));
        FD_CUDA(ctx, cudaMemcpyAsync(d_out_off.p, h_off, (nq + 1) * 8, cudaMemcpyHostToDevice, s));
        if (n_out) FD_LAUNCH(ctx, k3_gather, g2, 256, 0, d_hits.p, d_vals2.p, d_hit_off.p, d_out_off.p, d_out.p);
        h_hits = (fd_struct_hit *)malloc(std::max<uint64_t>(n_out, 1) * sizeof(fd_struct_hit));
        if (!h_hits) {
            free(h_off);
            return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
        }
        FD_CUDA(ctx, cudaMemcpyAsync(h_hits, d_out.p, n_out * sizeof(fd_struct_hit), cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, st.finish());
    }
    *out_hits = h_hits;
    *out_offsets = h_off;
    return FD_OK;
}

} // extern "C"
