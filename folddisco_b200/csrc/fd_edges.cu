// fd_edges.cu -- K4: candidate edge extraction (query path (ii), verification stage) + the HBM structure store.
//
// Replaces, per (query, candidate structure) pair, prefilter_amino_acid and retrieve_with_prefilter
// (reference src/controller/retrieve.rs:563-602, 52-156).  The reference re-reads and re-parses the candidate's
// file from disk for every query (retrieve.rs:375-376); here the compact structures live in HBM
// (fd_store_attach: N/CA/CB/aa SoA, ~37 B per residue) and are re-hashed in place.
//
// One CTA per candidate:
//   1. residues whose amino acid occurs as res1 / res2 of any query hash are compacted into two shared-memory
//      index lists (the BTreeSets of retrieve.rs:566-601; exact residue-name match, so modified residues are
//      left out exactly like the reference does);
//   2. the |set1| x |set2| pairs are screened with the cheap tests (CA distance <= cutoff, amino-acid pair seen in
//      the query, |d - d_query| < --ca-distance) and survivors are queued in shared memory;
//   3. survivors are drained densely: full feature + hash (fd_geom.cuh), membership in the query hash set
//      (binary search), output of (cand, i, j, hash) edges and of the (q_index, i, j) candidate pairs that the
//      residue-rescue step needs.
// Two passes (count, emit) size the outputs exactly; a radix sort on (cand, i, j[, k]) restores the reference's
// emission order, which fixes graph node numbering and the f32 summation order downstream.
#include <cub/cub.cuh>

#include <algorithm>

#include "fd_common.cuh"
#include "fd_geom.cuh"
#include "fd_hashtypes.cuh"

void fd_ctx_release_store(fd_ctx *ctx);

namespace {

constexpr int K4_THREADS = 128;
constexpr int K4_LIST_CAP = 4096;
constexpr int K4_CHUNK = K4_THREADS * 8;
constexpr int K4_MAX_AADIST = 512;
constexpr uint32_t PREFILTER_AA_SKIPPING_SIZE = 200; // retrieve.rs:24

struct StoreView {
    const uint64_t *row_offsets;
    const float *n_xyz, *ca_xyz, *cb_xyz;
    const uint8_t *aa, *cb_valid;
};

struct RQDesc { // per query
    uint32_t hash_begin, n_hashes;
    uint32_t aad_begin, n_aad;
    uint32_t aa1_mask, aa2_mask;
    uint32_t use_prefilter;
    // queries with more than K4_MAX_AADIST observed pairs (whole-structure queries, query.rs:226-233): their entries stay
    // in global memory grouped by amino-acid pair, run_begin indexes the 401-entry run table; 0xFFFFFFFF otherwise
    uint32_t run_begin;
};

struct AADist {
    uint8_t aa1, aa2;
    uint16_t k; // position inside its (aa1, aa2) list
    float dist;
    uint32_t q_index;
};

__device__ __forceinline__ fdg::V3 ld3(const float *p, uint64_t r) { return {p[3 * r], p[3 * r + 1], p[3 * r + 2]}; }


// Residues of every stored structure grouped by amino acid (FdDeviceStore::aa_rows / aa_dir): one warp per
// structure, counting sort over the 41 buckets, ascending residue index inside a bucket.
constexpr int AAD_WARPS = 8;
__device__ __forceinline__ uint32_t aa_bucket(uint8_t a, bool cb_ok) {
    if (a == 255 || !cb_ok) return 40u;
    const uint32_t code = a & 0x7fu;
    if (code >= 20u) return 40u;
    return (a & 0x80u) ? 20u + code : code;
}
__global__ void __launch_bounds__(AAD_WARPS * 32)
    k_store_aa_rows(const uint64_t *row_offsets, const uint8_t *aa, const uint8_t *cb_valid, uint64_t n_structs,
                    uint16_t *rows, uint16_t *dir) {
    __shared__ uint32_t s_cnt[AAD_WARPS][FD_AA_DIR];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t s = (uint64_t)blockIdx.x * AAD_WARPS + w;
    if (s >= n_structs) return;
    uint32_t *cnt = s_cnt[w];
    const uint64_t base = row_offsets[s];
    const uint32_t n = (uint32_t)(row_offsets[s + 1] - base);
    for (uint32_t b = lane; b < FD_AA_DIR; b += 32) cnt[b] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32)
        atomicAdd(&cnt[aa_bucket(aa[base + i], cb_valid == nullptr || cb_valid[base + i] != 0)], 1u);
    __syncwarp();
    if (lane == 0) { // exclusive scan: cnt[b] becomes the cursor of bucket b
        uint32_t run = 0;
        for (uint32_t b = 0; b < FD_AA_DIR - 1; b++) {
            const uint32_t c = cnt[b];
            cnt[b] = run;
            dir[s * FD_AA_DIR + b] = (uint16_t)run;
            run += c;
        }
        dir[s * FD_AA_DIR + FD_AA_DIR - 1] = (uint16_t)run;
    }
    __syncwarp();
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint32_t b = i < n ? aa_bucket(aa[base + i], cb_valid == nullptr || cb_valid[base + i] != 0) : 64u + lane;
        const uint32_t m = __match_any_sync(0xffffffffu, b);
        const uint32_t leader = (uint32_t)__ffs((int)m) - 1u;
        uint32_t at = 0;
        if (lane == leader && i < n) {
            at = cnt[b];
            cnt[b] = at + (uint32_t)__popc(m);
        }
        at = __shfl_sync(0xffffffffu, at, (int)leader);
        if (i < n) rows[base + at + (uint32_t)__popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
        __syncwarp();
    }
}

// MODE 0 = count, 1 = emit.  TYPED: any encoding of fd_hashtypes.cuh and `--multiple-bins` (tp); the default
// encoding with one bin pair keeps pair_hash_auto.  Edge key = cand << 36 | i << 20 | j << 4 | bin-pair index, so the
// sort restores the reference's emission order (pair order, then the order of the bin list: retrieve.rs:124-131).
template <int MODE, bool TYPED = false>
__global__ void __launch_bounds__(K4_THREADS)
    k4_candidate_edges(StoreView st, const RQDesc *rq, const uint32_t *q_hashes, const AADist *q_aad,
                       const uint32_t *cand_query, const uint32_t *cand_nid, uint32_t n_cand, fdg::HashParams hp,
                       float ca_cutoff, unsigned long long *n_edges, unsigned long long *n_pairs,
                       uint64_t *edge_keys, uint32_t *edge_hash, uint64_t *pair_keys, uint32_t *pair_q,
                       fdg::TypedParams tp = fdg::TypedParams(), const uint32_t *aad_runs = nullptr) {
    __shared__ uint16_t list1[K4_LIST_CAP], list2[K4_LIST_CAP];
    __shared__ uint32_t n1, n2;
    __shared__ uint32_t q_ij[K4_CHUNK];
    __shared__ float q_d[K4_CHUNK];
    __shared__ uint32_t q_n;
    __shared__ AADist aad[K4_MAX_AADIST];
    const int lane = threadIdx.x & 31;
    const uint32_t c = blockIdx.x;
    if (c >= n_cand) return;
    const RQDesc Q = rq[cand_query[c]];
    const uint32_t t = cand_nid[c];
    const uint64_t base = st.row_offsets[t];
    const uint32_t n = (uint32_t)(st.row_offsets[t + 1] - base);
    if (threadIdx.x == 0) {
        n1 = 0;
        n2 = 0;
        q_n = 0;
    }
    const bool big = Q.run_begin != 0xFFFFFFFFu; // entries read from global memory through the amino-acid-pair runs
    const AADist *gaad = q_aad + Q.aad_begin;
    const uint32_t *runs = big ? aad_runs + Q.run_begin : nullptr;
    if (!big)
        for (uint32_t k = threadIdx.x; k < Q.n_aad; k += K4_THREADS) aad[k] = q_aad[Q.aad_begin + k];
    __syncthreads();
    if (Q.n_hashes == 0 || Q.n_aad == 0) return;

    // ---- 1. prefilter sets (ascending order does not matter here: outputs are sorted afterwards) ----
    bool all_pairs = !Q.use_prefilter;
    if (!all_pairs) {
        // the lists are filled in ASCENDING residue order (block-wide prefix per step, no atomics): the CTAs that share
        // a candidate (gridDim.y) split the pair space by position in list1 x list2, so they must all see the same lists
        __shared__ uint32_t w1[K4_THREADS / 32], w2[K4_THREADS / 32];
        for (uint32_t r0 = 0; r0 < n; r0 += K4_THREADS) {
            const uint32_t r = r0 + threadIdx.x;
            bool in1 = false, in2 = false;
            if (r < n) {
                const uint8_t a = st.aa[base + r];
                // bit 7 of the stored code marks a residue whose name is not the canonical three-letter code
                const bool canonical = (a & 0x80u) == 0;
                const uint32_t code = a & 0x7Fu;
                in1 = canonical && code < 32 && ((Q.aa1_mask >> code) & 1u);
                in2 = canonical && code < 32 && ((Q.aa2_mask >> code) & 1u);
            }
            const uint32_t m1 = __ballot_sync(0xffffffffu, in1), m2 = __ballot_sync(0xffffffffu, in2);
            if (lane == 0) {
                w1[threadIdx.x >> 5] = __popc(m1);
                w2[threadIdx.x >> 5] = __popc(m2);
            }
            __syncthreads();
            uint32_t p1 = n1, p2 = n2;
            for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) {
                p1 += w1[w];
                p2 += w2[w];
            }
            p1 += __popc(m1 & ((1u << lane) - 1));
            p2 += __popc(m2 & ((1u << lane) - 1));
            if (in1 && p1 < K4_LIST_CAP) list1[p1] = (uint16_t)r;
            if (in2 && p2 < K4_LIST_CAP) list2[p2] = (uint16_t)r;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (uint32_t w = 0; w < K4_THREADS / 32; w++) {
                    n1 += w1[w];
                    n2 += w2[w];
                }
            }
            __syncthreads();
        }
        // CombinationVecIterator::is_empty -> fall back to every pair (retrieve.rs:145-151)
        if (n1 == 0 || n2 == 0) all_pairs = true;
    }
    // lists that overflow the shared-memory capacity are handled by masking inside the all-pairs sweep
    const bool masked_sweep = !all_pairs && (n1 > K4_LIST_CAP || n2 > K4_LIST_CAP);
    const uint64_t rows = (all_pairs || masked_sweep) ? n : n1;
    const uint64_t cols = (all_pairs || masked_sweep) ? n : n2;
    const uint64_t total = rows * cols;

    // the chunks of the pair space are dealt round-robin to the CTAs of the candidate (gridDim.y: a handful of large
    // all-pairs candidates would otherwise occupy a handful of SMs; outputs are appended with global atomics and
    // sorted afterwards, so any split is fine)
    for (uint64_t p0 = (uint64_t)blockIdx.y * K4_CHUNK; p0 < total; p0 += (uint64_t)gridDim.y * K4_CHUNK) {
        // ---- 2. cheap screen ----
        for (uint32_t u = 0; u < K4_CHUNK / K4_THREADS; u++) {
            const uint64_t p = p0 + (uint64_t)u * K4_THREADS + threadIdx.x;
            bool pass = false;
            uint32_t i = 0, j = 0;
            float d = 0.f;
            if (p < total) {
                const uint32_t a = (uint32_t)(p / cols), b = (uint32_t)(p % cols);
                i = (all_pairs || masked_sweep) ? a : list1[a];
                j = (all_pairs || masked_sweep) ? b : list2[b];
                bool ok = true;
                const uint8_t ai = st.aa[base + i], aj = st.aa[base + j];
                if (masked_sweep) {
                    ok = !(ai & 0x80u) && !(aj & 0x80u) && ((Q.aa1_mask >> (ai & 31u)) & 1u) &&
                         ((Q.aa2_mask >> (aj & 31u)) & 1u) && (ai & 0x7Fu) < 32 && (aj & 0x7Fu) < 32;
                }
                if (ok) {
                    d = fdg::dist(ld3(st.ca_xyz, base + i), ld3(st.ca_xyz, base + j));
                    if (d <= hp.dist_cutoff) {
                        const uint8_t ci = ai == 255 ? 255 : (ai & 0x7Fu), cj = aj == 255 ? 255 : (aj & 0x7Fu);
                        if (big) {
                            if (ci < 20 && cj < 20)
                                for (uint32_t k = runs[ci * 20u + cj]; k < runs[ci * 20u + cj + 1]; k++)
                                    if (fabsf(d - gaad[k].dist) < ca_cutoff) {
                                        pass = true;
                                        break;
                                    }
                        } else {
                            for (uint32_t k = 0; k < Q.n_aad; k++)
                                if (aad[k].aa1 == ci && aad[k].aa2 == cj && fabsf(d - aad[k].dist) < ca_cutoff) {
                                    pass = true;
                                    break;
                                }
                        }
                    }
                }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                uint32_t pos = 0;
                if (lane == 0) pos = atomicAdd(&q_n, __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1));
                if (pass) {
                    q_ij[pos] = (i << 16) | j;
                    q_d[pos] = d;
                }
            }
        }
        __syncthreads();
        // ---- 3. dense drain ----
        const uint32_t qn = q_n;
        for (uint32_t k0 = 0; k0 < qn; k0 += K4_THREADS) {
            const uint32_t k = k0 + threadIdx.x;
            if (k < qn) {
                const uint32_t i = q_ij[k] >> 16, j = q_ij[k] & 0xffffu;
                const float d = q_d[k];
                const uint64_t ri = base + i, rj = base + j;
                const uint8_t ai = st.aa[ri], aj = st.aa[rj];
                const bool cbok = st.cb_valid == nullptr || (st.cb_valid[ri] && st.cb_valid[rj]);
                // get_single_feature (feature.rs:11-24): i != j, both amino acids known, CB present -- and, for the
                // encodings whose cutoff is not on the CA distance, their own distance within the cutoff
                bool is_feature = i != j && ai != 255 && aj != 255 && (cbok || (TYPED && !fdg::ht_needs_cb(tp.type)));
                const bool need_nb = TYPED && fdg::ht_needs_neighbours(tp.type);
                if (need_nb) is_feature = is_feature && i > 0 && j > 0 && i + 1 < n && j + 1 < n; // feature.rs:113, 163
                float ds = d;
                if (TYPED && is_feature && (tp.type == fdg::HT_TRROSETTA || tp.type == fdg::HT_PPF)) {
                    ds = fdg::typed_screen_dist(tp.type, ld3(st.ca_xyz, ri), ld3(st.cb_xyz, ri), ld3(st.ca_xyz, rj),
                                                ld3(st.cb_xyz, rj));
                    is_feature = !(ds > hp.dist_cutoff);
                }
                if (is_feature) {
                    const uint8_t ci = ai & 0x7Fu, cj = aj & 0x7Fu;
                    uint32_t np = 0;
                    const uint32_t e0 = big ? ((ci < 20 && cj < 20) ? runs[ci * 20u + cj] : 0u) : 0u;
                    const uint32_t e1 = big ? ((ci < 20 && cj < 20) ? runs[ci * 20u + cj + 1] : 0u) : Q.n_aad;
                    for (uint32_t e = e0; e < e1; e++) {
                        const AADist a = big ? gaad[e] : aad[e];
                        if (a.aa1 == ci && a.aa2 == cj && fabsf(d - a.dist) < ca_cutoff) {
                            if (MODE == 1) {
                                const unsigned long long pos = atomicAdd(n_pairs, 1ull);
                                // (the low byte only orders the records of one residue pair; consumers count them)
                                pair_keys[pos] = ((uint64_t)c << 40) | ((uint64_t)i << 24) | ((uint64_t)j << 8) |
                                                 (uint64_t)(a.k & 0xffu);
                                pair_q[pos] = a.q_index;
                            }
                            np++;
                        }
                    }
                    if (MODE == 0 && np) atomicAdd(n_pairs, (unsigned long long)np);
                    float f[9];
                    if (TYPED) {
                        fdg::Nbr nb;
                        if (need_nb) {
                            nb.ca_pre1 = ld3(st.ca_xyz, ri - 1);
                            nb.ca_next1 = ld3(st.ca_xyz, ri + 1);
                            nb.ca_pre2 = ld3(st.ca_xyz, rj - 1);
                            nb.ca_next2 = ld3(st.ca_xyz, rj + 1);
                            nb.seq_dist = (float)j - (float)i;
                        }
                        fdg::typed_feature(tp.type, ld3(st.n_xyz, ri), ld3(st.ca_xyz, ri), ld3(st.cb_xyz, ri),
                                           ld3(st.n_xyz, rj), ld3(st.ca_xyz, rj), ld3(st.cb_xyz, rj), (float)ci, (float)cj,
                                           ds, f, &nb);
                    }
                    const uint32_t nb = TYPED ? tp.n_bins : 1u;
                    for (uint32_t bi = 0; bi < nb; bi++) {
                        const uint32_t h = TYPED ? fdg::typed_hash(tp.type, f, tp.nbd[bi], tp.nba[bi])
                                                 : fdg::pair_hash_auto(ld3(st.n_xyz, ri), ld3(st.ca_xyz, ri),
                                                                       ld3(st.cb_xyz, ri), ld3(st.n_xyz, rj),
                                                                       ld3(st.ca_xyz, rj), ld3(st.cb_xyz, rj), ci, cj, d, hp);
                        // membership in the query hash set
                        uint32_t lo = 0, hi = Q.n_hashes;
                        const uint32_t *hs = q_hashes + Q.hash_begin;
                        while (lo < hi) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (hs[mid] < h) lo = mid + 1;
                            else hi = mid;
                        }
                        if (lo < Q.n_hashes && hs[lo] == h) {
                            const unsigned long long pos = atomicAdd(n_edges, 1ull);
                            if (MODE == 1) {
                                edge_keys[pos] = ((uint64_t)c << 36) | ((uint64_t)i << 20) | ((uint64_t)j << 4) | bi;
                                edge_hash[pos] = h;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) q_n = 0;
        __syncthreads();
    }
}

__global__ void k4_unpack_edges(const uint64_t *keys, const uint32_t *hash, uint64_t n, fd_cand_edge *out) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t key = keys[k];
    out[k] = fd_cand_edge{(uint32_t)(key >> 36), (uint32_t)((key >> 20) & 0xffffu), (uint32_t)((key >> 4) & 0xffffu), hash[k]};
}
__global__ void k4_unpack_pairs(const uint64_t *keys, const uint32_t *qidx, uint64_t n, fd_cand_pair *out) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t key = keys[k];
    out[k] = fd_cand_pair{(uint32_t)(key >> 40), qidx[k], (uint32_t)((key >> 24) & 0xffffu),
                          (uint32_t)((key >> 8) & 0xffffu), (uint32_t)(key & 0xffu)};
}

int sort_pairs_u64_u32(fd_ctx *ctx, DevBuf<uint64_t> &keys, DevBuf<uint32_t> &vals, uint64_t n, int end_bit) {
    if (n == 0) return FD_OK;
    DevBuf<uint64_t> k2;
    DevBuf<uint32_t> v2;
    DevBuf<uint8_t> tmp;
    FD_CUDA(ctx, k2.alloc(n));
    FD_CUDA(ctx, v2.alloc(n));
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.p, k2.p, vals.p, v2.p, n, 0, end_bit, ctx->stream);
    FD_CUDA(ctx, tmp.alloc(tb));
    FD_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, k2.p, vals.p, v2.p, n, 0, end_bit, ctx->stream));
    ctx->launches += 8;
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::swap(keys.p, k2.p);
    std::swap(vals.p, v2.p);
    return FD_OK;
}

} // namespace

extern "C" {

int fd_store_attach(fd_ctx *ctx, const fd_struct_batch *batch) {
    if (!ctx) return FD_ERR_ARG;
    if (ctx->borrowed) return fd_fail(ctx, FD_ERR_STATE, "fd_store_attach: a forked context shares its parent's store");
    FD_ENTER(ctx);
    FdDeviceBatch d;
    FD_TRY(fd_upload_batch(ctx, batch, &d));
    for (uint64_t s = 0; s < batch->n_structs; s++)
        if (batch->row_offsets[s + 1] - batch->row_offsets[s] > 65535)
            return fd_fail(ctx, FD_ERR_LIMIT, "structure with more than 65535 residues (reference max_residue)");
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    fd_ctx_release_store(ctx);
    FdDeviceStore &st = ctx->store;
    st.n_structs = d.n_structs;
    st.n_res = d.n_res;
    st.row_offsets = d.row_offsets.take();
    st.n_xyz = d.n_xyz.take();
    st.ca_xyz = d.ca_xyz.take();
    st.cb_xyz = d.cb_xyz.take();
    st.aa = d.aa.take();
    if (batch->cb_valid) st.cb_valid = d.cb_valid.take();
    {
        DevBuf<uint16_t> rows, dir;
        FD_CUDA(ctx, rows.alloc(std::max<uint64_t>(st.n_res, 1)));
        FD_CUDA(ctx, dir.alloc(std::max<uint64_t>(st.n_structs, 1) * FD_AA_DIR));
        if (st.n_structs)
            FD_LAUNCH(ctx, k_store_aa_rows, fd_div_up(st.n_structs, (uint64_t)AAD_WARPS), AAD_WARPS * 32, 0, st.row_offsets,
                      st.aa, st.cb_valid, st.n_structs, rows.p, dir.p);
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        st.aa_rows = rows.take();
        st.aa_dir = dir.take();
    }
    st.h_row_offsets.assign(batch->row_offsets, batch->row_offsets + batch->n_structs + 1);
    st.attached = true;
    return FD_OK;
}

int fd_candidate_edges_batch(fd_ctx *ctx, const fd_retrieval_query *queries, uint32_t nq, const uint32_t *cand_query,
                             const uint32_t *cand_nid, uint64_t n_cand, const fd_hash_params *params,
                             float ca_dist_cutoff, fd_cand_edge **out_edges, uint64_t *out_n_edges,
                             fd_cand_pair **out_pairs, uint64_t *out_n_pairs) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->store.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_candidate_edges_batch: no structure store attached");
    if ((nq && !queries) || (n_cand && (!cand_query || !cand_nid)) || !params || !out_edges || !out_n_edges ||
        !out_pairs || !out_n_pairs)
        return fd_fail(ctx, FD_ERR_ARG, "fd_candidate_edges_batch: NULL argument");
    if (n_cand >= (1ull << 24))
        return fd_fail(ctx, FD_ERR_LIMIT, "at most 2^24 - 1 candidates per call; split the batch");
    fdg::TypedParams tp;
    if (const char *why = fdg::typed_params_from(params, &tp)) return fd_fail(ctx, FD_ERR_ARG, why);
    const bool typed = !fdg::ht_default_route(params);
    FD_ENTER(ctx);
    *out_edges = nullptr;
    *out_pairs = nullptr;
    *out_n_edges = 0;
    *out_n_pairs = 0;
    std::vector<RQDesc> descs(nq);
    std::vector<uint32_t> f_hash, f_runs;
    std::vector<AADist> f_aad;
    for (uint32_t q = 0; q < nq; q++) {
        const fd_retrieval_query &Q = queries[q];
        // no amino-acid prefilter for the encodings without amino-acid fields (retrieve.rs:385-390)
        RQDesc d{(uint32_t)f_hash.size(), Q.n_hashes, (uint32_t)f_aad.size(), Q.n_aa_dist, 0, 0,
                 Q.n_hashes <= PREFILTER_AA_SKIPPING_SIZE && fdg::ht_has_aa_index(tp.type) ? 1u : 0u, 0xFFFFFFFFu};
        for (uint32_t k = 0; k < Q.n_hashes; k++) {
            if (k && Q.hashes_sorted[k] <= Q.hashes_sorted[k - 1])
                return fd_fail(ctx, FD_ERR_ARG, "fd_retrieval_query: hashes_sorted must be strictly ascending");
            f_hash.push_back(Q.hashes_sorted[k]);
            uint32_t a1, a2; // prefilter_amino_acid reads them from reverse_hash_default (retrieve.rs:575-577)
            fdg::typed_hash_aa(tp.type, Q.hashes_sorted[k], &a1, &a2);
            d.aa1_mask |= 1u << (a1 & 31u);
            d.aa2_mask |= 1u << (a2 & 31u);
        }
        if (Q.n_aa_dist <= (uint32_t)K4_MAX_AADIST) {
            for (uint32_t k = 0; k < Q.n_aa_dist; k++) {
                uint16_t pos = 0;
                for (uint32_t m = 0; m < k; m++)
                    if (Q.aa1[m] == Q.aa1[k] && Q.aa2[m] == Q.aa2[k]) pos++;
                f_aad.push_back(AADist{Q.aa1[k], Q.aa2[k], pos, Q.ca_dist[k], Q.q_index[k]});
            }
        } else { // stable counting sort by amino-acid pair; k = position inside the pair's list, as above
            std::vector<uint32_t> start(401, 0);
            for (uint32_t k = 0; k < Q.n_aa_dist; k++) {
                if (Q.aa1[k] >= 20 || Q.aa2[k] >= 20)
                    return fd_fail(ctx, FD_ERR_ARG, "fd_retrieval_query: amino-acid code out of range");
                start[Q.aa1[k] * 20u + Q.aa2[k] + 1]++;
            }
            for (int b = 0; b < 400; b++) start[b + 1] += start[b];
            d.run_begin = (uint32_t)f_runs.size();
            f_runs.insert(f_runs.end(), start.begin(), start.end());
            const size_t base = f_aad.size();
            f_aad.resize(base + Q.n_aa_dist);
            std::vector<uint32_t> cur(start.begin(), start.end() - 1);
            for (uint32_t k = 0; k < Q.n_aa_dist; k++) {
                const uint32_t b = Q.aa1[k] * 20u + Q.aa2[k];
                f_aad[base + cur[b]] = AADist{Q.aa1[k], Q.aa2[k], (uint16_t)((cur[b] - start[b]) & 0xffffu), Q.ca_dist[k], Q.q_index[k]};
                cur[b]++;
            }
        }
        descs[q] = d;
    }
    for (uint64_t c = 0; c < n_cand; c++) {
        if (cand_query[c] >= nq) return fd_fail(ctx, FD_ERR_ARG, "cand_query out of range");
        if (cand_nid[c] >= ctx->store.n_structs) return fd_fail(ctx, FD_ERR_ARG, "cand_nid outside the attached store");
    }
    if (n_cand == 0) {
        *out_edges = (fd_cand_edge *)malloc(sizeof(fd_cand_edge));
        *out_pairs = (fd_cand_pair *)malloc(sizeof(fd_cand_pair));
        return FD_OK;
    }
    cudaStream_t s = ctx->stream;
    DevBuf<RQDesc> d_desc;
    DevBuf<uint32_t> d_hash, d_cq, d_cn, d_edge_hash, d_pair_q, d_runs;
    DevBuf<AADist> d_aad;
    DevBuf<unsigned long long> d_cnt;
    DevBuf<uint64_t> d_edge_keys, d_pair_keys;
    FD_CUDA(ctx, d_desc.alloc(nq));
    FD_CUDA(ctx, d_hash.alloc(f_hash.size()));
    FD_CUDA(ctx, d_aad.alloc(f_aad.size()));
    FD_CUDA(ctx, d_runs.alloc(f_runs.size()));
    if (!f_runs.empty())
        FD_CUDA(ctx, cudaMemcpyAsync(d_runs.p, f_runs.data(), f_runs.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    FD_CUDA(ctx, d_cq.alloc(n_cand));
    FD_CUDA(ctx, d_cn.alloc(n_cand));
    FD_CUDA(ctx, d_cnt.alloc(2));
    FD_CUDA(ctx, cudaMemcpyAsync(d_desc.p, descs.data(), nq * sizeof(RQDesc), cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_hash.p, f_hash.data(), f_hash.size() * 4, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_aad.p, f_aad.data(), f_aad.size() * sizeof(AADist), cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_cq.p, cand_query, n_cand * 4, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_cn.p, cand_nid, n_cand * 4, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemsetAsync(d_cnt.p, 0, 16, s));
    const FdDeviceStore &S = ctx->store;
    StoreView sv{S.row_offsets, S.n_xyz, S.ca_xyz, S.cb_xyz, S.aa, S.cb_valid};
    fdg::HashParams hp = fdg::make_params(params->nbin_dist, params->nbin_angle, params->dist_cutoff);
    StageTimer st(ctx, "edges");
    // few candidates (the rare ones beyond the fused kernels' limits): several CTAs per candidate fill the GPU
    uint32_t split = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(64, ((uint64_t)ctx->num_sms * 8) / std::max<uint64_t>(n_cand, 1)));
    if (const char *e = getenv("FD_K4_SPLIT")) split = (uint32_t)std::max(1, std::min(64, atoi(e)));
    const dim3 k4_grid((uint32_t)n_cand, split);
    if (typed)
        FD_LAUNCH(ctx, (k4_candidate_edges<0, true>), k4_grid, K4_THREADS, 0, sv, d_desc.p, d_hash.p, d_aad.p, d_cq.p,
                  d_cn.p, (uint32_t)n_cand, hp, ca_dist_cutoff, d_cnt.p, d_cnt.p + 1, (uint64_t *)nullptr,
                  (uint32_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr, tp, d_runs.p);
    else
        FD_LAUNCH(ctx, k4_candidate_edges<0>, k4_grid, K4_THREADS, 0, sv, d_desc.p, d_hash.p, d_aad.p, d_cq.p,
                  d_cn.p, (uint32_t)n_cand, hp, ca_dist_cutoff, d_cnt.p, d_cnt.p + 1, (uint64_t *)nullptr,
                  (uint32_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr, tp, d_runs.p);
    unsigned long long cnt[2] = {0, 0};
    FD_CUDA(ctx, cudaMemcpyAsync(cnt, d_cnt.p, 16, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, cudaStreamSynchronize(s));
    const uint64_t ne = cnt[0], np = cnt[1];
    FD_CUDA(ctx, d_edge_keys.alloc(ne));
    FD_CUDA(ctx, d_edge_hash.alloc(ne));
    FD_CUDA(ctx, d_pair_keys.alloc(np));
    FD_CUDA(ctx, d_pair_q.alloc(np));
    FD_CUDA(ctx, cudaMemsetAsync(d_cnt.p, 0, 16, s));
    if ((ne || np) && typed)
        FD_LAUNCH(ctx, (k4_candidate_edges<1, true>), k4_grid, K4_THREADS, 0, sv, d_desc.p, d_hash.p, d_aad.p,
                  d_cq.p, d_cn.p, (uint32_t)n_cand, hp, ca_dist_cutoff, d_cnt.p, d_cnt.p + 1, d_edge_keys.p,
                  d_edge_hash.p, d_pair_keys.p, d_pair_q.p, tp, d_runs.p);
    else if (ne || np)
        FD_LAUNCH(ctx, k4_candidate_edges<1>, k4_grid, K4_THREADS, 0, sv, d_desc.p, d_hash.p, d_aad.p,
                  d_cq.p, d_cn.p, (uint32_t)n_cand, hp, ca_dist_cutoff, d_cnt.p, d_cnt.p + 1, d_edge_keys.p,
                  d_edge_hash.p, d_pair_keys.p, d_pair_q.p, tp, d_runs.p);
    FD_TRY(sort_pairs_u64_u32(ctx, d_edge_keys, d_edge_hash, ne, 60));
    FD_TRY(sort_pairs_u64_u32(ctx, d_pair_keys, d_pair_q, np, 64));
    DevBuf<fd_cand_edge> d_oe;
    DevBuf<fd_cand_pair> d_op;
    FD_CUDA(ctx, d_oe.alloc(ne));
    FD_CUDA(ctx, d_op.alloc(np));
    if (ne) FD_LAUNCH(ctx, k4_unpack_edges, fd_div_up(ne, 256), 256, 0, d_edge_keys.p, d_edge_hash.p, ne, d_oe.p);
    if (np) FD_LAUNCH(ctx, k4_unpack_pairs, fd_div_up(np, 256), 256, 0, d_pair_keys.p, d_pair_q.p, np, d_op.p);
    fd_cand_edge *he = (fd_cand_edge *)malloc(std::max<uint64_t>(ne, 1) * sizeof(fd_cand_edge));
    fd_cand_pair *hp2 = (fd_cand_pair *)malloc(std::max<uint64_t>(np, 1) * sizeof(fd_cand_pair));
    if (!he || !hp2) {
        free(he);
        free(hp2);
        return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
    }
    FD_CUDA(ctx, cudaMemcpyAsync(he, d_oe.p, ne * sizeof(fd_cand_edge), cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, cudaMemcpyAsync(hp2, d_op.p, np * sizeof(fd_cand_pair), cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, st.finish());
    *out_edges = he;
    *out_n_edges = ne;
    *out_pairs = hp2;
    *out_n_pairs = np;
    return FD_OK;
}

} // extern "C"
