// fd_verify.cu -- K6: candidate verification (query path (ii), everything after the prefilter).
//
// Replaces retrieval_wrapper (reference src/controller/retrieve.rs:364-552) for a batch of
// (query, candidate structure) pairs, so that nothing but the final match records leaves HBM:
//   retrieve_with_prefilter   (:52-156)    re-hash of the candidate's residue pairs from the HBM structure store
//   create_index_graph + connected_components_with_given_node_count (graph.rs:16-50)
//   map_query_and_retrieved_residues (:604-702), calculate_subgraph_idf (:705-719), the rescue loop (:453-516)
//   rmsd_with_calpha_and_rottran -> kabsch (:756-834, kabsch.rs:157-554)
// The un-fused pieces (fd_candidate_edges_batch + host graph step + fd_kabsch_store_batch) remain as the general
// path; these kernels handle candidates that fit their shared-memory limits and flag the rest.
//
// Three kernels, one per kind of work (a single fused kernel was 145 KB of SASS and spent 40 % of its issue time
// waiting for instruction fetch, with 6 candidates per SM in flight -- profiles/r01a):
//   k6a_edges      CTA per candidate: amino-acid prefilter lists, pair screen through a 20x20 amino-acid-pair
//                  table, f32 feature + binary64 trig hash of the survivors (queued across chunks so the expensive
//                  hash runs with full lanes), membership in the query's hash set, bitonic sort by (i, j);
//                  the candidate's edges go to a compact pool in HBM.
//   k6b_components warp per candidate (no CTA barriers): graph components via 64-bit reachability masks
//                  (<= 64 nodes), residue mapping, rescue vote; one 128-byte spec per component.
//   k6c_kabsch     thread per component: f64 Kabsch, match record written at its final position
//                  (candidate order, component order) from an exclusive scan of the per-candidate counts.
#include <cub/cub.cuh>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <thread>

#include "fd_common.cuh"
#include "fd_geom.cuh"
#include "fd_kabsch.cuh"

namespace {

constexpr int VA_THREADS = 128;            // k6a CTA
constexpr int VA_COL_STAGE = 512;          // candidates with at most this many column residues stage them in shared memory
constexpr int VA_HASH_STAGE = 256;         // query hash sets up to this size are staged in shared memory
constexpr int V_LIST_CAP = 2048;           // prefilter-list capacity (larger candidates take the general path)
constexpr int VB_WARPS = 8;                // k6b: candidates per CTA (their components are shared by the CTA's warps)
constexpr int V_MAX_AAD = 255; // start and count of an amino-acid pair's run each fit 8 bits
constexpr int V_MAX_E = 256;
constexpr int V_MAX_NODES = 64;
constexpr int V_MAX_C = 64; // <= 64 nodes give at most 32 SCCs + 32 weak components of size >= 2: the limit never binds
constexpr int V_MAX_NQ = 16;
constexpr uint32_t V_PREFILTER_SKIP = 200; // retrieve.rs:24

struct StoreView {
    const uint64_t *row_offsets;
    const float *n_xyz, *ca_xyz, *cb_xyz;
    const uint8_t *aa, *cb_valid;
    const uint16_t *aa_rows, *aa_dir; // residues grouped by amino acid (FdDeviceStore)
};

struct VQDesc {                 // per query
    uint32_t hash_begin, n_hashes; // sorted hash set + per-hash info
    uint32_t aad_begin, n_aad;
    uint32_t aa1_mask, aa2_mask;
    uint32_t use_prefilter;
    uint32_t idx_begin, n_idx;   // Q.indices as dense query-residue ids (may repeat)
    uint32_t n_dq;               // number of dense query-residue ids (ascending residue index)
    uint32_t qres_base;          // first dense residue of this query in q_ca / q_cb
};
struct VHash {                  // per sorted query hash
    uint32_t hash;
    float idf;
    uint8_t dqi, dqj, sym, pad;
};
struct VAad {
    uint8_t aa1, aa2, dq, pad; // dq: dense id of the entry's query residue
    float dist;
};

__device__ __forceinline__ fdg::V3 ld3(const float *p, uint64_t r) { return {p[3 * r], p[3 * r + 1], p[3 * r + 2]}; }
__device__ __forceinline__ int ctz64(uint64_t v) { return __ffsll((long long)v) - 1; }

// lexicographic order of the sorted node lists of two different component masks (graph.rs:45-46)
__device__ __forceinline__ bool mask_less(uint64_t a, uint64_t b) {
    const uint64_t x = a ^ b;
    const int d = ctz64(x);
    const uint64_t above = ~((2ull << d) - 1ull);
    if ((a >> d) & 1ull) return (b & above) != 0; // a has d; smaller unless b is a proper prefix of a
    return (a & above) == 0;                      // b has d; a smaller only if a is a prefix of b
}

struct CompSpec { // one connected component of one candidate, k6b -> k6c (128 bytes)
    uint32_t cand;
    uint16_t ci, nal;        // component number inside the candidate; residue pairs to superpose
    float idf;               // calculate_subgraph_idf
    uint32_t pad;
    uint32_t res[V_MAX_NQ];  // reported target residue index + 1 per query position, 0 = none
    uint8_t aq[V_MAX_NQ];    // alignment: dense query residue ids ...
    uint16_t at[V_MAX_NQ];   // ... and target residue indices
};
static_assert(sizeof(CompSpec) == 128, "CompSpec layout");

template <class I>
struct GatherPointsT { // CA, CB interleaved, gathered by residue index: point 2k = CA(res[k]), 2k+1 = CB(res[k])
    const float *ca, *cb;
    const I *res;
    uint64_t base;
    __device__ fdk::P3 operator()(uint32_t i) const {
        const uint64_t r = base + res[i >> 1];
        const float *s = (i & 1) ? cb : ca;
        return {(double)s[3 * r], (double)s[3 * r + 1], (double)s[3 * r + 2]};
    }
};

// amino-acid-pair table of a query in shared memory: aad[] sorted by (aa1, aa2), aa_range[aa1 * 20 + aa2] =
// start << 8 | count of the pair's run (0 = pair not in the query).  NT cooperating threads, caller synchronises.
template <int NT>
__device__ __forceinline__ void load_aad_phase1(const VQDesc &Q, const VAad *vaad, VAad *aad, uint16_t *aa_range, int t) {
    for (uint32_t k = t; k < Q.n_aad; k += NT) aad[k] = vaad[Q.aad_begin + k];
    for (uint32_t k = t; k < 400; k += NT) aa_range[k] = 0;
}
template <int NT>
__device__ __forceinline__ void load_aad_phase2(const VQDesc &Q, const VAad *aad, uint16_t *aa_range, int t) {
    for (uint32_t k = t; k < Q.n_aad; k += NT) { // heads of runs of equal (aa1, aa2)
        const uint32_t key = aad[k].aa1 * 20u + aad[k].aa2;
        if (k == 0 || aad[k - 1].aa1 * 20u + aad[k - 1].aa2 != key) {
            uint32_t e = k + 1;
            while (e < Q.n_aad && aad[e].aa1 * 20u + aad[e].aa2 == key) e++;
            aa_range[key] = (uint16_t)((k << 8) | (e - k));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k6a: candidate edges
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VA_THREADS, 7)
    k6a_edges(StoreView st, const VQDesc *vq, const VHash *vhash, const VAad *vaad, const uint32_t *cand_query,
              const uint32_t *cand_nid, uint32_t n_cand, fdg::HashParams hp, float ca_cutoff, uint32_t *pool_key,
              uint16_t *pool_ent, unsigned int *pool_count, uint32_t pool_cap, uint32_t *cand_ebegin,
              uint32_t *cand_ne, uint8_t *cand_flags) {
    __shared__ uint16_t list1[V_LIST_CAP], list2[V_LIST_CAP];
    __shared__ uint32_t n1, n2, n_e, s_base;
    __shared__ int s_dmax_bits, s_dmin_bits; // range of the query's CA distances (positive floats order like ints)
    __shared__ uint32_t wq_ij[VA_THREADS / 32][64], wb_ij[VA_THREADS / 32][64]; // warp-private queues A and B
    __shared__ float wq_d[VA_THREADS / 32][64], wb_d[VA_THREADS / 32][64];
    __shared__ uint16_t wb_lo[VA_THREADS / 32][64];
    __shared__ uint32_t s_hash[VA_HASH_STAGE];
    __shared__ float4 s_col[VA_COL_STAGE]; // column residues: CA xyz, w = residue index << 8 | amino acid
    __shared__ VAad aad[V_MAX_AAD];
    __shared__ uint16_t aa_range[400];
    __shared__ uint32_t e_key[V_MAX_E]; // i << 16 | j
    __shared__ uint16_t e_ent[V_MAX_E]; // index of the edge's hash inside the query's sorted hash set

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t c = blockIdx.x;
    if (c >= n_cand) return;
    const VQDesc Q = vq[cand_query[c]];
    const uint32_t t = cand_nid[c];
    const uint64_t base = st.row_offsets[t];
    const uint32_t n = (uint32_t)(st.row_offsets[t + 1] - base);
    const VHash *H = vhash + Q.hash_begin;
    if (tid == 0) {
        n1 = n2 = n_e = 0;
        s_dmax_bits = 0;
        s_dmin_bits = 0x7f7fffff;
    }
    load_aad_phase1<VA_THREADS>(Q, vaad, aad, aa_range, tid);
    __syncthreads();
    if (Q.n_hashes == 0 || Q.n_aad == 0) return;
    load_aad_phase2<VA_THREADS>(Q, aad, aa_range, tid);
    {
        int bmax = 0, bmin = 0x7f7fffff;
        for (uint32_t k = tid; k < Q.n_aad; k += VA_THREADS) {
            const int bits = __float_as_int(fmaxf(aad[k].dist, 0.f));
            bmax = max(bmax, bits);
            bmin = min(bmin, bits);
        }
        bmax = __reduce_max_sync(0xffffffffu, bmax);
        bmin = __reduce_min_sync(0xffffffffu, bmin);
        if (lane == 0) {
            atomicMax(&s_dmax_bits, bmax);
            atomicMin(&s_dmin_bits, bmin);
        }
    }
    if (Q.n_hashes <= VA_HASH_STAGE)
        for (uint32_t k = tid; k < Q.n_hashes; k += VA_THREADS) s_hash[k] = H[k].hash;
    __syncthreads();
    const float s_dmax = __int_as_float(s_dmax_bits), s_dmin = __int_as_float(s_dmin_bits);

    // ---- prefilter sets (prefilter_amino_acid, retrieve.rs:563-602) ----
    bool all_pairs = !Q.use_prefilter;
    if (!all_pairs) {
        for (uint32_t r0 = 0; r0 < n; r0 += VA_THREADS) {
            const uint32_t r = r0 + tid;
            bool in1 = false, in2 = false;
            if (r < n) {
                const uint8_t a = st.aa[base + r];
                const bool canonical = (a & 0x80u) == 0;
                in1 = canonical && ((Q.aa1_mask >> (a & 31u)) & 1u);
                in2 = canonical && ((Q.aa2_mask >> (a & 31u)) & 1u);
            }
            const uint32_t m1 = __ballot_sync(0xffffffffu, in1), m2 = __ballot_sync(0xffffffffu, in2);
            uint32_t p1 = 0, p2 = 0;
            if (lane == 0) {
                if (m1) p1 = atomicAdd(&n1, __popc(m1));
                if (m2) p2 = atomicAdd(&n2, __popc(m2));
            }
            p1 = __shfl_sync(0xffffffffu, p1, 0) + __popc(m1 & ((1u << lane) - 1));
            p2 = __shfl_sync(0xffffffffu, p2, 0) + __popc(m2 & ((1u << lane) - 1));
            if (in1 && p1 < V_LIST_CAP) list1[p1] = (uint16_t)r;
            if (in2 && p2 < V_LIST_CAP) list2[p2] = (uint16_t)r;
        }
        __syncthreads();
        if (n1 == 0 || n2 == 0) all_pairs = true; // CombinationVecIterator::is_empty (retrieve.rs:145-151)
        else if (n1 > V_LIST_CAP || n2 > V_LIST_CAP) {
            if (tid == 0) cand_flags[c] = 16; // prefilter list beyond the shared-memory capacity: // handled by the general path
            return;
        }
    }
    const uint32_t rows = all_pairs ? n : n1, cols = all_pairs ? n : n2;

    // ---- retrieve_with_prefilter: screen, hash, keep pairs whose hash is in the query set ----
    // The four warps of the CTA work on the candidate independently (no CTA barrier inside the screen).  Inside a warp,
    // lane -> (row of the current row block, column phase): the row's residue stays in registers and the lanes of a row
    // block walk the same column (one broadcast load).  A pair reaches sqrt + the exact |d - d_q| test only if its
    // squared CA distance lies inside the query's distance range.  Survivors go through two warp-private queues:
    //   A -> stage 1 (no trigonometry): bits 12..31 of the hash (amino acids, CA / CB distance bins) must occur in the
    //        query's hash set;
    //   B -> stage 2: the full hash (binary64 trigonometry) on 32 full lanes, then membership.
    const uint32_t warp = (uint32_t)tid >> 5;
    const uint32_t RP = max(1u, min(rows, 32u));
    const uint32_t CSW = 32u / RP;                     // column phases inside a warp
    const uint32_t PH = CSW * (VA_THREADS / 32);       // column phases of the CTA
    const bool t_active = (uint32_t)lane < RP * CSW;
    const uint32_t t_row = (uint32_t)lane % RP, t_ph = warp * CSW + (uint32_t)lane / RP;
    const float pre_hi = fminf(hp.dist_cutoff, s_dmax + ca_cutoff), pre_lo = fmaxf(s_dmin - ca_cutoff, 0.f);
    const float pre_hi2 = pre_hi * pre_hi * 1.0001f, pre_lo2 = pre_lo * pre_lo * 0.9999f;
    const bool staged = Q.n_hashes <= VA_HASH_STAGE; // the hash set is in shared memory (s_hash), else global
    auto hash_at = [&](uint32_t k) -> uint32_t { return staged ? s_hash[k] : H[k].hash; };
    const bool cstaged = cols <= VA_COL_STAGE;
    if (cstaged) {
        for (uint32_t b = tid; b < cols; b += VA_THREADS) {
            const uint32_t j = all_pairs ? b : list2[b];
            const fdg::V3 cj = ld3(st.ca_xyz, base + j);
            s_col[b] = make_float4(cj.x, cj.y, cj.z, __uint_as_float((j << 8) | st.aa[base + j]));
        }
        __syncthreads();
    }
    uint32_t *qa_ij = wq_ij[warp], *qb_ij = wb_ij[warp];
    float *qa_d = wq_d[warp], *qb_d = wb_d[warp];
    uint16_t *qb_lo = wb_lo[warp];
    uint32_t na = 0, nb = 0; // queue fill, warp-uniform
    const uint32_t lt = (1u << lane) - 1u;
    auto stage2 = [&](uint32_t first, uint32_t cnt) { // entries [first, first + cnt) of B, cnt <= 32
        if ((uint32_t)lane < cnt) {
            const uint32_t ij = qb_ij[first + lane];
            const uint64_t ri = base + (ij >> 16), rj = base + (ij & 0xffffu);
            const uint32_t h = fdg::pair_hash_auto(ld3(st.n_xyz, ri), ld3(st.ca_xyz, ri), ld3(st.cb_xyz, ri),
                                              ld3(st.n_xyz, rj), ld3(st.ca_xyz, rj), ld3(st.cb_xyz, rj),
                                              st.aa[ri] & 0x7Fu, st.aa[rj] & 0x7Fu, qb_d[first + lane], hp);
            // the hashes that share the upper bits start at the stage-1 lower bound
            for (uint32_t k = qb_lo[first + lane]; k < Q.n_hashes; k++) {
                const uint32_t hk = hash_at(k);
                if (hk >= h) {
                    if (hk == h) {
                        const uint32_t pos = atomicAdd(&n_e, 1u);
                        if (pos < V_MAX_E) {
                            e_key[pos] = ij;
                            e_ent[pos] = (uint16_t)k;
                        }
                    }
                    break;
                }
            }
        }
        __syncwarp();
    };
    auto stage1 = [&](uint32_t first, uint32_t cnt) { // entries [first, first + cnt) of A, cnt <= 32
        bool keep = false;
        uint32_t ij = 0, lo = 0;
        float d = 0.f;
        if ((uint32_t)lane < cnt) {
            ij = qa_ij[first + lane];
            d = qa_d[first + lane];
            const uint64_t ri = base + (ij >> 16), rj = base + (ij & 0xffffu);
            if (st.cb_valid == nullptr || (st.cb_valid[ri] && st.cb_valid[rj])) {
                const uint32_t up = fdg::hash_upper(ld3(st.cb_xyz, ri), ld3(st.cb_xyz, rj), st.aa[ri] & 0x7Fu,
                                                    st.aa[rj] & 0x7Fu, d, hp);
                uint32_t hi = Q.n_hashes;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if ((hash_at(mid) >> 12) < up) lo = mid + 1;
                    else hi = mid;
                }
                keep = lo < Q.n_hashes && (hash_at(lo) >> 12) == up;
            }
        }
        const uint32_t km = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const uint32_t pos = nb + __popc(km & lt);
            qb_ij[pos] = ij;
            qb_d[pos] = d;
            qb_lo[pos] = (uint16_t)lo;
        }
        nb += __popc(km);
        __syncwarp();
        if (nb >= 32) {
            stage2(nb - 32, 32);
            nb -= 32;
        }
    };
    for (uint32_t row0 = 0; row0 < rows; row0 += RP) {
        const uint32_t a = row0 + t_row;
        bool row_ok = t_active && a < rows;
        uint32_t i = 0, aim = 0;
        fdg::V3 cai{0.f, 0.f, 0.f};
        if (row_ok) {
            i = all_pairs ? a : list1[a];
            const uint8_t ai = st.aa[base + i];
            row_ok = ai != 255;
            aim = (ai & 0x7Fu) * 20u;
            cai = ld3(st.ca_xyz, base + i);
        }
        for (uint32_t b0 = 0; b0 < cols; b0 += PH) {
            const uint32_t b = b0 + t_ph;
            bool surv = false;
            uint32_t j = 0;
            float d = 0.f;
            if (row_ok && b < cols) {
                fdg::V3 caj;
                uint32_t aj;
                if (cstaged) {
                    const float4 cj = s_col[b];
                    const uint32_t w = __float_as_uint(cj.w);
                    j = w >> 8;
                    aj = w & 0xffu;
                    caj = fdg::V3{cj.x, cj.y, cj.z};
                } else {
                    j = all_pairs ? b : list2[b];
                    aj = st.aa[base + j];
                    caj = ld3(st.ca_xyz, base + j);
                }
                const uint32_t rg = (j == i || aj == 255) ? 0u : aa_range[aim + (aj & 0x7Fu)];
                if (rg != 0) {
                    const float d2 = fdg::dist2(cai, caj);
                    if (d2 <= pre_hi2 && d2 >= pre_lo2) {
                        d = FD_SQRT(d2); // == fdg::dist
                        if (d <= hp.dist_cutoff)
                            for (uint32_t k = rg >> 8, ke = (rg >> 8) + (rg & 0xffu); k < ke; k++)
                                if (fabsf(d - aad[k].dist) < ca_cutoff) {
                                    surv = true;
                                    break;
                                }
                    }
                }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, surv);
            if (m) {
                if (surv) {
                    const uint32_t pos = na + __popc(m & lt);
                    qa_ij[pos] = (i << 16) | j;
                    qa_d[pos] = d;
                }
                na += __popc(m);
                __syncwarp();
                if (na >= 32) {
                    stage1(na - 32, 32);
                    na -= 32;
                }
            }
        }
    }
    if (na) stage1(0, na);
    if (nb > 32) {
        stage2(32, nb - 32);
        nb = 32;
    }
    if (nb) stage2(0, nb);
    __syncthreads();
    const uint32_t ne = n_e;
    if (ne == 0) return;
    if (ne > V_MAX_E || Q.n_idx > V_MAX_NQ || Q.n_dq > V_MAX_NQ) {
        if (tid == 0) cand_flags[c] = 2; // more matching edges (or query residues) than the fused kernels hold
        return;
    }
    // ---- sort edges by (i, j): the reference's emission order (graph node numbering, f32 sum order) ----
    uint32_t sort_n = 2;
    while (sort_n < ne) sort_n <<= 1;
    for (uint32_t k = ne + tid; k < sort_n; k += VA_THREADS) e_key[k] = 0xffffffffu;
    __syncthreads();
    for (uint32_t size = 2; size <= sort_n; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t k = tid; k < sort_n / 2; k += VA_THREADS) {
                const uint32_t lo_i = 2 * k - (k & (stride - 1));
                const uint32_t hi_i = lo_i + stride;
                const bool up = (lo_i & size) == 0;
                const uint32_t a = e_key[lo_i], b = e_key[hi_i];
                if ((a > b) == up) {
                    e_key[lo_i] = b;
                    e_key[hi_i] = a;
                    const uint16_t ta = e_ent[lo_i];
                    e_ent[lo_i] = e_ent[hi_i];
                    e_ent[hi_i] = ta;
                }
            }
            __syncthreads();
        }
    if (tid == 0) s_base = atomicAdd(pool_count, ne);
    __syncthreads();
    const uint32_t b0 = s_base;
    if ((uint64_t)b0 + ne > pool_cap) return; // pool too small: the host sizes it exactly and runs again
    for (uint32_t k = tid; k < ne; k += VA_THREADS) {
        pool_key[b0 + k] = e_key[k];
        pool_ent[b0 + k] = e_ent[k];
    }
    if (tid == 0) {
        cand_ebegin[c] = b0;
        cand_ne[c] = ne | (all_pairs ? 0x80000000u : 0u); // k6b's rescue scan needs to know which pair domain was used
    }
}

// ------------------------------------------------------------------------------------------------
// k6a_table: candidate edges from the store's PAIR TABLE (fd_store_build_pair_table) -- one warp per candidate
// ------------------------------------------------------------------------------------------------
// retrieve_with_prefilter keeps the residue pairs (i, j) of the candidate whose geometric hash is in the query's hash
// set (plus the amino-acid / CA-distance conditions of retrieve.rs:85-143).  k6a_edges finds them by screening n1 x n2
// pairs and re-hashing the survivors -- work the index build already did once per structure.  With the table the
// question is turned round: for every query hash, which pairs of this structure carry it?  One read of the
// structure's amino-acid-pair directory + a binary search inside that pair's run (~75 entries) per query hash; the
// conditions that do not follow from hash equality (exact residue names under the amino-acid prefilter, the pair's
// presence in the observed-distance map, |d - d_q| < --ca-distance) are checked on the few matches.  Same outputs as
// k6a_edges: the candidate's edges sorted by (i, j) in the pool, cand_ne with the pair-domain flag.
constexpr int VT_WARPS = 4;
struct TableWarp { // per-warp shared memory
    uint32_t e_key[V_MAX_E];
    uint16_t e_ent[V_MAX_E];
    VAad aad[V_MAX_AAD];
    uint16_t aa_range[400];
    uint32_t n_e, has1, has2;
};
struct PairTableView {
    const uint64_t *offsets;
    const uint32_t *hash, *ij, *dir;
};

__global__ void __launch_bounds__(VT_WARPS * 32)
    k6a_table(StoreView st, PairTableView pt, const VQDesc *vq, const VHash *vhash, const VAad *vaad,
              const uint32_t *cand_query, const uint32_t *cand_nid, uint32_t n_cand, fdg::HashParams hp, float ca_cutoff,
              uint32_t *pool_key, uint16_t *pool_ent, unsigned int *pool_count, uint32_t pool_cap, uint32_t *cand_ebegin,
              uint32_t *cand_ne, uint8_t *cand_flags) {
    __shared__ TableWarp tw[VT_WARPS];
    const int lane = threadIdx.x & 31;
    TableWarp &W = tw[threadIdx.x >> 5];
    const uint32_t c = blockIdx.x * VT_WARPS + (threadIdx.x >> 5);
    if (c >= n_cand) return;
    const VQDesc Q = vq[cand_query[c]];
    if (Q.n_hashes == 0 || Q.n_aad == 0) return;
    const uint32_t t = cand_nid[c];
    const uint64_t base = st.row_offsets[t];
    const uint32_t n = (uint32_t)(st.row_offsets[t + 1] - base);
    const VHash *H = vhash + Q.hash_begin;
    if (lane == 0) {
        W.n_e = 0;
        W.has1 = 0;
        W.has2 = 0;
    }
    load_aad_phase1<32>(Q, vaad, W.aad, W.aa_range, lane);
    __syncwarp();
    load_aad_phase2<32>(Q, W.aad, W.aa_range, lane);
    // prefilter_amino_acid (retrieve.rs:563-602): with it the pair domain is list1 x list2 (residues whose exact name
    // is an amino acid of the query); an empty list means every pair (CombinationVecIterator::is_empty, :145-151)
    bool all_pairs = !Q.use_prefilter;
    if (!all_pairs) {
        bool any1 = false, any2 = false;
        for (uint32_t r = lane; r < n; r += 32) {
            const uint8_t a = st.aa[base + r];
            const bool canonical = (a & 0x80u) == 0;
            any1 = any1 || (canonical && ((Q.aa1_mask >> (a & 31u)) & 1u));
            any2 = any2 || (canonical && ((Q.aa2_mask >> (a & 31u)) & 1u));
        }
        if (!__any_sync(0xffffffffu, any1) || !__any_sync(0xffffffffu, any2)) all_pairs = true;
    }
    __syncwarp();
    const uint64_t seg = pt.offsets[t];
    const uint32_t *th = pt.hash + seg;
    const uint32_t *tij = pt.ij + seg;
    const uint32_t *dir = pt.dir + (uint64_t)t * FD_PT_DIR;
    for (uint32_t k0 = 0; k0 < Q.n_hashes; k0 += 32) {
        const uint32_t k = k0 + lane;
        if (k < Q.n_hashes) {
            const uint32_t h = H[k].hash;
            const uint32_t p = h >> 20;
            uint32_t lo = dir[p], hi = dir[p + 1];
            const uint32_t end = hi;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (th[mid] < h) lo = mid + 1;
                else hi = mid;
            }
            for (uint32_t e = lo; e < end && th[e] == h; e++) {
                const uint32_t ij = tij[e];
                const uint32_t i = ij >> 16, j = ij & 0xffffu;
                const uint8_t ai = st.aa[base + i], aj = st.aa[base + j];
                if (!all_pairs && ((ai | aj) & 0x80u)) continue; // modified residues are not in the prefilter lists
                const uint32_t rg = W.aa_range[(ai & 0x7Fu) * 20u + (aj & 0x7Fu)];
                if (rg == 0) continue; // the pair's amino acids carry no observed distance (substituted hashes)
                const float d = fdg::dist(ld3(st.ca_xyz, base + i), ld3(st.ca_xyz, base + j));
                bool ok = false;
                for (uint32_t a = rg >> 8, ae = (rg >> 8) + (rg & 0xffu); a < ae; a++)
                    if (fabsf(d - W.aad[a].dist) < ca_cutoff) {
                        ok = true;
                        break;
                    }
                if (!ok) continue;
                const uint32_t pos = atomicAdd(&W.n_e, 1u);
                if (pos < V_MAX_E) {
                    W.e_key[pos] = ij;
                    W.e_ent[pos] = (uint16_t)k;
                }
            }
        }
    }
    __syncwarp();
    const uint32_t ne = W.n_e;
    if (ne == 0) return;
    if (ne > V_MAX_E || Q.n_idx > V_MAX_NQ || Q.n_dq > V_MAX_NQ) {
        if (lane == 0) cand_flags[c] = 2;
        return;
    }
    // ---- sort edges by (i, j): the reference's emission order (graph node numbering, f32 sum order) ----
    uint32_t sort_n = 2;
    while (sort_n < ne) sort_n <<= 1;
    for (uint32_t k = ne + lane; k < sort_n; k += 32) W.e_key[k] = 0xffffffffu;
    __syncwarp();
    for (uint32_t size = 2; size <= sort_n; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t k = lane; k < sort_n / 2; k += 32) {
                const uint32_t lo_i = 2 * k - (k & (stride - 1));
                const uint32_t hi_i = lo_i + stride;
                const bool up = (lo_i & size) == 0;
                const uint32_t a = W.e_key[lo_i], b = W.e_key[hi_i];
                if ((a > b) == up) {
                    W.e_key[lo_i] = b;
                    W.e_key[hi_i] = a;
                    const uint16_t ta = W.e_ent[lo_i];
                    W.e_ent[lo_i] = W.e_ent[hi_i];
                    W.e_ent[hi_i] = ta;
                }
            }
            __syncwarp();
        }
    uint32_t b0 = 0;
    if (lane == 0) b0 = atomicAdd(pool_count, ne);
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    if ((uint64_t)b0 + ne > pool_cap) return; // pool too small: the host sizes it exactly and runs again
    for (uint32_t k = lane; k < ne; k += 32) {
        pool_key[b0 + k] = W.e_key[k];
        pool_ent[b0 + k] = W.e_ent[k];
    }
    if (lane == 0) {
        cand_ebegin[c] = b0;
        cand_ne[c] = ne | (all_pairs ? 0x80000000u : 0u);
    }
}

__global__ void k6_iota(uint32_t *out, uint32_t n) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = k;
}

// ------------------------------------------------------------------------------------------------
// k6b: components, residue mapping, rescue -- one warp per candidate
// ------------------------------------------------------------------------------------------------
struct WarpState { // per-warp shared memory
    uint32_t e_key[V_MAX_E];
    uint16_t e_ent[V_MAX_E];
    uint8_t e_a[V_MAX_E], e_b[V_MAX_E];
    uint64_t comp_mask[V_MAX_C];
    uint16_t node_res[V_MAX_NODES];
    uint32_t counts[V_MAX_NQ * V_MAX_NODES / 2]; // votes[q][node], two 16-bit counters per word
    float e_idf[V_MAX_E];
    uint8_t best_c[V_MAX_NQ], best_r[V_MAX_NQ];
    uint16_t aa_dir[FD_AA_DIR + 2]; // the candidate's amino-acid directory (FdDeviceStore::aa_dir)
    uint32_t dq_aa1[V_MAX_NQ]; // amino acids (bit set) that carry an entry of query residue dq: rows of the rescue scan
    uint32_t n_nodes, n_comp, s_flag;
    uint32_t cand, ne_word; // the warp's own candidate and its edge-count word (owner state, like the arrays above)
    float pre_hi2, pre_lo2; // squared CA-distance window outside which a pair cannot support a rescue (owner state)
    uint32_t r_dq, r_need, r_nridx, s_out_base;
    uint16_t r_ridx[V_MAX_NQ];
    float r_x[V_MAX_NQ], r_y[V_MAX_NQ], r_z[V_MAX_NQ]; // CA of the matched target residues (rescue)
    uint8_t r_aa[V_MAX_NQ];                            // their amino acid, 0xff = cannot pair
    uint8_t m_qidx[V_MAX_NQ], m_ridx[V_MAX_NQ]; // mapping of the current component: dense query id, node id
    uint32_t m_n;
    uint32_t f_qscan[V_MAX_NQ], f_rscan[V_MAX_NQ], f_nscan, f_res[V_MAX_NQ], f_nres;
    uint32_t c_res[V_MAX_NQ]; // res_vec_from_hash of the current component
    float c_idf;
};

__global__ void __launch_bounds__(VB_WARPS * 32)
    k6b_components(StoreView st, const VQDesc *vq, const VHash *vhash, const VAad *vaad, const uint16_t *aa_ranges,
                   const uint8_t *idx_dense,
                   const uint32_t *cand_query, const uint32_t *cand_nid, uint32_t n_cand, fdg::HashParams hp,
                   float ca_cutoff, int skip_ca_match, const uint32_t *pool_key, const uint16_t *pool_ent,
                   const uint32_t *cand_ebegin, const uint32_t *cand_ne, CompSpec *specs, unsigned int *spec_count,
                   uint32_t spec_cap, uint32_t *cand_ncomp, uint8_t *cand_flags, const uint32_t *order) {
    extern __shared__ __align__(16) unsigned char k6b_smem[];
    __shared__ uint32_t s_next; // next (candidate, component) work item of the CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpState *states = reinterpret_cast<WarpState *>(k6b_smem);
    WarpState &W = states[warp];
    if (threadIdx.x == 0) s_next = 0;
    // ---- phase 1: every warp builds the graph of its own candidate (edges, node numbering, component masks) ----
    const uint32_t slot = blockIdx.x * VB_WARPS + warp;
    // candidates are taken in descending order of their edge count (longest first: a candidate with dozens of
    // components that started last kept one SM busy for half of the launch, profiles/r03h_ncu_k6.txt)
    const uint32_t c_own = slot < n_cand ? (order ? order[slot] : slot) : 0u;
    const uint32_t ne_word_own = slot < n_cand ? cand_ne[c_own] : 0u;
    if (lane == 0) {
        W.s_flag = 0;
        W.n_comp = 0;
        W.cand = c_own;
        W.ne_word = ne_word_own;
    }
    __syncwarp();
    if ((ne_word_own & 0x7fffffffu) != 0) {
    const uint32_t c = c_own;
    const uint32_t ne = ne_word_own & 0x7fffffffu;
    const VQDesc Q = vq[cand_query[c]];
    const uint32_t t = cand_nid[c];
    const VHash *H = vhash + Q.hash_begin;
    const uint32_t eb = cand_ebegin[c];
    for (uint32_t k = lane; k < ne; k += 32) {
        const uint16_t ent = pool_ent[eb + k];
        W.e_key[k] = pool_key[eb + k];
        W.e_ent[k] = ent;
        W.e_idf[k] = H[ent].idf;
    }
    const VAad *AAD = vaad + Q.aad_begin;
    if (lane < V_MAX_NQ) W.dq_aa1[lane] = 0;
    for (uint32_t b = lane; b < FD_AA_DIR; b += 32) W.aa_dir[b] = st.aa_dir[(uint64_t)t * FD_AA_DIR + b];
    __syncwarp();
    for (uint32_t k = lane; k < Q.n_aad; k += 32)
        if (AAD[k].dq < V_MAX_NQ) atomicOr(&W.dq_aa1[AAD[k].dq], 1u << (AAD[k].aa1 & 31u));
    { // range of the query's CA distances: a pair outside it cannot support a rescue
        float dmax = 0.f, dmin = 3.0e38f;
        for (uint32_t k = lane; k < Q.n_aad; k += 32) {
            dmax = fmaxf(dmax, AAD[k].dist);
            dmin = fminf(dmin, AAD[k].dist);
        }
        for (int o = 16; o > 0; o >>= 1) {
            dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
            dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        }
        const float pre_hi = fminf(hp.dist_cutoff, dmax + ca_cutoff), pre_lo = fmaxf(dmin - ca_cutoff, 0.f);
        if (lane == 0) {
            W.pre_hi2 = pre_hi * pre_hi * 1.0001f;
            W.pre_lo2 = pre_lo * pre_lo * 0.9999f;
        }
    }
    __syncwarp();

    // ---- graph: nodes by first appearance, components = SCCs U weak components, size >= 2 ----
    // All lanes cooperate: lane v keeps nodes v and v + 32 in registers (residue index, directed and undirected
    // adjacency -> closure).  Node numbering = first appearance in the (i, j)-sorted edge list (graph.rs:16-31).
    {
        const uint32_t FULL = 0xffffffffu;
        uint32_t nn = 0;
        bool overflow = false;
        uint32_t nr0 = 0xffffffffu, nr1 = 0xffffffffu; // residue of node lane / lane + 32 (edge ends are < 65536)
        uint64_t r0 = 0, r1 = 0, u0 = 0, u1 = 0;
        for (uint32_t k = 0; k < ne && !overflow; k++) {
            const uint32_t key = W.e_key[k];
            uint32_t ids[2] = {0, 0};
            for (int s = 0; s < 2; s++) {
                const uint32_t end = s ? (key & 0xffffu) : (key >> 16);
                const uint32_t m0 = __ballot_sync(FULL, nr0 == end), m1 = __ballot_sync(FULL, nr1 == end);
                uint32_t f;
                if (m0) f = (uint32_t)__ffs((int)m0) - 1u;
                else if (m1) f = 31u + (uint32_t)__ffs((int)m1);
                else {
                    if (nn == V_MAX_NODES) {
                        overflow = true;
                        break;
                    }
                    f = nn;
                    if ((uint32_t)lane == (nn & 31u)) {
                        if (nn < 32) nr0 = end;
                        else nr1 = end;
                    }
                    nn++;
                }
                ids[s] = f;
            }
            if (overflow) break;
            const uint32_t a = ids[0], b = ids[1];
            if ((uint32_t)lane == (a & 31u)) {
                if (a < 32) {
                    r0 |= 1ull << b;
                    u0 |= 1ull << b;
                } else {
                    r1 |= 1ull << b;
                    u1 |= 1ull << b;
                }
            }
            if ((uint32_t)lane == (b & 31u)) {
                if (b < 32) u0 |= 1ull << a;
                else u1 |= 1ull << a;
            }
            if (lane == 0) {
                W.e_a[k] = (uint8_t)a;
                W.e_b[k] = (uint8_t)b;
            }
        }
        if (overflow) {
            if (lane == 0) W.s_flag = 4; // more than 64 graph nodes
        } else {
            if ((uint32_t)lane < nn) {
                W.node_res[lane] = (uint16_t)nr0;
                r0 |= 1ull << lane;
                u0 |= 1ull << lane;
            }
            if ((uint32_t)lane + 32 < nn) {
                W.node_res[lane + 32] = (uint16_t)nr1;
                r1 |= 1ull << (lane + 32);
                u1 |= 1ull << (lane + 32);
            }
            // Warshall closure, one pass over the pivot k (rows are 64-bit masks)
            for (uint32_t k = 0; k < nn; k++) {
                const unsigned long long rk = __shfl_sync(FULL, (unsigned long long)(k < 32 ? r0 : r1), (int)(k & 31u));
                const unsigned long long uk = __shfl_sync(FULL, (unsigned long long)(k < 32 ? u0 : u1), (int)(k & 31u));
                if ((r0 >> k) & 1ull) r0 |= rk;
                if ((r1 >> k) & 1ull) r1 |= rk;
                if ((u0 >> k) & 1ull) u0 |= uk;
                if ((u1 >> k) & 1ull) u1 |= uk;
            }
            // strongly connected component of v: the w reachable from v that reach v
            uint64_t s0 = 0, s1 = 0;
            for (uint32_t k = 0; k < nn; k++) {
                const unsigned long long rk = __shfl_sync(FULL, (unsigned long long)(k < 32 ? r0 : r1), (int)(k & 31u));
                if (((r0 >> k) & 1ull) && ((rk >> lane) & 1ull)) s0 |= 1ull << k;
                if (((r1 >> k) & 1ull) && ((rk >> (lane + 32)) & 1ull)) s1 |= 1ull << k;
            }
            // distinct masks of size >= 2: every component is added by its lowest node; a weak component that is
            // strongly connected equals its SCC and is added once
            const bool v0 = (uint32_t)lane < nn, v1 = (uint32_t)lane + 32 < nn;
            const bool sr0 = v0 && __popcll(s0) >= 2 && ctz64(s0) == lane;
            const bool sr1 = v1 && __popcll(s1) >= 2 && ctz64(s1) == lane + 32;
            const bool ur0 = v0 && __popcll(u0) >= 2 && ctz64(u0) == lane && !(sr0 && s0 == u0);
            const bool ur1 = v1 && __popcll(u1) >= 2 && ctz64(u1) == lane + 32 && !(sr1 && s1 == u1);
            const uint32_t bs0 = __ballot_sync(FULL, sr0), bs1 = __ballot_sync(FULL, sr1);
            const uint32_t bu0 = __ballot_sync(FULL, ur0), bu1 = __ballot_sync(FULL, ur1);
            const uint32_t nc = __popc(bs0) + __popc(bs1) + __popc(bu0) + __popc(bu1);
            if (nc > V_MAX_C) {
                if (lane == 0) W.s_flag = 8; // more components than V_MAX_C (cannot happen with <= 64 nodes)
            } else {
                const uint32_t lt = (1u << lane) - 1u;
                uint32_t pos = __popc(bs0 & lt);
                if (sr0) W.comp_mask[pos] = s0;
                pos = __popc(bs0) + __popc(bs1 & lt);
                if (sr1) W.comp_mask[pos] = s1;
                pos = __popc(bs0) + __popc(bs1) + __popc(bu0 & lt);
                if (ur0) W.comp_mask[pos] = u0;
                pos = __popc(bs0) + __popc(bs1) + __popc(bu0) + __popc(bu1 & lt);
                if (ur1) W.comp_mask[pos] = u1;
                __syncwarp();
                if (lane == 0) {
                    for (uint32_t a = 1; a < nc; a++) { // insertion sort, lexicographic by sorted node list
                        const uint64_t m = W.comp_mask[a];
                        uint32_t b = a;
                        while (b > 0 && mask_less(m, W.comp_mask[b - 1])) {
                            W.comp_mask[b] = W.comp_mask[b - 1];
                            b--;
                        }
                        W.comp_mask[b] = m;
                    }
                    W.n_nodes = nn;
                    W.n_comp = nc;
                }
            }
        }
    }
    __syncwarp();
    if (W.s_flag) {
        if (lane == 0) {
            cand_flags[c] = (uint8_t)W.s_flag;
            W.n_comp = 0;
        }
    } else if (lane == 0 && W.n_comp) {
        W.s_out_base = atomicAdd(spec_count, W.n_comp);
        cand_ncomp[c] = W.n_comp;
    }
    } // phase 1
    __syncthreads();

    // ---- phase 2: the components of the CTA's candidates are work items shared by its warps (a candidate with
    // dozens of components no longer runs them one after another on a single warp); the component's owner state
    // O (edges, nodes, masks) is read-only here, the mapping / rescue scratch is this warp's own W ----
    uint32_t comp_begin[VB_WARPS + 1];
    comp_begin[0] = 0;
#pragma unroll
    for (int w = 0; w < VB_WARPS; w++) comp_begin[w + 1] = comp_begin[w] + states[w].n_comp;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(&s_next, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= comp_begin[VB_WARPS]) break;
        int ow = 0;
#pragma unroll
        for (int w = 1; w < VB_WARPS; w++)
            if (item >= comp_begin[w]) ow = w;
        const WarpState &O = states[ow];
        const uint32_t ci = item - comp_begin[ow];
        const uint32_t c = O.cand;
        const uint32_t ne = O.ne_word & 0x7fffffffu;
        const bool all_pairs = (O.ne_word >> 31) != 0;
        const VQDesc Q = vq[cand_query[c]];
        const uint32_t t = cand_nid[c];
        const uint64_t base = st.row_offsets[t];
        const VHash *H = vhash + Q.hash_begin;
        const VAad *AAD = vaad + Q.aad_begin; // (amino-acid pair -> CA distances) of the query: global, L1-resident
        const uint16_t *AAR = aa_ranges + (size_t)cand_query[c] * 400u;
        const uint8_t *IDX = idx_dense + Q.idx_begin;
        const uint32_t out_base = O.s_out_base;
        const float pre_hi2 = O.pre_hi2, pre_lo2 = O.pre_lo2;
        // ---- mapping: votes (lanes over edges), best per query residue (lane per residue), greedy assignment ----
        const uint64_t mask = O.comp_mask[ci];
        for (uint32_t k = lane; k < Q.n_dq * (V_MAX_NODES / 2); k += 32) W.counts[k] = 0;
        __syncwarp();
        // votes[q][r]: 16-bit counters, two per word (a cell receives at most 2 * V_MAX_E votes); the reference's u8
        // saturating counters (retrieve.rs:628-660) are min(count, 255) of these
        for (uint32_t k = lane; k < ne; k += 32) {
            const uint32_t a = O.e_a[k], b = O.e_b[k];
            if (!((mask >> a) & 1ull) || !((mask >> b) & 1ull)) continue;
            const VHash h = H[O.e_ent[k]];
            uint32_t pq[2], pr[2];
            if (h.sym) {
                pq[0] = min((uint32_t)h.dqi, (uint32_t)h.dqj);
                pq[1] = max((uint32_t)h.dqi, (uint32_t)h.dqj);
                const bool ab = O.node_res[a] < O.node_res[b];
                pr[0] = ab ? a : b;
                pr[1] = ab ? b : a;
            } else {
                pq[0] = h.dqi;
                pq[1] = h.dqj;
                pr[0] = a;
                pr[1] = b;
            }
            for (int s = 0; s < 2; s++) {
                const uint32_t cell = pq[s] * V_MAX_NODES + pr[s];
                atomicAdd(&W.counts[cell >> 1], 1u << (16u * (cell & 1u)));
            }
        }
        __syncwarp();
        // best[q] = (largest count, smallest target residue among the nodes that reach it): the state the reference's
        // running update `cnt > best.0 || (cnt == best.0 && r < best.1)` ends in, whatever the edge order
        if ((uint32_t)lane < Q.n_dq) {
            uint32_t bc = 0, br = 0xff, bres = 0xffffffffu;
            for (uint64_t m = mask; m; m &= m - 1) {
                const uint32_t v = (uint32_t)ctz64(m);
                const uint32_t cell = (uint32_t)lane * V_MAX_NODES + v;
                const uint32_t cnt = min((W.counts[cell >> 1] >> (16u * (cell & 1u))) & 0xffffu, 255u);
                if (cnt == 0) continue;
                const uint32_t res = O.node_res[v];
                if (cnt > bc || (cnt == bc && res < bres)) {
                    bc = cnt;
                    br = v;
                    bres = res;
                }
            }
            W.best_c[lane] = (uint8_t)bc;
            W.best_r[lane] = (uint8_t)br;
        }
        __syncwarp();
        if (lane == 0) {
            const uint32_t nq_d = Q.n_dq;
            const uint8_t *best_c = W.best_c, *best_r = W.best_r;
            float idf = 0.f; // calculate_subgraph_idf: f32 sum in edge order
            for (uint32_t k = 0; k < ne; k++) {
                const uint32_t a = O.e_a[k], b = O.e_b[k];
                if (((mask >> a) & 1ull) && ((mask >> b) & 1ull)) idf += O.e_idf[k];
            }
            W.c_idf = idf;
            // order: count descending, dense query id ascending (bucket sort of retrieve.rs:667-675)
            uint8_t order[V_MAX_NQ];
            uint32_t no = 0;
            for (uint32_t q = 0; q < nq_d; q++)
                if (best_c[q] > 0) {
                    uint32_t b = no++;
                    while (b > 0 && best_c[order[b - 1]] < best_c[q]) {
                        order[b] = order[b - 1];
                        b--;
                    }
                    order[b] = (uint8_t)q;
                }
            const uint32_t node_count = (uint32_t)__popcll(mask);
            uint64_t r_used = 0;
            uint32_t nm = 0;
            for (uint32_t k = 0; k < no && nm < node_count; k++) {
                const uint32_t q = order[k], r = best_r[q];
                if (!((r_used >> r) & 1ull)) { // q_used is implied: each q appears once in `order`
                    W.m_qidx[nm] = (uint8_t)q;
                    W.m_ridx[nm] = (uint8_t)r;
                    r_used |= 1ull << r;
                    nm++;
                }
            }
            W.m_n = nm;
            W.f_nscan = 0;
            W.f_nres = 0;
            W.r_nridx = nm;
            for (uint32_t k = 0; k < nm; k++) W.r_ridx[k] = O.node_res[W.m_ridx[k]];
        }
        __syncwarp();
        if ((uint32_t)lane < W.r_nridx) { // the matched residues' data, once per component
            const uint32_t rj = W.r_ridx[lane];
            const uint8_t aj = st.aa[base + rj];
            bool ok = aj != 255 && (st.cb_valid == nullptr || st.cb_valid[base + rj]);
            if (!all_pairs) ok = ok && ((aj & 0x80u) == 0) && ((Q.aa2_mask >> (aj & 31u)) & 1u);
            const fdg::V3 c = ld3(st.ca_xyz, base + rj);
            W.r_x[lane] = c.x;
            W.r_y[lane] = c.y;
            W.r_z[lane] = c.z;
            W.r_aa[lane] = ok ? (uint8_t)(aj & 0x7Fu) : (uint8_t)0xff;
        }
        __syncwarp();
        // ---- rescue loop over the query residues (retrieve.rs:453-516) ----
        for (uint32_t pos = 0; pos < Q.n_idx; pos++) {
            if (lane == 0) {
                const uint32_t dq = IDX[pos];
                int mapped = -1;
                for (uint32_t k = 0; k < W.m_n; k++)
                    if (W.m_qidx[k] == dq) mapped = (int)O.node_res[W.m_ridx[k]];
                W.r_need = 0;
                if (mapped >= 0) {
                    const uint32_t ri = (uint32_t)mapped;
                    W.c_res[pos] = ri + 1; // res_vec_from_hash
                    uint32_t pp = 0;
                    while (pp < W.f_nscan && W.f_rscan[pp] != ri) pp++;
                    if (pp == W.f_nscan) {
                        W.f_res[W.f_nres++] = ri + 1;
                        W.f_qscan[W.f_nscan] = dq;
                        W.f_rscan[W.f_nscan++] = ri;
                    } else {
                        W.f_res[pp] = 0; // sic (retrieve.rs:475)
                        W.f_res[W.f_nres++] = ri + 1;
                        for (uint32_t k = pp; k + 1 < W.f_nscan; k++) {
                            W.f_qscan[k] = W.f_qscan[k + 1];
                            W.f_rscan[k] = W.f_rscan[k + 1];
                        }
                        W.f_qscan[W.f_nscan - 1] = dq;
                        W.f_rscan[W.f_nscan - 1] = ri;
                    }
                } else {
                    W.c_res[pos] = 0;
                    W.r_need = 1;
                    W.r_dq = dq;
                }
            }
            __syncwarp();
            if (W.r_need) {
                // count_map[i] = #(entries of this query residue, matched target residues rj) compatible with (i, rj),
                // over the rows of the pair iteration (all residues, or those whose amino acid is a res1 of the query)
                const uint32_t dq = W.r_dq;
                const uint32_t row_aa = O.dq_aa1[dq]; // an entry (aa1, aa2, dq) only matches rows whose amino acid is aa1
                auto row_count = [&](uint32_t i) -> uint32_t {
                    const uint32_t cia = (st.aa[base + i] & 0x7Fu) * 20u;
                    const fdg::V3 cai = ld3(st.ca_xyz, base + i);
                    uint32_t cnt = 0;
                    for (uint32_t k = 0; k < W.r_nridx; k++) {
                        const uint32_t aj = W.r_aa[k];
                        if (aj == 0xffu || W.r_ridx[k] == i) continue;
                        const uint32_t rg = AAR[cia + aj];
                        if (rg == 0) continue;
                        const float d2 = fdg::dist2(cai, fdg::V3{W.r_x[k], W.r_y[k], W.r_z[k]});
                        if (d2 > pre_hi2 || d2 < pre_lo2) continue;
                        const float d = FD_SQRT(d2); // == fdg::dist
                        if (!(d <= hp.dist_cutoff)) continue;
                        for (uint32_t e = rg >> 8, ee = (rg >> 8) + (rg & 0xffu); e < ee; e++)
                            if (AAD[e].dq == dq && fabsf(d - AAD[e].dist) < ca_cutoff) cnt++;
                    }
                    return cnt;
                };
                // per lane the maximum count, how many rows reach it, and one such row
                uint32_t l_max = 0, l_n = 0, l_arg = 0;
                auto take = [&](uint32_t i) {
                    const uint32_t cnt = row_count(i);
                    if (cnt > l_max) {
                        l_max = cnt;
                        l_n = 1;
                        l_arg = i;
                    } else if (cnt == l_max && cnt > 0) {
                        l_n++;
                    }
                };
                // The rows that can count at all -- amino acid carries an entry of dq (in practice ONE amino acid: that
                // of the query residue, plus its substitutions), in the pair iteration, has a CB -- come from the
                // store's amino-acid directory: ~n / 20 rows instead of a scan over all n.
                for (uint32_t am = row_aa & 0xfffffu; am; am &= am - 1u) {
                    const uint32_t a = (uint32_t)__ffs((int)am) - 1u;
                    for (uint32_t mod = 0; mod < (all_pairs ? 2u : 1u); mod++) {
                        if (!all_pairs && !((Q.aa1_mask >> a) & 1u)) continue;
                        const uint32_t b = a + 20u * mod;
                        for (uint32_t r = O.aa_dir[b] + lane, re = O.aa_dir[b + 1]; r < re; r += 32) take(st.aa_rows[base + r]);
                    }
                }
                const uint32_t mx = __reduce_max_sync(0xffffffffu, l_max);
                const uint32_t mine = (mx > 0 && l_max == mx) ? l_n : 0u;
                const uint32_t nmax = __reduce_add_sync(0xffffffffu, mine);
                const uint32_t who = __ballot_sync(0xffffffffu, mine != 0);
                const uint32_t arg = who ? __shfl_sync(0xffffffffu, l_arg, __ffs(who) - 1) : 0u;
                if (lane == 0) {
                    bool ok = mx >= 2 && nmax == 1;
                    if (ok)
                        for (uint32_t k = 0; k < W.f_nscan; k++)
                            if (W.f_rscan[k] == arg) ok = false;
                    if (ok) {
                        W.f_res[W.f_nres++] = arg + 1;
                        W.f_qscan[W.f_nscan] = W.r_dq;
                        W.f_rscan[W.f_nscan++] = arg;
                    } else {
                        W.f_res[W.f_nres++] = 0;
                    }
                }
            }
            __syncwarp();
        }
        // ---- choose the alignment and the reported residues; write the component spec ----
        if (lane == 0) {
            bool same = true;
            for (uint32_t k = 0; k < Q.n_idx; k++) same = same && W.f_res[k] == W.c_res[k];
            const uint32_t slot = out_base + ci;
            if (slot < spec_cap) {
                CompSpec sp;
                sp.cand = c;
                sp.ci = (uint16_t)ci;
                sp.idf = W.c_idf;
                sp.pad = 0;
                uint32_t nal;
                if (skip_ca_match || same) {
                    nal = W.m_n;
                    for (uint32_t k = 0; k < nal; k++) {
                        sp.aq[k] = W.m_qidx[k];
                        sp.at[k] = O.node_res[W.m_ridx[k]];
                    }
                } else {
                    nal = W.f_nscan;
                    for (uint32_t k = 0; k < nal; k++) {
                        sp.aq[k] = (uint8_t)W.f_qscan[k];
                        sp.at[k] = (uint16_t)W.f_rscan[k];
                    }
                }
                for (uint32_t k = nal; k < V_MAX_NQ; k++) {
                    sp.aq[k] = 0;
                    sp.at[k] = 0;
                }
                sp.nal = (uint16_t)nal;
                for (uint32_t k = 0; k < V_MAX_NQ; k++)
                    sp.res[k] = k < Q.n_idx ? (skip_ca_match ? W.c_res[k] : W.f_res[k]) : 0u;
                specs[slot] = sp;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// k6c: Kabsch per component, records in (candidate, component) order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    k6c_kabsch(StoreView st, const VQDesc *vq, const float *q_ca, const float *q_cb, const uint32_t *cand_query,
               const uint32_t *cand_nid, const CompSpec *specs, const unsigned int *spec_count, uint32_t spec_cap,
               uint32_t cand_base, const uint32_t *cand_first, fd_match_record *out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= min(*spec_count, spec_cap)) return; // the grid is sized for the capacity: no host round trip for the count
    const CompSpec sp = specs[k];
    const VQDesc Q = vq[cand_query[sp.cand]];
    const uint64_t base = st.row_offsets[cand_nid[sp.cand]];
    fd_match_record rec;
    rec.cand = sp.cand + cand_base;
    rec.idf = sp.idf;
    uint32_t nodes = 0;
    for (uint32_t j = 0; j < V_MAX_NQ; j++) {
        rec.res[j] = sp.res[j];
        nodes += sp.res[j] != 0;
    }
    rec.node_count = nodes;
    GatherPointsT<uint16_t> mov{st.ca_xyz, st.cb_xyz, sp.at, base};
    GatherPointsT<uint8_t> ref{q_ca, q_cb, sp.aq, (uint64_t)Q.qres_base};
    fdk::kabsch_one(mov, ref, 2u * sp.nal, rec.U, rec.t, &rec.rmsd);
    const uint32_t pos = cand_first[sp.cand] + sp.ci;
    if (pos < spec_cap) out[pos] = rec; // (an overflowing chunk is re-issued with exact sizes)
}

} // namespace

// The query side of the verification (independent of the candidates): tables flattened on the host and resident on
// the device.  Built once per query batch (fd_verify_prepare) and reused by every verification call on it.
struct fd_verify_prepared {
    uint32_t nq = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<uint8_t> q_unfit; // queries outside the kernels' limits: their candidates take the general path
    VQDesc *d_desc = nullptr;
    VHash *d_hash = nullptr;
    VAad *d_aad = nullptr;
    uint16_t *d_aar = nullptr; // [nq][400] run of every amino-acid pair in d_aad: start << 8 | count (0 = pair not in the query)
    uint8_t *d_idx = nullptr;
    float *d_qca = nullptr, *d_qcb = nullptr;
    uint64_t h2d_bytes = 0;
};

static void verify_prepared_release(fd_verify_prepared *P) {
    if (!P) return;
    cudaSetDevice(P->device);
    void *ptrs[7] = {P->d_desc, P->d_hash, P->d_aad, P->d_idx, P->d_qca, P->d_qcb, P->d_aar};
    // The tables may outlive the context (and stream) that uploaded them, and every search that used them has
    // returned: they go back to the device's pool through the default stream (cudaFree would synchronise the whole
    // device at every released batch).
    for (void *q : ptrs)
        if (q && cudaFreeAsync(q, (cudaStream_t)0) != cudaSuccess) cudaFree(q);
    cudaGetLastError();
    delete P;
}

static int verify_prepare(fd_ctx *ctx, const fd_verify_query *queries, uint32_t nq, fd_verify_prepared **out) {
    if (!ctx || !out) return FD_ERR_ARG;
    *out = nullptr;
    if (nq && !queries) return fd_fail(ctx, FD_ERR_ARG, "fd_verify_prepare: NULL argument");
    FD_ENTER(ctx);
    auto h_now = std::chrono::steady_clock::now();
    auto h_mark = [&](const char *name) {
        const auto t = std::chrono::steady_clock::now();
        ctx->stages[name].ms += std::chrono::duration<double, std::milli>(t - h_now).count();
        h_now = t;
    };
    // flatten queries (two passes, query-parallel, no per-query heap traffic); a query outside the kernel's limits
    // marks all of its candidates for the general path
    std::vector<VQDesc> descs(nq);
    std::vector<uint8_t> q_unfit(nq, 0);
    struct DenseIds { // sorted distinct residue indices of one query (at most V_MAX_NQ when the query fits)
        uint32_t v[V_MAX_NQ + 1];
        uint32_t n;
    };
    std::vector<DenseIds> dqs(nq);
    std::atomic<int> bad{0}; // 1 hashes not ascending, 2 amino-acid code, 3 residue index
    auto parallel_queries = [&](const std::function<void(uint32_t)> &fn) {
        int nt = std::min(fd_default_host_threads(), 16);
        if (nq < 64) nt = 1;
        std::atomic<uint32_t> next{0};
        auto worker = [&] {
            for (uint32_t q0; (q0 = next.fetch_add(16)) < nq;)
                for (uint32_t q = q0; q < std::min(nq, q0 + 16); q++) fn(q);
        };
        fd_parallel(nt, [&](int) { worker(); });
    };
    // pass 1: dense ids over every query residue index that occurs (ascending residue index), limits, masks
    parallel_queries([&](uint32_t q) {
        const fd_verify_query &Q = queries[q];
        DenseIds &D = dqs[q];
        D.n = 0;
        bool overflow = false;
        auto add = [&](uint32_t r) {
            if (overflow) return;
            uint32_t p = 0;
            while (p < D.n && D.v[p] < r) p++;
            if (p < D.n && D.v[p] == r) return;
            if (D.n == V_MAX_NQ + 1) {
                overflow = true;
                return;
            }
            for (uint32_t k = D.n; k > p; k--) D.v[k] = D.v[k - 1];
            D.v[p] = r;
            D.n++;
        };
        for (uint32_t k = 0; k < Q.n_hashes; k++) {
            add(Q.hash_qi[k]);
            add(Q.hash_qj[k]);
        }
        for (uint32_t k = 0; k < Q.n_indices; k++) add(Q.indices[k]);
        for (uint32_t k = 0; k < Q.n_aa_dist; k++) add(Q.q_index[k]);
        VQDesc d{};
        d.n_hashes = Q.n_hashes;
        d.n_aad = Q.n_aa_dist;
        d.use_prefilter = Q.n_hashes <= V_PREFILTER_SKIP ? 1u : 0u;
        d.n_idx = Q.n_indices;
        d.n_dq = D.n;
        if (overflow || D.n > V_MAX_NQ || Q.n_indices > V_MAX_NQ || Q.n_aa_dist > V_MAX_AAD || Q.n_hashes > 65535) {
            q_unfit[q] = 1;
            d.n_aad = 0; // the kernels return immediately for this query's candidates
            D.n = std::min<uint32_t>(D.n, V_MAX_NQ);
            d.n_dq = D.n;
        }
        for (uint32_t k = 0; k < Q.n_hashes; k++) {
            if (k && Q.hashes_sorted[k] <= Q.hashes_sorted[k - 1]) bad = 1;
            d.aa1_mask |= 1u << ((Q.hashes_sorted[k] >> 25) & 31u);
            d.aa2_mask |= 1u << ((Q.hashes_sorted[k] >> 20) & 31u);
        }
        for (uint32_t k = 0; k < D.n; k++)
            if (D.v[k] >= Q.n_residues) bad = 3;
        descs[q] = d;
    });
    if (bad == 1) return fd_fail(ctx, FD_ERR_ARG, "fd_verify_query: hashes_sorted must be strictly ascending");
    if (bad == 3) return fd_fail(ctx, FD_ERR_ARG, "fd_verify_query: residue index outside the query structure");
    uint64_t n_fh = 0, n_fa = 0, n_fi = 0, n_fr = 0;
    for (uint32_t q = 0; q < nq; q++) {
        VQDesc &d = descs[q];
        d.hash_begin = (uint32_t)n_fh;
        d.aad_begin = (uint32_t)n_fa;
        d.idx_begin = (uint32_t)n_fi;
        d.qres_base = (uint32_t)n_fr;
        n_fh += d.n_hashes;
        n_fa += d.n_aad;
        n_fi += d.n_idx;
        n_fr += d.n_dq;
    }
    if (n_fh > 0xffffffffull || n_fr > 0xffffffffull) return fd_fail(ctx, FD_ERR_LIMIT, "query batch too large");
    std::vector<VHash> f_hash(n_fh);
    std::vector<VAad> f_aad(n_fa);
    std::vector<uint16_t> f_aar((size_t)nq * 400, 0);
    std::vector<uint8_t> f_idx(n_fi);
    std::vector<float> q_ca(3 * n_fr), q_cb(3 * n_fr);
    // pass 2: fill
    parallel_queries([&](uint32_t q) {
        const fd_verify_query &Q = queries[q];
        const DenseIds &D = dqs[q];
        const VQDesc &d = descs[q];
        auto dense = [&](uint32_t r) {
            uint32_t p = 0;
            while (p < D.n && D.v[p] < r) p++;
            return (uint8_t)p;
        };
        for (uint32_t k = 0; k < Q.n_hashes; k++)
            f_hash[d.hash_begin + k] = VHash{Q.hashes_sorted[k], Q.hash_idf[k], dense(Q.hash_qi[k]), dense(Q.hash_qj[k]),
                                             Q.hash_symmetric[k], 0};
        if (!q_unfit[q]) { // entries grouped by amino-acid pair (the kernels index them through a 20x20 table):
            // stable counting sort by aa1 * 20 + aa2
            uint16_t start[401];
            memset(start, 0, sizeof(start));
            for (uint32_t k = 0; k < Q.n_aa_dist; k++) {
                if (Q.aa1[k] >= 20 || Q.aa2[k] >= 20) {
                    bad = 2;
                    return;
                }
                start[Q.aa1[k] * 20u + Q.aa2[k] + 1]++;
            }
            for (int k = 0; k < 400; k++) start[k + 1] += start[k];
            for (int k = 0; k < 400; k++) // at most V_MAX_AAD = 255 entries per query: start and count fit 8 bits each
                if (start[k + 1] > start[k]) f_aar[(size_t)q * 400 + k] = (uint16_t)((start[k] << 8) | (start[k + 1] - start[k]));
            for (uint32_t k = 0; k < Q.n_aa_dist; k++)
                f_aad[d.aad_begin + start[Q.aa1[k] * 20u + Q.aa2[k]]++] =
                    VAad{Q.aa1[k], Q.aa2[k], dense(Q.q_index[k]), 0, Q.ca_dist[k]};
        }
        for (uint32_t k = 0; k < Q.n_indices; k++) f_idx[d.idx_begin + k] = dense(Q.indices[k]);
        for (uint32_t k = 0; k < D.n; k++) { // only the residues the query touches travel to the device
            const uint32_t r = D.v[k];
            for (int x = 0; x < 3; x++) {
                q_ca[3 * ((size_t)d.qres_base + k) + x] = Q.ca_xyz[3 * r + x];
                q_cb[3 * ((size_t)d.qres_base + k) + x] = Q.cb_xyz[3 * r + x];
            }
        }
    });
    if (bad == 2) return fd_fail(ctx, FD_ERR_ARG, "fd_verify_query: amino-acid code out of range");
    h_mark("hv_flatten");
    fd_verify_prepared *P = new fd_verify_prepared();
    P->nq = nq;
    P->device = ctx->device;
    P->stream = ctx->stream;
    P->q_unfit = std::move(q_unfit);
    cudaStream_t s = ctx->stream;
    // the flattened tables travel through one pinned staging buffer (pageable uploads of a few MB ran at ~5 GB/s)
    const size_t stage_bytes = nq * sizeof(VQDesc) + f_hash.size() * sizeof(VHash) + f_aad.size() * sizeof(VAad) +
                               f_aar.size() * sizeof(uint16_t) + f_idx.size() + (q_ca.size() + q_cb.size()) * 4 + 7 * 16;
    uint8_t *stage = nullptr;
    if (fd_pinned(ctx, 8, stage_bytes, (void **)&stage) != FD_OK) {
        delete P;
        return FD_ERR_NOMEM;
    }
    size_t stage_pos = 0;
    auto up = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMallocAsync(dst, std::max<size_t>(bytes, 16), s);
        if (e != cudaSuccess) return e;
        P->h2d_bytes += bytes;
        if (!bytes) return cudaSuccess;
        memcpy(stage + stage_pos, src, bytes);
        e = cudaMemcpyAsync(*dst, stage + stage_pos, bytes, cudaMemcpyHostToDevice, s);
        stage_pos += (bytes + 15) & ~(size_t)15;
        return e;
    };
    cudaError_t e = up((void **)&P->d_desc, descs.data(), nq * sizeof(VQDesc));
    if (e == cudaSuccess) e = up((void **)&P->d_hash, f_hash.data(), f_hash.size() * sizeof(VHash));
    if (e == cudaSuccess) e = up((void **)&P->d_aad, f_aad.data(), f_aad.size() * sizeof(VAad));
    if (e == cudaSuccess) e = up((void **)&P->d_aar, f_aar.data(), f_aar.size() * sizeof(uint16_t));
    if (e == cudaSuccess) e = up((void **)&P->d_idx, f_idx.data(), f_idx.size());
    if (e == cudaSuccess) e = up((void **)&P->d_qca, q_ca.data(), q_ca.size() * 4);
    if (e == cudaSuccess) e = up((void **)&P->d_qcb, q_cb.data(), q_cb.size() * 4);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s); // the host vectors above go out of scope
    if (e != cudaSuccess) {
        verify_prepared_release(P);
        return fd_fail(ctx, FD_ERR_CUDA, std::string("fd_verify_prepare: ") + cudaGetErrorString(e));
    }
    h_mark("hv_upload");
    *out = P;
    return FD_OK;
}


static void verify_keep_release(fd_ctx *ctx) {
    FdVerifyKeep &K = ctx->vkeep;
    if (K.recs) cudaFreeAsync(K.recs, ctx->stream);
    if (K.d_cand_query) cudaFreeAsync(K.d_cand_query, ctx->stream);
    K = FdVerifyKeep();
}

// ------------------------------------------------------------------------------------------------
// k6d: result rows on the device -- one CTA per query.  Replaces the host's row assembly (per-candidate summary of
// retrieve.rs:539-551, the default sorts of sort.rs:218-222 / 454-458, residue labels): the host only receives the
// finished arrays.  Queries the kernel cannot order (pool larger than its shared memory, NaN / -0.0 sort keys) are
// flagged and left in emission order for the host.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t VD_MAX_MATCHES = 2048; // bitonic sort of a query's matches in shared memory
constexpr uint32_t VD_MAX_CANDS = 2048;   // summaries of a query's candidates in shared memory
constexpr int VD_THREADS = 256;

__device__ __forceinline__ uint32_t vd_sortable(float f) { // order-preserving integer image of a float
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ bool vd_plain(float f) { return f == f && !(f == 0.f && (__float_as_uint(f) >> 31)); }

__global__ void __launch_bounds__(VD_THREADS)
    k6d_rows(const VQDesc *vq, const uint32_t *cand_query, const uint64_t *cand_off, const fd_struct_hit *hits,
             const uint32_t *first, const fd_match_record *recs, const uint64_t *store_row_offsets,
             const uint8_t *label_chain, const uint64_t *label_serial, const uint64_t *res_off,
             fd_struct_row *structs_tmp, fd_struct_row *structs, fd_match_row *matches, unsigned long long *match_order, fd_residue_row *residues, uint8_t *needs_host, uint32_t q_base) {
    __shared__ unsigned long long s_key[VD_MAX_MATCHES];
    __shared__ uint16_t s_idx[VD_MAX_MATCHES];
    __shared__ float s_idf[VD_MAX_CANDS], s_rmsd[VD_MAX_CANDS];
    __shared__ uint32_t s_flag;
    const uint32_t q = q_base + blockIdx.x, tid = threadIdx.x;
    const uint64_t c0 = cand_off[q], c1 = cand_off[q + 1];
    const uint32_t nc = (uint32_t)(c1 - c0);
    if (tid == 0) s_flag = 0;
    __syncthreads();
    if (nc == 0) {
        if (tid == 0) needs_host[q] = 0;
        return;
    }
    const uint32_t m0 = first[c0], m1 = first[c1], nm = m1 - m0;
    const uint32_t n_res = vq[cand_query[c0]].n_idx;
    const bool cands_fit = nc <= VD_MAX_CANDS, matches_fit = nm <= VD_MAX_MATCHES;
    // ---- per-candidate summary: max node count, smallest RMSD among the matches that reach it ----
    for (uint32_t k = tid; k < nc; k += blockDim.x) {
        const uint64_t c = c0 + k;
        uint32_t max_node = 0;
        float min_rmsd = 0.f;
        for (uint32_t m = first[c]; m < first[c + 1]; m++) {
            const uint32_t node = recs[m].node_count;
            const float r = recs[m].rmsd;
            if (node > max_node) {
                max_node = node;
                min_rmsd = r;
            } else if (node == max_node && r < min_rmsd) {
                min_rmsd = r;
            }
        }
        const fd_struct_hit h = hits[c];
        if (cands_fit) {
            s_idf[k] = h.idf;
            s_rmsd[k] = min_rmsd;
        }
        if (!vd_plain(h.idf) || !vd_plain(min_rmsd)) atomicOr(&s_flag, 1u);
        // provisional position = count_query order; moved below when the run of equal idf needs reordering
        fd_struct_row sr;
        memset(&sr, 0, sizeof(sr)); // (padding bytes travel to the host)
        sr.nid = h.nid;
        sr.total_match_count = h.match_count;
        sr.node_count = h.node_count;
        sr.edge_count = h.edge_count;
        sr.idf = h.idf;
        sr.max_matching_node_count = max_node;
        sr.min_rmsd_with_max_match = min_rmsd;
        sr.match_begin = first[c];
        sr.match_end = first[c + 1];
        structs_tmp[c] = sr;
    }
    // ---- match rows + matched residues, emission order ----
    for (uint32_t k = tid; k < nm; k += blockDim.x) {
        const uint32_t m = m0 + k;
        const fd_match_record r = recs[m];
        const uint32_t nid = hits[r.cand].nid;
        fd_match_row mr;
        mr.nid = nid;
        mr.node_count = r.node_count;
        mr.idf = r.idf;
        mr.rmsd = r.rmsd;
#pragma unroll
        for (int x = 0; x < 9; x++) mr.U[x] = r.U[x];
#pragma unroll
        for (int x = 0; x < 3; x++) mr.t[x] = r.t[x];
        const uint64_t rb = res_off[q] + (uint64_t)k * n_res;
        mr.res_begin = rb;
        matches[m] = mr;
        const uint64_t base = store_row_offsets[nid];
        for (uint32_t x = 0; x < n_res; x++) {
            const uint32_t v = r.res[x];
            fd_residue_row rr;
            memset(&rr, 0, sizeof(rr));
            rr.some = v != 0;
            rr.serial = v ? (uint64_t)(v - 1) : 0;
            if (v && label_chain) {
                rr.chain = label_chain[base + v - 1];
                rr.serial = label_serial[base + v - 1];
            }
            residues[rb + x] = rr;
        }
        if (!vd_plain(r.idf) || !vd_plain(r.rmsd)) atomicOr(&s_flag, 1u);
        if (matches_fit) {
            s_key[k] = ((unsigned long long)(~vd_sortable(r.idf)) << 32) | vd_sortable(r.rmsd);
            s_idx[k] = (uint16_t)k;
        }
    }
    __syncthreads();
    const bool host_sorts = s_flag != 0 || !cands_fit || !matches_fit;
    if (tid == 0) needs_host[q] = host_sorts ? 1 : 0;
    if (host_sorts) {
        __syncthreads();
        for (uint32_t k = tid; k < nc; k += blockDim.x) structs[c0 + k] = structs_tmp[c0 + k];
        for (uint32_t k = tid; k < nm; k += blockDim.x) match_order[m0 + k] = m0 + k;
        return;
    }
    // ---- structure rows: idf desc, min_rmsd asc, stable.  They arrive idf desc (ties by nid): only runs of equal idf
    // can need reordering (StructureSortStrategy::default, sort.rs:454-458): rank inside the run, copy to the final array
    for (uint32_t k = tid; k < nc; k += blockDim.x) {
        const float idf = s_idf[k], rm = s_rmsd[k];
        uint32_t lo = k, hi = k + 1;
        while (lo > 0 && s_idf[lo - 1] == idf) lo--;
        while (hi < nc && s_idf[hi] == idf) hi++;
        uint32_t rank = k;
        if (hi - lo > 1) {
            rank = lo;
            for (uint32_t j = lo; j < hi; j++)
                if (s_rmsd[j] < rm || (s_rmsd[j] == rm && j < k)) rank++;
        }
        structs[c0 + rank] = structs_tmp[c0 + k];
    }
    // ---- match order: idf desc, rmsd asc, emission index (MatchSortStrategy::default, sort.rs:218-222) ----
    uint32_t mpow = 32;
    while (mpow < nm) mpow <<= 1;
    for (uint32_t k = nm + tid; k < mpow; k += blockDim.x) {
        s_key[k] = ~0ull;
        s_idx[k] = (uint16_t)k;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= mpow; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t k = tid; k < (mpow >> 1); k += blockDim.x) {
                const uint32_t lo_i = 2 * k - (k & (stride - 1));
                const uint32_t hi_i = lo_i + stride;
                const bool up = (lo_i & size) == 0;
                const unsigned long long x = s_key[lo_i], y = s_key[hi_i];
                const uint16_t ix = s_idx[lo_i], iy = s_idx[hi_i];
                const bool gt = x > y || (x == y && ix > iy); // emission index breaks ties: a total order
                if (gt == up) {
                    s_key[lo_i] = y;
                    s_key[hi_i] = x;
                    s_idx[lo_i] = iy;
                    s_idx[hi_i] = ix;
                }
            }
            __syncthreads();
        }
    for (uint32_t k = tid; k < nm; k += blockDim.x) match_order[m0 + k] = (unsigned long long)m0 + s_idx[k];
}

// Shared body of the entry points.  Outputs are views of ctx's pinned staging buffers.
static int verify_run(fd_ctx *ctx, const fd_verify_prepared *P, const uint32_t *cand_query, const uint32_t *cand_nid,
                      uint64_t n_cand, const fd_hash_params *params, float ca_dist_cutoff, int skip_ca_match,
                      const fd_match_record **out_records, uint64_t *out_n, const uint32_t **out_first,
                      const uint8_t **out_flags, bool keep_device = false, fd_rows_plan *plan = nullptr) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->store.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_verify_candidates: no structure store attached");
    if (!P || (n_cand && (!cand_query || !cand_nid)) || !params || !out_records || !out_n || !out_flags || !out_first)
        return fd_fail(ctx, FD_ERR_ARG, "fd_verify_candidates: NULL argument");
    if (P->device != ctx->device) return fd_fail(ctx, FD_ERR_ARG, "fd_verify_candidates: tables prepared on another device");
    if (n_cand > 0xfffffff0ull) return fd_fail(ctx, FD_ERR_LIMIT, "too many candidates in one call");
    FD_ENTER(ctx);
    *out_records = nullptr;
    *out_flags = nullptr;
    *out_first = nullptr;
    *out_n = 0;
    verify_keep_release(ctx);
    auto h_now = std::chrono::steady_clock::now();
    auto h_mark = [&](const char *name) {
        const auto t = std::chrono::steady_clock::now();
        ctx->stages[name].ms += std::chrono::duration<double, std::milli>(t - h_now).count();
        h_now = t;
    };
    const uint32_t nq = P->nq;
    const std::vector<uint8_t> &q_unfit = P->q_unfit;
    uint8_t *h_flags = nullptr;
    uint32_t *h_first = nullptr;
    FD_TRY(fd_pinned(ctx, 1, std::max<uint64_t>(n_cand, 1), (void **)&h_flags));
    FD_TRY(fd_pinned(ctx, 2, (n_cand + 1) * 4, (void **)&h_first));
    memset(h_flags, 0, std::max<uint64_t>(n_cand, 1));
    memset(h_first, 0, (n_cand + 1) * 4);
    for (uint64_t c = 0; c < n_cand; c++) {
        if (cand_query[c] >= nq || cand_nid[c] >= ctx->store.n_structs) {
            return fd_fail(ctx, FD_ERR_ARG, "candidate query / structure id out of range");
        }
    }
    for (uint64_t c = 0; c < n_cand; c++)
        if (q_unfit[cand_query[c]]) h_flags[c] = 1;
    if (n_cand == 0) {
        fd_match_record *none = nullptr;
        FD_TRY(fd_pinned(ctx, 0, sizeof(fd_match_record), (void **)&none));
        *out_records = none;
        *out_flags = h_flags;
        *out_first = h_first;
        return FD_OK;
    }
    // ---- chunked pipeline ----
    // The candidates are cut into a few contiguous chunks issued back to back on two compute streams without any
    // host synchronisation in between (k6c sizes itself from the device-side component count).  The host then walks
    // the chunks in order: waits for a chunk's event, reads its counts and starts the copy of its records on a third
    // stream -- so the record copy of chunk i (PCIe, the longest serial piece of the old sequence) runs under the
    // kernels of chunks i+1...  Pools are sized for 48 edges / 2 components per candidate; a chunk that overflows is
    // re-issued with exact sizes.
    cudaStream_t s0 = ctx->stream;
    for (auto &st : ctx->aux_stream)
        if (!st) FD_CUDA(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaStream_t s1 = ctx->aux_stream[0], sc = ctx->aux_stream[1];
    uint32_t n_chunks = 4;
    if (const char *e = getenv("FD_VERIFY_CHUNKS")) n_chunks = (uint32_t)std::max(1, std::min(atoi(e), 16));
    n_chunks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_chunks, n_cand / 4096));
    auto event_at = [&](size_t k) -> cudaEvent_t {
        while (ctx->ev_pool.size() <= k) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            ctx->ev_pool.push_back(e);
        }
        return ctx->ev_pool[k];
    };
    if (!event_at(5 * n_chunks + 3)) return fd_fail(ctx, FD_ERR_CUDA, "cudaEventCreate failed");
    struct Chunk {
        uint64_t c0 = 0, n = 0;
        cudaStream_t st = nullptr;
        uint64_t pool_cap = 0, spec_cap = 0;
        DevBuf<uint32_t> pool_key, ncomp, first;
        DevBuf<uint16_t> pool_ent;
        DevBuf<unsigned int> counters; // [0] edge pool, [1] component specs
        DevBuf<CompSpec> specs;
        DevBuf<fd_match_record> out;
        DevBuf<uint8_t> tmp, sort_tmp;
        DevBuf<uint32_t> order, ne_sorted; // candidates in descending order of their edge count
        unsigned int *h_counters = nullptr; // pinned
        uint32_t *h_first_rel = nullptr;    // pinned, n + 1
        uint64_t rec_base = 0;
    };
    // with a rows plan the chunks are cut at query boundaries, so that a finished chunk holds whole queries
    std::vector<uint64_t> cut(n_chunks + 1);
    std::vector<uint32_t> cut_q(n_chunks + 1, 0);
    for (uint32_t k = 0; k <= n_chunks; k++) cut[k] = n_cand * k / n_chunks;
    bool rows_on = plan != nullptr;
    if (plan) {
        plan->done = 0;
        if (plan->cand_offsets[0] != 0 || plan->cand_offsets[plan->n_queries] != n_cand)
            return fd_fail(ctx, FD_ERR_ARG, "fd_verify_candidates_rows: cand_offsets must cover the candidates");
        std::vector<uint64_t> c2{0};
        std::vector<uint32_t> q2{0};
        for (uint32_t k = 1; k < n_chunks; k++) {
            const uint64_t *lo = std::lower_bound(plan->cand_offsets, plan->cand_offsets + plan->n_queries + 1, cut[k]);
            const uint32_t q = (uint32_t)(lo - plan->cand_offsets);
            if (*lo > c2.back() && *lo < n_cand) {
                c2.push_back(*lo);
                q2.push_back(q);
            }
        }
        c2.push_back(n_cand);
        q2.push_back(plan->n_queries);
        n_chunks = (uint32_t)c2.size() - 1;
        cut = c2;
        cut_q = q2;
        for (uint64_t c = 0; c < n_cand && rows_on; c++) rows_on = h_flags[c] == 0; // a query outside the kernels' limits
    }
    std::vector<Chunk> chunks(n_chunks);
    DevBuf<uint8_t> d_flags;
    DevBuf<uint32_t> d_cq, d_cn, d_ebegin, d_ne;
    FD_CUDA(ctx, d_cq.alloc(n_cand));
    FD_CUDA(ctx, d_cn.alloc(n_cand));
    FD_CUDA(ctx, d_flags.alloc(n_cand));
    FD_CUDA(ctx, d_ebegin.alloc(n_cand));
    FD_CUDA(ctx, d_ne.alloc(n_cand));
    FD_CUDA(ctx, cudaMemcpyAsync(d_cq.p, cand_query, n_cand * 4, cudaMemcpyHostToDevice, s0));
    FD_CUDA(ctx, cudaMemcpyAsync(d_cn.p, cand_nid, n_cand * 4, cudaMemcpyHostToDevice, s0));
    FD_CUDA(ctx, cudaMemsetAsync(d_flags.p, 0, n_cand, s0));
    FD_CUDA(ctx, cudaMemsetAsync(d_ne.p, 0, n_cand * 4, s0));
    // pinned staging: records (upper bound 2 per candidate, grown on overflow), per-chunk counters, relative firsts
    uint8_t *h_kflags = nullptr;
    unsigned int *h_counters_all = nullptr;
    uint32_t *h_first_rel_all = nullptr;
    FD_TRY(fd_pinned(ctx, 3, n_cand, (void **)&h_kflags));
    FD_TRY(fd_pinned(ctx, 4, 2 * sizeof(unsigned int) * n_chunks, (void **)&h_counters_all));
    FD_TRY(fd_pinned(ctx, 5, (n_cand + n_chunks) * 4, (void **)&h_first_rel_all));
    const FdDeviceStore &S = ctx->store;
    StoreView sv{S.row_offsets, S.n_xyz, S.ca_xyz, S.cb_xyz, S.aa, S.cb_valid, S.aa_rows, S.aa_dir};
    fdg::HashParams hp = fdg::make_params(params->nbin_dist, params->nbin_angle, params->dist_cutoff);
    // the store's pair table answers "which pairs carry this hash" directly -- if it was built with these hash parameters
    const FdPairTable &PT = S.pt;
    const bool use_table = PT.built && PT.nbin_dist == params->nbin_dist && PT.nbin_angle == params->nbin_angle &&
                           PT.dist_cutoff == params->dist_cutoff &&
                           !(getenv("FD_VERIFY_TABLE") && atoi(getenv("FD_VERIFY_TABLE")) == 0);
    const PairTableView ptv{PT.offsets, PT.hash, PT.ij, PT.dir};
    const size_t smem_b = sizeof(WarpState) * VB_WARPS;
    const bool lpt = !(getenv("FD_VERIFY_LPT") && atoi(getenv("FD_VERIFY_LPT")) == 0);
    uint64_t max_chunk = 0;
    FD_CUDA(ctx, cudaFuncSetAttribute(k6b_components, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    for (uint32_t k = 0; k < n_chunks; k++) {
        Chunk &C = chunks[k];
        C.c0 = cut[k];
        C.n = cut[k + 1] - C.c0;
        C.st = (k & 1) ? s1 : s0;
        // pools sized from what the previous calls on this context needed (+50 %), at least 64 edges and 3 components
        // per candidate: a chunk that overflows is re-issued with exact sizes, which doubles its cost and stalls the
        // pipeline behind it (at 4x the database the bench's 2.01 components per candidate overflowed the old 2 x n)
        C.pool_cap = std::max<uint64_t>(1u << 16, (uint64_t)(std::max(64.0, 1.5 * ctx->verify_edges_per_cand) * (double)C.n));
        C.spec_cap = std::max<uint64_t>(1024, (uint64_t)(std::max(3.0, 1.5 * ctx->verify_comps_per_cand) * (double)C.n));
        C.h_counters = h_counters_all + 2 * k;
        C.h_first_rel = h_first_rel_all + C.c0 + k;
        FD_CUDA(ctx, C.ncomp.alloc(C.n + 1));
        FD_CUDA(ctx, C.first.alloc(C.n + 1));
        FD_CUDA(ctx, C.counters.alloc(2));
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, C.ncomp.p, C.first.p, C.n + 1, s0);
        FD_CUDA(ctx, C.tmp.alloc(tb));
        if (lpt) {
            FD_CUDA(ctx, C.order.alloc(C.n + 1));
            FD_CUDA(ctx, C.ne_sorted.alloc(C.n + 1));
            size_t tbs = 0;
            cub::DeviceRadixSort::SortPairsDescending(nullptr, tbs, d_ne.p, C.ne_sorted.p, C.order.p, C.order.p, (int)C.n, 0, 9, s0);
            FD_CUDA(ctx, C.sort_tmp.alloc(tbs));
            max_chunk = std::max<uint64_t>(max_chunk, C.n);
        }
    }
    DevBuf<uint32_t> d_iota;
    if (lpt) {
        FD_CUDA(ctx, d_iota.alloc(max_chunk + 1));
        FD_LAUNCH_ON(ctx, s0, k6_iota, fd_div_up(max_chunk, 256), 256, 0, d_iota.p, (uint32_t)max_chunk);
    }
    // rows plan: the row kernel's inputs and outputs on the device (filled chunk by chunk on the copy stream)
    DevBuf<uint64_t> r_coff, r_roff;
    DevBuf<fd_struct_hit> r_hits;
    DevBuf<uint32_t> r_first;
    DevBuf<fd_struct_row> r_structs, r_structs_tmp;
    DevBuf<fd_match_row> r_matches;
    DevBuf<unsigned long long> r_order;
    DevBuf<fd_residue_row> r_res;
    DevBuf<uint8_t> r_need;
    uint64_t *h_roff = nullptr; // pinned: res_offsets of the plan
    uint64_t res_produced = 0;
    if (rows_on) {
        const uint32_t pq = plan->n_queries;
        FD_TRY(fd_pinned(ctx, 9, (pq + 1) * 8ull, (void **)&h_roff));
        h_roff[0] = 0;
        FD_CUDA(ctx, r_coff.alloc(pq + 1));
        FD_CUDA(ctx, r_roff.alloc(pq + 1));
        FD_CUDA(ctx, r_hits.alloc(n_cand));
        FD_CUDA(ctx, r_first.alloc(n_cand + 1));
        FD_CUDA(ctx, r_structs.alloc(n_cand));
        FD_CUDA(ctx, r_structs_tmp.alloc(n_cand));
        FD_CUDA(ctx, r_matches.alloc(std::max<uint64_t>(plan->match_capacity, 1)));
        FD_CUDA(ctx, r_order.alloc(std::max<uint64_t>(plan->match_capacity, 1)));
        FD_CUDA(ctx, r_res.alloc(std::max<uint64_t>(plan->residue_capacity, 1)));
        FD_CUDA(ctx, r_need.alloc(std::max<uint32_t>(pq, 1)));
        FD_CUDA(ctx, cudaMemcpyAsync(r_coff.p, plan->cand_offsets, (pq + 1) * 8ull, cudaMemcpyHostToDevice, s0));
        FD_CUDA(ctx, cudaMemcpyAsync(r_hits.p, plan->hits, n_cand * sizeof(fd_struct_hit), cudaMemcpyHostToDevice, s0));
    }
    // allocations and uploads were ordered on s0: the other streams start after them
    cudaEvent_t ev_begin = event_at(5 * n_chunks), ev_end = event_at(5 * n_chunks + 1), ev_tmp = event_at(5 * n_chunks + 2);
    FD_CUDA(ctx, cudaEventRecord(ev_begin, s0));
    FD_CUDA(ctx, cudaStreamWaitEvent(s1, ev_begin, 0));
    FD_CUDA(ctx, cudaStreamWaitEvent(sc, ev_begin, 0));
    h_mark("hv_candidates");
    auto issue = [&](uint32_t k) -> int {
        Chunk &C = chunks[k];
        cudaStream_t st = C.st;
        const uint32_t n32 = (uint32_t)C.n;
        // pools are (re)allocated stream-ordered on s0; on the first pass that happened before ev_begin
        FD_CUDA(ctx, C.pool_key.alloc(C.pool_cap));
        FD_CUDA(ctx, C.pool_ent.alloc(C.pool_cap));
        FD_CUDA(ctx, C.specs.alloc(C.spec_cap));
        FD_CUDA(ctx, C.out.alloc(C.spec_cap));
        if (st != s0) {
            FD_CUDA(ctx, cudaEventRecord(ev_tmp, s0));
            FD_CUDA(ctx, cudaStreamWaitEvent(st, ev_tmp, 0));
        }
        FD_CUDA(ctx, cudaEventRecord(event_at(5 * k), st));
        FD_CUDA(ctx, cudaMemsetAsync(C.counters.p, 0, 8, st));
        FD_CUDA(ctx, cudaMemsetAsync(C.ncomp.p, 0, (C.n + 1) * 4, st));
        if (use_table)
            FD_LAUNCH_ON(ctx, st, k6a_table, fd_div_up(C.n, VT_WARPS), VT_WARPS * 32, 0, sv, ptv, P->d_desc, P->d_hash,
                         P->d_aad, d_cq.p + C.c0, d_cn.p + C.c0, n32, hp, ca_dist_cutoff, C.pool_key.p, C.pool_ent.p,
                         C.counters.p, (uint32_t)std::min<uint64_t>(C.pool_cap, 0xffffffffu), d_ebegin.p + C.c0,
                         d_ne.p + C.c0, d_flags.p + C.c0);
        else
            FD_LAUNCH_ON(ctx, st, k6a_edges, n32, VA_THREADS, 0, sv, P->d_desc, P->d_hash, P->d_aad, d_cq.p + C.c0,
                         d_cn.p + C.c0, n32, hp, ca_dist_cutoff, C.pool_key.p, C.pool_ent.p, C.counters.p,
                         (uint32_t)std::min<uint64_t>(C.pool_cap, 0xffffffffu), d_ebegin.p + C.c0, d_ne.p + C.c0,
                         d_flags.p + C.c0);
        FD_CUDA(ctx, cudaEventRecord(event_at(5 * k + 1), st));
        if (lpt) { // edge counts are at most V_MAX_E = 256 (bit 31 = pair-domain flag): nine key bits
            size_t tbs = C.sort_tmp.n;
            FD_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(C.sort_tmp.p, tbs, d_ne.p + C.c0, C.ne_sorted.p, d_iota.p,
                                                                   C.order.p, (int)n32, 0, 9, st));
            ctx->launches += 2;
        }
        FD_LAUNCH_ON(ctx, st, k6b_components, fd_div_up(C.n, VB_WARPS), VB_WARPS * 32, smem_b, sv, P->d_desc, P->d_hash,
                     P->d_aad, P->d_aar, P->d_idx, d_cq.p + C.c0, d_cn.p + C.c0, n32, hp, ca_dist_cutoff, skip_ca_match,
                     C.pool_key.p, C.pool_ent.p, d_ebegin.p + C.c0, d_ne.p + C.c0, C.specs.p, C.counters.p + 1,
                     (uint32_t)std::min<uint64_t>(C.spec_cap, 0xffffffffu), C.ncomp.p, d_flags.p + C.c0,
                     lpt ? C.order.p : (const uint32_t *)nullptr);
        FD_CUDA(ctx, cudaEventRecord(event_at(5 * k + 2), st));
        size_t tb = C.tmp.n;
        FD_CUDA(ctx, cub::DeviceScan::ExclusiveSum(C.tmp.p, tb, C.ncomp.p, C.first.p, C.n + 1, st));
        ctx->launches += 2;
        FD_LAUNCH_ON(ctx, st, k6c_kabsch, fd_div_up(C.spec_cap, 128), 128, 0, sv, P->d_desc, P->d_qca, P->d_qcb,
                     d_cq.p + C.c0, d_cn.p + C.c0, C.specs.p, C.counters.p + 1,
                     (uint32_t)std::min<uint64_t>(C.spec_cap, 0xffffffffu), (uint32_t)C.c0, C.first.p, C.out.p);
        FD_CUDA(ctx, cudaMemcpyAsync(C.h_counters, C.counters.p, 8, cudaMemcpyDeviceToHost, st));
        FD_CUDA(ctx, cudaMemcpyAsync(C.h_first_rel, C.first.p, (C.n + 1) * 4, cudaMemcpyDeviceToHost, st));
        FD_CUDA(ctx, cudaMemcpyAsync(h_kflags + C.c0, d_flags.p + C.c0, C.n, cudaMemcpyDeviceToHost, st));
        FD_CUDA(ctx, cudaEventRecord(event_at(5 * k + 3), st));
        return FD_OK;
    };
    for (uint32_t k = 0; k < n_chunks; k++) FD_TRY(issue(k));
    h_mark("hv_issue");
    fd_match_record *h_out = nullptr; // all records in emission order: pinned host staging, or (keep_device) device memory
    uint64_t h_out_cap = 0, produced = 0;
    {
        uint64_t cap = 0;
        for (auto &C : chunks) cap += C.spec_cap;
        if (keep_device) FD_CUDA(ctx, cudaMallocAsync((void **)&h_out, std::max<uint64_t>(cap, 1) * sizeof(fd_match_record), sc));
        else FD_TRY(fd_pinned(ctx, 0, cap * sizeof(fd_match_record), (void **)&h_out));
        h_out_cap = cap;
    }
    for (uint32_t k = 0; k < n_chunks; k++) {
        Chunk &C = chunks[k];
        for (int attempt = 0;; attempt++) {
            FD_CUDA(ctx, cudaEventSynchronize(event_at(5 * k + 3)));
            const bool pool_over = C.h_counters[0] > C.pool_cap, spec_over = C.h_counters[1] > C.spec_cap;
            if (!pool_over && !spec_over) break;
            if (attempt == 2) return fd_fail(ctx, FD_ERR_STATE, "fd_verify_candidates: pool sizing did not converge");
            if (pool_over) C.pool_cap = C.h_counters[0];
            if (spec_over) C.spec_cap = C.h_counters[1]; // upper bound: the first pass may have stopped counting early
            if (pool_over) C.spec_cap = std::max<uint64_t>(C.spec_cap, 2 * C.n);
            FD_CUDA(ctx, cudaMemsetAsync(d_ne.p + C.c0, 0, C.n * 4, C.st));
            FD_CUDA(ctx, cudaMemsetAsync(d_flags.p + C.c0, 0, C.n, C.st));
            FD_TRY(issue(k));
        }
        C.rec_base = produced;
        const uint64_t np = C.h_counters[1];
        if (produced + np > h_out_cap) { // only after a re-issue with more components than the first estimate
            const uint64_t new_cap = produced + np + (n_cand - C.c0) * 2;
            if (keep_device) {
                fd_match_record *bigger = nullptr;
                FD_CUDA(ctx, cudaMallocAsync((void **)&bigger, new_cap * sizeof(fd_match_record), sc));
                FD_CUDA(ctx, cudaMemcpyAsync(bigger, h_out, produced * sizeof(fd_match_record), cudaMemcpyDeviceToDevice, sc));
                FD_CUDA(ctx, cudaFreeAsync(h_out, sc));
                h_out = bigger;
            } else {
                FD_CUDA(ctx, cudaStreamSynchronize(sc));
                fd_match_record *bigger = nullptr;
                std::vector<fd_match_record> keep(h_out, h_out + produced);
                FD_TRY(fd_pinned(ctx, 0, new_cap * sizeof(fd_match_record), (void **)&bigger));
                memcpy(bigger, keep.data(), produced * sizeof(fd_match_record));
                h_out = bigger;
            }
            h_out_cap = new_cap;
        }
        if (np) {
            // the records are complete (event above); copy them on the copy stream, under the other chunks' kernels
            FD_CUDA(ctx, cudaMemcpyAsync(h_out + produced, C.out.p, np * sizeof(fd_match_record),
                                         keep_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, sc));
        }
        for (uint64_t c = 0; c <= C.n; c++) h_first[C.c0 + c] = (uint32_t)(produced + C.h_first_rel[c]);
        if (rows_on) {
            // rows of the chunk's queries: assembled and copied on the copy stream, under the next chunks' kernels
            const uint32_t qa = cut_q[k], qb = cut_q[k + 1];
            for (uint64_t c = 0; c < C.n && rows_on; c++) rows_on = h_kflags[C.c0 + c] == 0; // general-path candidate
            uint64_t res_end = res_produced;
            for (uint32_t q = qa; q < qb && rows_on; q++) {
                const uint64_t nm_q = h_first[plan->cand_offsets[q + 1]] - h_first[plan->cand_offsets[q]];
                res_end += nm_q * plan->n_res[q];
                h_roff[q + 1] = res_end;
            }
            if (produced + np > plan->match_capacity || res_end > plan->residue_capacity) rows_on = false;
            if (rows_on) {
                FD_CUDA(ctx, cudaMemcpyAsync(r_first.p + C.c0, h_first + C.c0, (C.n + 1) * 4ull, cudaMemcpyHostToDevice, sc));
                FD_CUDA(ctx, cudaMemcpyAsync(r_roff.p + qa, h_roff + qa, (qb - qa + 1) * 8ull, cudaMemcpyHostToDevice, sc));
                if (qb > qa)
                    FD_LAUNCH_ON(ctx, sc, k6d_rows, qb - qa, VD_THREADS, 0, P->d_desc, d_cq.p, r_coff.p, r_hits.p, r_first.p,
                                 h_out, S.row_offsets, S.label_chain, S.label_serial, r_roff.p, r_structs_tmp.p, r_structs.p,
                                 r_matches.p, r_order.p, r_res.p, r_need.p, qa);
                FD_CUDA(ctx, cudaMemcpyAsync(plan->needs_host_sort + qa, r_need.p + qa, qb - qa, cudaMemcpyDeviceToHost, sc));
                FD_CUDA(ctx, cudaMemcpyAsync(plan->structs + C.c0, r_structs.p + C.c0, C.n * sizeof(fd_struct_row), cudaMemcpyDeviceToHost, sc));
                if (np) {
                    FD_CUDA(ctx, cudaMemcpyAsync(plan->matches + produced, r_matches.p + produced, np * sizeof(fd_match_row), cudaMemcpyDeviceToHost, sc));
                    FD_CUDA(ctx, cudaMemcpyAsync(plan->match_order + produced, r_order.p + produced, np * 8ull, cudaMemcpyDeviceToHost, sc));
                }
                if (res_end > res_produced)
                    FD_CUDA(ctx, cudaMemcpyAsync(plan->residues + res_produced, r_res.p + res_produced,
                                                 (res_end - res_produced) * sizeof(fd_residue_row), cudaMemcpyDeviceToHost, sc));
                res_produced = res_end;
            }
        }
        produced += np;
    }
    h_mark("hv_wait_chunks");
    FD_CUDA(ctx, cudaEventRecord(ev_tmp, sc));
    FD_CUDA(ctx, cudaStreamWaitEvent(s0, ev_tmp, 0));
    FD_CUDA(ctx, cudaEventRecord(ev_tmp, s1));
    FD_CUDA(ctx, cudaStreamWaitEvent(s0, ev_tmp, 0));
    FD_CUDA(ctx, cudaEventRecord(ev_end, s0));
    FD_CUDA(ctx, cudaEventSynchronize(ev_end)); // every stream is idle: the DevBufs can be released on s0
    FD_CUDA(ctx, cudaGetLastError());
    h_mark("hv_copy_tail");
    {
        const char *names[3] = {"verify_edges", "verify_components", "verify_kabsch"};
        for (uint32_t k = 0; k < n_chunks; k++)
            for (int j = 0; j < 3; j++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, event_at(5 * k + j), event_at(5 * k + j + 1));
                FdStage &stg = ctx->stages[names[j]];
                stg.ms += ms; // chunks overlap on two streams: these intervals add up to more than the wall time
                stg.launches += k == 0 ? 1 : 0;
            }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev_begin, ev_end);
        FdStage &stg = ctx->stages["verify"];
        stg.ms += ms;
        stg.launches += 1;
    }
    for (uint64_t c = 0; c < n_cand; c++) h_flags[c] |= h_kflags[c];
    {
        uint64_t edges = 0;
        for (auto &C : chunks) edges += C.h_counters[0];
        ctx->verify_edges_per_cand = std::max(ctx->verify_edges_per_cand * 0.9, (double)edges / (double)n_cand);
        ctx->verify_comps_per_cand = std::max(ctx->verify_comps_per_cand * 0.9, (double)produced / (double)n_cand);
    }
    if (plan) {
        plan->done = rows_on ? 1 : 0;
        if (rows_on) memcpy(plan->res_offsets, h_roff, (plan->n_queries + 1) * 8ull);
    }
    if (keep_device) {
        ctx->vkeep.recs = h_out;
        ctx->vkeep.n = produced;
        ctx->vkeep.cap = h_out_cap;
        ctx->vkeep.n_cand = n_cand;
        ctx->vkeep.d_cand_query = d_cq.take();
        ctx->vkeep.prepared = P;
        *out_records = nullptr;
    } else {
        *out_records = h_out;
    }
    *out_n = produced;
    *out_first = h_first;
    *out_flags = h_flags;
    return FD_OK;
}

// one-shot form: prepare the query tables, verify, release
static int verify_core(fd_ctx *ctx, const fd_verify_query *queries, uint32_t nq, const uint32_t *cand_query,
                       const uint32_t *cand_nid, uint64_t n_cand, const fd_hash_params *params, float ca_dist_cutoff,
                       int skip_ca_match, const fd_match_record **out_records, uint64_t *out_n,
                       const uint32_t **out_first, const uint8_t **out_flags) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->store.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_verify_candidates: no structure store attached");
    fd_verify_prepared *P = nullptr;
    FD_TRY(verify_prepare(ctx, queries, nq, &P));
    const int rc = verify_run(ctx, P, cand_query, cand_nid, n_cand, params, ca_dist_cutoff, skip_ca_match, out_records,
                              out_n, out_first, out_flags);
    verify_prepared_release(P);
    return rc;
}

extern "C" int fd_verify_prepare(fd_ctx *ctx, const fd_verify_query *queries, uint32_t nq, fd_verify_prepared **out) {
    return verify_prepare(ctx, queries, nq, out);
}

extern "C" void fd_verify_prepared_free(fd_verify_prepared *p) { verify_prepared_release(p); }

extern "C" uint64_t fd_verify_prepared_bytes(const fd_verify_prepared *p) { return p ? p->h2d_bytes : 0; }

extern "C" int fd_verify_candidates_prepared(fd_ctx *ctx, const fd_verify_prepared *prepared, const uint32_t *cand_query,
                                             const uint32_t *cand_nid, uint64_t n_cand, const fd_hash_params *params,
                                             float ca_dist_cutoff, int skip_ca_match,
                                             const fd_match_record **out_records, uint64_t *out_n,
                                             const uint32_t **out_first, const uint8_t **out_flags) {
    return verify_run(ctx, prepared, cand_query, cand_nid, n_cand, params, ca_dist_cutoff, skip_ca_match, out_records,
                      out_n, out_first, out_flags);
}

extern "C" int fd_verify_candidates_device(fd_ctx *ctx, const fd_verify_prepared *prepared, const uint32_t *cand_query,
                                           const uint32_t *cand_nid, uint64_t n_cand, const fd_hash_params *params,
                                           float ca_dist_cutoff, int skip_ca_match, uint64_t *out_n,
                                           const uint32_t **out_first, const uint8_t **out_flags) {
    const fd_match_record *none = nullptr;
    return verify_run(ctx, prepared, cand_query, cand_nid, n_cand, params, ca_dist_cutoff, skip_ca_match, &none, out_n,
                      out_first, out_flags, true);
}

extern "C" uint64_t fd_verify_match_capacity(const fd_ctx *ctx, uint64_t n_cand) {
    const double per = std::max(3.0, 1.5 * (ctx ? ctx->verify_comps_per_cand : 0.0));
    return (uint64_t)(per * (double)n_cand) + n_cand + 16 * 1024;
}

extern "C" int fd_verify_candidates_rows(fd_ctx *ctx, const fd_verify_prepared *prepared, const uint32_t *cand_query,
                                         const uint32_t *cand_nid, uint64_t n_cand, const fd_hash_params *params,
                                         float ca_dist_cutoff, int skip_ca_match, fd_rows_plan *plan, uint64_t *out_n,
                                         const uint32_t **out_first, const uint8_t **out_flags) {
    if (!plan || !plan->cand_offsets || !plan->n_res || !plan->needs_host_sort || !plan->res_offsets ||
        (n_cand && (!plan->hits || !plan->structs)) || (plan->match_capacity && (!plan->matches || !plan->match_order)) ||
        (plan->residue_capacity && !plan->residues))
        return fd_fail(ctx, FD_ERR_ARG, "fd_verify_candidates_rows: NULL argument");
    plan->done = 0;
    const fd_match_record *none = nullptr;
    return verify_run(ctx, prepared, cand_query, cand_nid, n_cand, params, ca_dist_cutoff, skip_ca_match, &none, out_n,
                      out_first, out_flags, true, n_cand ? plan : nullptr);
}

extern "C" int fd_verify_records_fetch(fd_ctx *ctx, const fd_match_record **out_records) {
    if (!ctx || !out_records) return FD_ERR_ARG;
    FD_ENTER(ctx);
    const FdVerifyKeep &K = ctx->vkeep;
    fd_match_record *h = nullptr;
    FD_TRY(fd_pinned(ctx, 0, std::max<uint64_t>(K.n, 1) * sizeof(fd_match_record), (void **)&h));
    if (K.n) {
        if (!K.recs) return fd_fail(ctx, FD_ERR_STATE, "fd_verify_records_fetch: no records kept on the device");
        FD_CUDA(ctx, cudaMemcpyAsync(h, K.recs, K.n * sizeof(fd_match_record), cudaMemcpyDeviceToHost, ctx->stream));
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *out_records = h;
    return FD_OK;
}

extern "C" int fd_verify_rows(fd_ctx *ctx, const fd_rows_request *rq) {
    if (!ctx || !rq) return FD_ERR_ARG;
    FD_ENTER(ctx);
    const FdVerifyKeep &K = ctx->vkeep;
    if (!K.prepared || (K.n && !K.recs)) return fd_fail(ctx, FD_ERR_STATE, "fd_verify_rows: no records kept on the device (call fd_verify_candidates_device first)");
    if (!rq->cand_offsets || !rq->res_offsets || !rq->needs_host_sort || (K.n_cand && (!rq->hits || !rq->structs)) ||
        (K.n && (!rq->matches || !rq->match_order)))
        return fd_fail(ctx, FD_ERR_ARG, "fd_verify_rows: NULL argument");
    const uint32_t nq = rq->n_queries;
    if (rq->cand_offsets[0] != 0 || rq->cand_offsets[nq] != K.n_cand)
        return fd_fail(ctx, FD_ERR_ARG, "fd_verify_rows: cand_offsets must cover the verified candidates");
    const fd_verify_prepared *P = (const fd_verify_prepared *)K.prepared;
    cudaStream_t s = ctx->stream;
    const FdDeviceStore &S = ctx->store;
    const uint64_t n_cand = K.n_cand, n_rec = K.n, n_res = rq->res_offsets[nq];
    const uint32_t *h_first = (const uint32_t *)ctx->pinned[2].p; // the first[] of the verification call (pinned)
    DevBuf<uint64_t> d_coff, d_roff;
    DevBuf<fd_struct_hit> d_hits;
    DevBuf<uint32_t> d_first;
    DevBuf<fd_struct_row> d_structs, d_structs_tmp;
    DevBuf<fd_match_row> d_matches;
    DevBuf<unsigned long long> d_order;
    DevBuf<fd_residue_row> d_res;
    DevBuf<uint8_t> d_need;
    FD_CUDA(ctx, d_coff.alloc(nq + 1));
    FD_CUDA(ctx, d_roff.alloc(nq + 1));
    FD_CUDA(ctx, d_hits.alloc(std::max<uint64_t>(n_cand, 1)));
    FD_CUDA(ctx, d_first.alloc(n_cand + 1));
    FD_CUDA(ctx, d_structs.alloc(std::max<uint64_t>(n_cand, 1)));
    FD_CUDA(ctx, d_structs_tmp.alloc(std::max<uint64_t>(n_cand, 1)));
    FD_CUDA(ctx, d_matches.alloc(std::max<uint64_t>(n_rec, 1)));
    FD_CUDA(ctx, d_order.alloc(std::max<uint64_t>(n_rec, 1)));
    FD_CUDA(ctx, d_res.alloc(std::max<uint64_t>(n_res, 1)));
    FD_CUDA(ctx, d_need.alloc(std::max<uint32_t>(nq, 1)));
    StageTimer st(ctx, "rows");
    FD_CUDA(ctx, cudaMemcpyAsync(d_coff.p, rq->cand_offsets, (nq + 1) * 8ull, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_roff.p, rq->res_offsets, (nq + 1) * 8ull, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_hits.p, rq->hits, n_cand * sizeof(fd_struct_hit), cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_first.p, h_first, (n_cand + 1) * 4ull, cudaMemcpyHostToDevice, s));
    if (nq)
        FD_LAUNCH(ctx, k6d_rows, nq, VD_THREADS, 0, P->d_desc, K.d_cand_query, d_coff.p, d_hits.p, d_first.p,
                  (const fd_match_record *)K.recs, S.row_offsets, S.label_chain, S.label_serial, d_roff.p, d_structs_tmp.p,
                  d_structs.p, d_matches.p, d_order.p, d_res.p, d_need.p, 0u);
    FD_CUDA(ctx, cudaMemcpyAsync(rq->needs_host_sort, d_need.p, nq, cudaMemcpyDeviceToHost, s));
    if (n_cand) FD_CUDA(ctx, cudaMemcpyAsync(rq->structs, d_structs.p, n_cand * sizeof(fd_struct_row), cudaMemcpyDeviceToHost, s));
    if (n_rec) {
        FD_CUDA(ctx, cudaMemcpyAsync(rq->matches, d_matches.p, n_rec * sizeof(fd_match_row), cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, cudaMemcpyAsync(rq->match_order, d_order.p, n_rec * 8ull, cudaMemcpyDeviceToHost, s));
    }
    if (n_res) FD_CUDA(ctx, cudaMemcpyAsync(rq->residues, d_res.p, n_res * sizeof(fd_residue_row), cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, st.finish());
    verify_keep_release(ctx);
    return FD_OK;
}

extern "C" int fd_store_has_labels(const fd_ctx *ctx) { return ctx && ctx->store.label_chain ? 1 : 0; }

extern "C" int fd_store_attach_labels(fd_ctx *ctx, const uint8_t *chain, const uint64_t *serial, uint64_t n_residues) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->store.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_store_attach_labels: no structure store attached");
    if (ctx->borrowed) return fd_fail(ctx, FD_ERR_STATE, "fd_store_attach_labels: a forked context shares its parent's store");
    if (!chain || !serial || n_residues != ctx->store.n_res)
        return fd_fail(ctx, FD_ERR_ARG, "fd_store_attach_labels: one label per residue of the attached store");
    FD_ENTER(ctx);
    FdDeviceStore &S = ctx->store;
    cudaFree(S.label_chain);
    cudaFree(S.label_serial);
    S.label_chain = nullptr;
    S.label_serial = nullptr;
    FD_CUDA(ctx, cudaMalloc((void **)&S.label_chain, std::max<uint64_t>(n_residues, 1)));
    FD_CUDA(ctx, cudaMalloc((void **)&S.label_serial, std::max<uint64_t>(n_residues, 1) * 8));
    FD_TRY(fd_copy_to_device_staged(ctx, S.label_chain, chain, n_residues));
    FD_TRY(fd_copy_to_device_staged(ctx, S.label_serial, serial, n_residues * 8));
    return FD_OK;
}

extern "C" int fd_verify_candidates_view(fd_ctx *ctx, const fd_verify_query *queries, uint32_t nq,
                                         const uint32_t *cand_query, const uint32_t *cand_nid, uint64_t n_cand,
                                         const fd_hash_params *params, float ca_dist_cutoff, int skip_ca_match,
                                         const fd_match_record **out_records, uint64_t *out_n,
                                         const uint32_t **out_first, const uint8_t **out_flags) {
    return verify_core(ctx, queries, nq, cand_query, cand_nid, n_cand, params, ca_dist_cutoff, skip_ca_match,
                       out_records, out_n, out_first, out_flags);
}

extern "C" int fd_verify_candidates_batch(fd_ctx *ctx, const fd_verify_query *queries, uint32_t nq,
                                          const uint32_t *cand_query, const uint32_t *cand_nid, uint64_t n_cand,
                                          const fd_hash_params *params, float ca_dist_cutoff, int skip_ca_match,
                                          fd_match_record **out_records, uint64_t *out_n, uint8_t **out_flags) {
    if (!out_records || !out_n || !out_flags) return fd_fail(ctx, FD_ERR_ARG, "fd_verify_candidates_batch: NULL argument");
    const fd_match_record *recs = nullptr;
    const uint32_t *first = nullptr;
    const uint8_t *flags = nullptr;
    uint64_t n = 0;
    *out_records = nullptr;
    *out_flags = nullptr;
    *out_n = 0;
    FD_TRY(verify_core(ctx, queries, nq, cand_query, cand_nid, n_cand, params, ca_dist_cutoff, skip_ca_match, &recs, &n,
                       &first, &flags));
    fd_match_record *r = (fd_match_record *)malloc(std::max<uint64_t>(n, 1) * sizeof(fd_match_record));
    uint8_t *f = (uint8_t *)malloc(std::max<uint64_t>(n_cand, 1));
    if (!r || !f) {
        free(r);
        free(f);
        return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
    }
    memcpy(r, recs, n * sizeof(fd_match_record));
    memcpy(f, flags, n_cand);
    *out_records = r;
    *out_n = n;
    *out_flags = f;
    return FD_OK;
}
