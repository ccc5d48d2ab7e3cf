// fd_verify.cu -- K6: fused candidate verification (query path (ii), everything after the prefilter).
//
// Replaces retrieval_wrapper (reference src/controller/retrieve.rs:364-552) for a batch of
// (query, candidate structure) pairs in ONE kernel, so that nothing but the final match records leaves HBM:
//   retrieve_with_prefilter   (:52-156)    re-hash of the candidate's residue pairs from the HBM structure store
//   create_index_graph + connected_components_with_given_node_count (graph.rs:16-50)
//   map_query_and_retrieved_residues (:604-702), calculate_subgraph_idf (:705-719), the rescue loop (:453-516)
//   rmsd_with_calpha_and_rottran -> kabsch (:756-834, kabsch.rs:157-554)
// The un-fused pieces (fd_candidate_edges_batch + host graph step + fd_kabsch_store_batch) remain as the general
// path; this kernel handles candidates that fit its shared-memory limits and flags the rest (V_OVERFLOW).
//
// One CTA (128 threads) per candidate.  Parallel phases (pair screen, hashing, rescue voting, Kabsch per
// component) use the whole CTA; the small irregular graph phase runs on thread 0 over shared-memory state:
// graphs have <= 64 nodes, so reachability closures, components and "used" sets are single 64-bit masks.
#include <cub/cub.cuh>

#include <algorithm>

#include "fd_common.cuh"
#include "fd_geom.cuh"
#include "fd_kabsch.cuh"

namespace {

constexpr int V_THREADS = 128;
constexpr int V_LIST_CAP = 4096;
constexpr int V_CHUNK = V_THREADS * 8;
constexpr int V_MAX_AAD = 255; // start and count of an amino-acid pair's run each fit 8 bits
constexpr int V_MAX_E = 256;
constexpr int V_MAX_NODES = 64;
constexpr int V_MAX_C = 16;
constexpr int V_MAX_NQ = 16;
constexpr uint32_t V_PREFILTER_SKIP = 200; // retrieve.rs:24

struct StoreView {
    const uint64_t *row_offsets;
    const float *n_xyz, *ca_xyz, *cb_xyz;
    const uint8_t *aa, *cb_valid;
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

__global__ void __launch_bounds__(V_THREADS)
    k6_verify(StoreView st, const VQDesc *vq, const VHash *vhash, const VAad *vaad, const uint8_t *idx_dense,
              const float *q_ca, const float *q_cb, const uint32_t *cand_query,
              const uint32_t *cand_nid, uint32_t n_cand, fdg::HashParams hp, float ca_cutoff, int skip_ca_match,
              fd_match_record *out, unsigned int *out_count, uint32_t out_cap, uint8_t *cand_flags) {
    __shared__ uint16_t list1[V_LIST_CAP], list2[V_LIST_CAP];
    __shared__ uint32_t n1, n2, q_n, n_e;
    __shared__ uint32_t q_ij[V_CHUNK];
    __shared__ float q_d[V_CHUNK];
    __shared__ VAad aad[V_MAX_AAD];        // sorted by (aa1, aa2) on the host
    __shared__ uint16_t aa_range[400];      // aa1 * 20 + aa2 -> start << 8 | count of its entries (0 = pair not in query)
    __shared__ uint32_t e_key[V_MAX_E]; // i << 16 | j, sorted
    __shared__ uint16_t e_ent[V_MAX_E]; // index of the edge's hash inside the query's sorted hash set
    __shared__ uint8_t e_a[V_MAX_E], e_b[V_MAX_E];
    __shared__ uint16_t node_res[V_MAX_NODES];
    __shared__ uint64_t comp_mask[V_MAX_C];
    __shared__ uint64_t reach[V_MAX_NODES], und[V_MAX_NODES]; // directed / undirected reachability closures
    __shared__ uint32_t n_nodes, n_comp, s_flag;
    __shared__ uint8_t counts[V_MAX_NQ * V_MAX_NODES];
    // per-component results
    __shared__ float c_idf[V_MAX_C];
    __shared__ uint32_t c_res[V_MAX_C][V_MAX_NQ]; // target residue index + 1 per query position
    __shared__ uint32_t c_aq[V_MAX_C][V_MAX_NQ], c_at[V_MAX_C][V_MAX_NQ];
    __shared__ uint32_t c_nal[V_MAX_C], c_nodes[V_MAX_C];
    // rescue scratch
    __shared__ uint32_t r_dq, r_max, r_nmax, r_arg, r_need, r_nridx;
    __shared__ uint16_t r_ridx[V_MAX_NQ];
    __shared__ unsigned int s_out_base;
    __shared__ uint8_t m_qidx[V_MAX_NQ], m_ridx[V_MAX_NQ]; // mapping of the current component: dense query id, node id
    __shared__ uint32_t m_n;
    __shared__ uint32_t f_qscan[V_MAX_NQ], f_rscan[V_MAX_NQ], f_nscan, f_res[V_MAX_NQ], f_nres;

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t c = blockIdx.x;
    if (c >= n_cand) return;
    const VQDesc Q = vq[cand_query[c]];
    const uint32_t t = cand_nid[c];
    const uint64_t base = st.row_offsets[t];
    const uint32_t n = (uint32_t)(st.row_offsets[t + 1] - base);
    const VHash *H = vhash + Q.hash_begin;
    if (tid == 0) {
        n1 = n2 = q_n = n_e = 0;
        s_flag = 0;
        n_comp = 0;
    }
    for (uint32_t k = tid; k < Q.n_aad; k += V_THREADS) aad[k] = vaad[Q.aad_begin + k];
    for (uint32_t k = tid; k < 400; k += V_THREADS) aa_range[k] = 0;
    __syncthreads();
    if (Q.n_hashes == 0 || Q.n_aad == 0) return;
    for (uint32_t k = tid; k < Q.n_aad; k += V_THREADS) { // heads of runs of equal (aa1, aa2)
        const uint32_t key = aad[k].aa1 * 20u + aad[k].aa2;
        if (k == 0 || aad[k - 1].aa1 * 20u + aad[k - 1].aa2 != key) {
            uint32_t e = k + 1;
            while (e < Q.n_aad && aad[e].aa1 * 20u + aad[e].aa2 == key) e++;
            aa_range[key] = (uint16_t)((k << 8) | (e - k));
        }
    }
    __syncthreads();

    // ---- prefilter sets (prefilter_amino_acid, retrieve.rs:563-602) ----
    bool all_pairs = !Q.use_prefilter;
    if (!all_pairs) {
        for (uint32_t r0 = 0; r0 < n; r0 += V_THREADS) {
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
            if (tid == 0) cand_flags[c] = 1; // V_OVERFLOW: handled by the general path
            return;
        }
    }
    const uint32_t rows = all_pairs ? n : n1, cols = all_pairs ? n : n2;
    const uint64_t total = (uint64_t)rows * cols; // < 2^32: both factors are at most 65535

    // ---- retrieve_with_prefilter: screen, hash, keep pairs whose hash is in the query set ----
    for (uint64_t p0 = 0; p0 < total; p0 += V_CHUNK) {
        for (uint32_t u = 0; u < V_CHUNK / V_THREADS; u++) {
            const uint64_t p = p0 + (uint64_t)u * V_THREADS + tid;
            bool pass = false;
            uint32_t i = 0, j = 0;
            float d = 0.f;
            if (p < total) {
                const uint32_t p32 = (uint32_t)p;
                const uint32_t a = p32 / cols, b = p32 - a * cols;
                i = all_pairs ? a : list1[a];
                j = all_pairs ? b : list2[b];
                const uint8_t ai = st.aa[base + i], aj = st.aa[base + j];
                if (i != j && ai != 255 && aj != 255 && aa_range[(ai & 0x7Fu) * 20u + (aj & 0x7Fu)] != 0) {
                    d = fdg::dist(ld3(st.ca_xyz, base + i), ld3(st.ca_xyz, base + j));
                    if (d <= hp.dist_cutoff) {
                        const uint32_t rg = aa_range[(ai & 0x7Fu) * 20u + (aj & 0x7Fu)];
                        for (uint32_t k = rg >> 8, ke = (rg >> 8) + (rg & 0xffu); k < ke; k++)
                            if (fabsf(d - aad[k].dist) < ca_cutoff) {
                                pass = true;
                                break;
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
        const uint32_t qn = q_n;
        for (uint32_t k = tid; k < qn; k += V_THREADS) {
            const uint32_t i = q_ij[k] >> 16, j = q_ij[k] & 0xffffu;
            const uint64_t ri = base + i, rj = base + j;
            const bool cbok = st.cb_valid == nullptr || (st.cb_valid[ri] && st.cb_valid[rj]);
            if (!cbok) continue;
            const uint32_t h = fdg::pair_hash(ld3(st.n_xyz, ri), ld3(st.ca_xyz, ri), ld3(st.cb_xyz, ri), ld3(st.n_xyz, rj),
                                              ld3(st.ca_xyz, rj), ld3(st.cb_xyz, rj), st.aa[ri] & 0x7Fu,
                                              st.aa[rj] & 0x7Fu, q_d[k], hp);
            uint32_t lo = 0, hi = Q.n_hashes;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (H[mid].hash < h) lo = mid + 1;
                else hi = mid;
            }
            if (lo < Q.n_hashes && H[lo].hash == h) {
                const uint32_t pos = atomicAdd(&n_e, 1u);
                if (pos < V_MAX_E) {
                    e_key[pos] = q_ij[k];
                    e_ent[pos] = (uint16_t)lo;
                }
            }
        }
        __syncthreads();
        if (tid == 0) q_n = 0;
        __syncthreads();
    }
    const uint32_t ne = n_e;
    if (ne == 0) return;
    if (ne > V_MAX_E || Q.n_idx > V_MAX_NQ || Q.n_dq > V_MAX_NQ) {
        if (tid == 0) cand_flags[c] = 1;
        return;
    }
    // ---- sort edges by (i, j): the reference's emission order (graph node numbering, f32 sum order) ----
    uint32_t sort_n = 2;
    while (sort_n < ne) sort_n <<= 1;
    for (uint32_t k = ne + tid; k < sort_n; k += V_THREADS) e_key[k] = 0xffffffffu;
    __syncthreads();
    for (uint32_t size = 2; size <= sort_n; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t k = tid; k < sort_n / 2; k += V_THREADS) {
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

    // ---- graph: nodes by first appearance, components = SCCs U weak components, size >= 2 ----
    if (tid == 0) {
        uint32_t nn = 0;
        bool overflow = false;
        for (uint32_t k = 0; k < ne && !overflow; k++) {
            const uint16_t ends[2] = {(uint16_t)(e_key[k] >> 16), (uint16_t)(e_key[k] & 0xffffu)};
            uint8_t ids[2];
            for (int s = 0; s < 2; s++) {
                uint32_t f = 0;
                while (f < nn && node_res[f] != ends[s]) f++;
                if (f == nn) {
                    if (nn == V_MAX_NODES) {
                        overflow = true;
                        break;
                    }
                    node_res[nn++] = ends[s];
                }
                ids[s] = (uint8_t)f;
            }
            e_a[k] = ids[0];
            e_b[k] = ids[1];
        }
        if (overflow) {
            s_flag = 1;
        } else {
            n_nodes = nn;
            for (uint32_t v = 0; v < nn; v++) reach[v] = und[v] = 1ull << v;
            bool changed = true;
            while (changed) {
                changed = false;
                for (uint32_t k = 0; k < ne; k++) {
                    const uint32_t a = e_a[k], b = e_b[k];
                    const uint64_t ra = reach[a] | reach[b];
                    if (ra != reach[a]) {
                        reach[a] = ra;
                        changed = true;
                    }
                    const uint64_t u = und[a] | und[b];
                    if (u != und[a] || u != und[b]) {
                        und[a] = und[b] = u;
                        changed = true;
                    }
                }
            }
            uint32_t nc = 0;
            auto add_mask = [&](uint64_t m) {
                if (__popcll(m) < 2) return;
                for (uint32_t k = 0; k < nc; k++)
                    if (comp_mask[k] == m) return;
                if (nc == V_MAX_C) {
                    s_flag = 1;
                    return;
                }
                comp_mask[nc++] = m;
            };
            for (uint32_t v = 0; v < nn; v++) {
                uint64_t scc = 0;
                for (uint64_t r = reach[v]; r; r &= r - 1) {
                    const int w = ctz64(r);
                    if ((reach[w] >> v) & 1ull) scc |= 1ull << w;
                }
                add_mask(scc);
                add_mask(und[v]);
            }
            for (uint32_t a = 1; a < nc; a++) { // insertion sort, lexicographic by sorted node list
                const uint64_t m = comp_mask[a];
                uint32_t b = a;
                while (b > 0 && mask_less(m, comp_mask[b - 1])) {
                    comp_mask[b] = comp_mask[b - 1];
                    b--;
                }
                comp_mask[b] = m;
            }
            n_comp = nc;
        }
    }
    __syncthreads();
    if (s_flag) {
        if (tid == 0) cand_flags[c] = 1;
        return;
    }
    const uint32_t ncomp = n_comp;
    const uint8_t *IDX = idx_dense + Q.idx_begin;

    for (uint32_t ci = 0; ci < ncomp; ci++) {
        // ---- mapping (thread 0): votes, best per query residue, greedy assignment ----
        const uint64_t mask = comp_mask[ci];
        for (uint32_t k = tid; k < Q.n_dq * V_MAX_NODES; k += V_THREADS) counts[k] = 0;
        __syncthreads();
        if (tid == 0) {
            const uint32_t nq_d = Q.n_dq;
            uint8_t best_c[V_MAX_NQ], best_r[V_MAX_NQ];
            for (uint32_t q = 0; q < nq_d; q++) {
                best_c[q] = 0;
                best_r[q] = 0xff;
            }
            float idf = 0.f;
            for (uint32_t k = 0; k < ne; k++) {
                const uint32_t a = e_a[k], b = e_b[k];
                if (!((mask >> a) & 1ull) || !((mask >> b) & 1ull)) continue;
                const VHash h = H[e_ent[k]];
                idf += h.idf;
                uint32_t pq[2], pr[2];
                if (h.sym) {
                    pq[0] = min((uint32_t)h.dqi, (uint32_t)h.dqj);
                    pq[1] = max((uint32_t)h.dqi, (uint32_t)h.dqj);
                    const bool ab = node_res[a] < node_res[b];
                    pr[0] = ab ? a : b;
                    pr[1] = ab ? b : a;
                } else {
                    pq[0] = h.dqi;
                    pq[1] = h.dqj;
                    pr[0] = a;
                    pr[1] = b;
                }
                for (int s = 0; s < 2; s++) {
                    uint8_t &cnt = counts[pq[s] * V_MAX_NODES + pr[s]];
                    if (cnt != 255) cnt++;
                    const uint32_t q = pq[s];
                    // best = (count, target residue) starts at (0, 0): the tie rule `r < best.1` only ever compares
                    // against a residue that already has a vote (retrieve.rs:655-660)
                    if (cnt > best_c[q] || (cnt == best_c[q] && node_res[pr[s]] < node_res[best_r[q]])) {
                        best_c[q] = cnt;
                        best_r[q] = (uint8_t)pr[s];
                    }
                }
            }
            c_idf[ci] = idf;
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
                    m_qidx[nm] = (uint8_t)q;
                    m_ridx[nm] = (uint8_t)r;
                    r_used |= 1ull << r;
                    nm++;
                }
            }
            m_n = nm;
            c_nodes[ci] = node_count;
            f_nscan = 0;
            f_nres = 0;
            r_nridx = nm;
            for (uint32_t k = 0; k < nm; k++) r_ridx[k] = node_res[m_ridx[k]];
        }
        __syncthreads();
        // ---- rescue loop over the query residues (retrieve.rs:453-516) ----
        for (uint32_t pos = 0; pos < Q.n_idx; pos++) {
            if (tid == 0) {
                const uint32_t dq = IDX[pos];
                int mapped = -1;
                for (uint32_t k = 0; k < m_n; k++)
                    if (m_qidx[k] == dq) mapped = (int)node_res[m_ridx[k]];
                r_need = 0;
                if (mapped >= 0) {
                    const uint32_t ri = (uint32_t)mapped;
                    c_res[ci][pos] = ri + 1; // res_vec_from_hash
                    uint32_t pp = 0;
                    while (pp < f_nscan && f_rscan[pp] != ri) pp++;
                    if (pp == f_nscan) {
                        f_res[f_nres++] = ri + 1;
                        f_qscan[f_nscan] = dq;
                        f_rscan[f_nscan++] = ri;
                    } else {
                        f_res[pp] = 0; // sic (retrieve.rs:475)
                        f_res[f_nres++] = ri + 1;
                        for (uint32_t k = pp; k + 1 < f_nscan; k++) {
                            f_qscan[k] = f_qscan[k + 1];
                            f_rscan[k] = f_rscan[k + 1];
                        }
                        f_qscan[f_nscan - 1] = dq;
                        f_rscan[f_nscan - 1] = ri;
                    }
                } else {
                    c_res[ci][pos] = 0;
                    r_need = 1;
                    r_dq = dq;
                    r_max = 0;
                    r_nmax = 0;
                    r_arg = 0;
                }
            }
            __syncthreads();
            if (r_need) {
                // count_map[i] = #(entries of this query residue, matched target residues rj) compatible with (i, rj)
                const uint32_t dq = r_dq;
                const uint32_t nrows = all_pairs ? n : n1;
                auto row_count = [&](uint32_t a) -> uint32_t {
                    const uint32_t i = all_pairs ? a : list1[a];
                    const uint8_t ai = st.aa[base + i];
                    if (ai == 255 || (st.cb_valid != nullptr && !st.cb_valid[base + i])) return 0u;
                    const uint8_t cia = ai & 0x7Fu;
                    const fdg::V3 cai = ld3(st.ca_xyz, base + i);
                    uint32_t cnt = 0;
                    for (uint32_t k = 0; k < r_nridx; k++) {
                        const uint32_t rj = r_ridx[k];
                        if (rj == i) continue;
                        const uint8_t aj = st.aa[base + rj];
                        if (aj == 255 || (st.cb_valid != nullptr && !st.cb_valid[base + rj])) continue;
                        if (!all_pairs && !(((aj & 0x80u) == 0) && ((Q.aa2_mask >> (aj & 31u)) & 1u))) continue;
                        const float d = fdg::dist(cai, ld3(st.ca_xyz, base + rj));
                        if (!(d <= hp.dist_cutoff)) continue;
                        const uint32_t rg = aa_range[cia * 20u + (aj & 0x7Fu)];
                        for (uint32_t e = rg >> 8, ee = (rg >> 8) + (rg & 0xffu); e < ee; e++)
                            if (aad[e].dq == dq && fabsf(d - aad[e].dist) < ca_cutoff) cnt++;
                    }
                    return cnt;
                };
                for (uint32_t a = tid; a < nrows; a += V_THREADS) { // pass 1: the maximum count
                    const uint32_t cnt = row_count(a);
                    if (cnt) atomicMax(&r_max, cnt);
                }
                __syncthreads();
                const uint32_t mx = r_max;
                if (mx > 0)
                    for (uint32_t a = tid; a < nrows; a += V_THREADS) // pass 2: who reaches it
                        if (row_count(a) == mx) {
                            atomicAdd(&r_nmax, 1u);
                            r_arg = all_pairs ? a : list1[a];
                        }
                __syncthreads();
                if (tid == 0) {
                    bool ok = r_max >= 2 && r_nmax == 1;
                    if (ok)
                        for (uint32_t k = 0; k < f_nscan; k++)
                            if (f_rscan[k] == r_arg) ok = false;
                    if (ok) {
                        f_res[f_nres++] = r_arg + 1;
                        f_qscan[f_nscan] = r_dq;
                        f_rscan[f_nscan++] = r_arg;
                    } else {
                        f_res[f_nres++] = 0;
                    }
                }
            }
            __syncthreads();
        }
        // ---- choose the alignment and the reported residues ----
        if (tid == 0) {
            bool same = true;
            for (uint32_t k = 0; k < Q.n_idx; k++) same = same && f_res[k] == c_res[ci][k];
            uint32_t nal;
            if (skip_ca_match || same) {
                nal = m_n;
                for (uint32_t k = 0; k < nal; k++) {
                    c_aq[ci][k] = m_qidx[k];
                    c_at[ci][k] = node_res[m_ridx[k]];
                }
            } else {
                nal = f_nscan;
                for (uint32_t k = 0; k < nal; k++) {
                    c_aq[ci][k] = f_qscan[k];
                    c_at[ci][k] = f_rscan[k];
                }
            }
            c_nal[ci] = nal;
            if (!skip_ca_match)
                for (uint32_t k = 0; k < Q.n_idx; k++) c_res[ci][k] = f_res[k];
        }
        __syncthreads();
    }

    // ---- Kabsch per component (thread ci), append match records ----
    if (tid == 0) {
        const unsigned int b = atomicAdd(out_count, ncomp);
        s_out_base = b;
    }
    __syncthreads();
    if ((uint32_t)tid < ncomp) {
        const uint32_t ci = tid;
        const unsigned int slot = s_out_base + ci;
        if (slot < out_cap) {
            fd_match_record rec;
            rec.cand = c;
            rec.idf = c_idf[ci];
            uint32_t nodes = 0;
            for (uint32_t k = 0; k < V_MAX_NQ; k++) {
                rec.res[k] = k < Q.n_idx ? c_res[ci][k] : 0;
                nodes += rec.res[k] != 0;
            }
            rec.node_count = nodes;
            fdk::GatherPoints mov{st.ca_xyz, st.cb_xyz, c_at[ci], base};
            fdk::GatherPoints ref{q_ca, q_cb, c_aq[ci], (uint64_t)Q.qres_base};
            fdk::kabsch_one(mov, ref, 2 * c_nal[ci], rec.U, rec.t, &rec.rmsd);
            out[slot] = rec;
        }
    }
}

} // namespace

extern "C" int fd_verify_candidates_batch(fd_ctx *ctx, const fd_verify_query *queries, uint32_t nq,
                                          const uint32_t *cand_query, const uint32_t *cand_nid, uint64_t n_cand,
                                          const fd_hash_params *params, float ca_dist_cutoff, int skip_ca_match,
                                          fd_match_record **out_records, uint64_t *out_n, uint8_t **out_flags) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->store.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_verify_candidates_batch: no structure store attached");
    if ((nq && !queries) || (n_cand && (!cand_query || !cand_nid)) || !params || !out_records || !out_n || !out_flags)
        return fd_fail(ctx, FD_ERR_ARG, "fd_verify_candidates_batch: NULL argument");
    if (n_cand > 0xfffffff0ull) return fd_fail(ctx, FD_ERR_LIMIT, "too many candidates in one call");
    FD_ENTER(ctx);
    *out_records = nullptr;
    *out_flags = nullptr;
    *out_n = 0;
    uint8_t *h_flags = (uint8_t *)calloc(std::max<uint64_t>(n_cand, 1), 1);
    if (!h_flags) return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
    // flatten queries; a query outside the kernel's limits marks all of its candidates for the general path
    std::vector<VQDesc> descs(nq);
    std::vector<VHash> f_hash;
    std::vector<VAad> f_aad;
    std::vector<uint8_t> f_idx;
    std::vector<float> q_ca, q_cb;
    std::vector<uint8_t> q_unfit(nq, 0);
    for (uint32_t q = 0; q < nq; q++) {
        const fd_verify_query &Q = queries[q];
        VQDesc d{};
        d.hash_begin = (uint32_t)f_hash.size();
        d.n_hashes = Q.n_hashes;
        d.aad_begin = (uint32_t)f_aad.size();
        d.n_aad = Q.n_aa_dist;
        d.use_prefilter = Q.n_hashes <= V_PREFILTER_SKIP ? 1u : 0u;
        d.idx_begin = (uint32_t)f_idx.size();
        d.n_idx = Q.n_indices;
        d.qres_base = (uint32_t)(q_ca.size() / 3);
        // dense ids over every query residue index that occurs (ascending residue index)
        std::vector<uint32_t> dq;
        for (uint32_t k = 0; k < Q.n_hashes; k++) {
            dq.push_back(Q.hash_qi[k]);
            dq.push_back(Q.hash_qj[k]);
        }
        for (uint32_t k = 0; k < Q.n_indices; k++) dq.push_back(Q.indices[k]);
        for (uint32_t k = 0; k < Q.n_aa_dist; k++) dq.push_back(Q.q_index[k]);
        std::sort(dq.begin(), dq.end());
        dq.erase(std::unique(dq.begin(), dq.end()), dq.end());
        d.n_dq = (uint32_t)dq.size();
        if (dq.size() > V_MAX_NQ || Q.n_indices > V_MAX_NQ || Q.n_aa_dist > V_MAX_AAD || Q.n_hashes > 65535) q_unfit[q] = 1;
        auto dense = [&](uint32_t r) { return (uint8_t)(std::lower_bound(dq.begin(), dq.end(), r) - dq.begin()); };
        for (uint32_t k = 0; k < Q.n_hashes; k++) {
            if (k && Q.hashes_sorted[k] <= Q.hashes_sorted[k - 1]) {
                free(h_flags);
                return fd_fail(ctx, FD_ERR_ARG, "fd_verify_query: hashes_sorted must be strictly ascending");
            }
            f_hash.push_back(VHash{Q.hashes_sorted[k], Q.hash_idf[k], dense(Q.hash_qi[k]), dense(Q.hash_qj[k]),
                                   Q.hash_symmetric[k], 0});
            d.aa1_mask |= 1u << ((Q.hashes_sorted[k] >> 25) & 31u);
            d.aa2_mask |= 1u << ((Q.hashes_sorted[k] >> 20) & 31u);
        }
        if (!q_unfit[q]) { // entries grouped by amino-acid pair (the kernel indexes them through a 20x20 table)
            std::vector<uint32_t> ord(Q.n_aa_dist);
            for (uint32_t k = 0; k < Q.n_aa_dist; k++) ord[k] = k;
            std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) {
                return Q.aa1[a] * 20u + Q.aa2[a] < Q.aa1[b] * 20u + Q.aa2[b];
            });
            for (uint32_t k : ord) {
                if (Q.aa1[k] >= 20 || Q.aa2[k] >= 20) {
                    free(h_flags);
                    return fd_fail(ctx, FD_ERR_ARG, "fd_verify_query: amino-acid code out of range");
                }
                f_aad.push_back(VAad{Q.aa1[k], Q.aa2[k], dense(Q.q_index[k]), 0, Q.ca_dist[k]});
            }
        }
        if (q_unfit[q]) d.n_aad = 0; // kernel returns immediately for this query's candidates
        for (uint32_t k = 0; k < Q.n_indices; k++) f_idx.push_back(dense(Q.indices[k]));
        for (uint32_t r : dq) { // only the residues the query touches travel to the device
            if (r >= Q.n_residues) {
                free(h_flags);
                return fd_fail(ctx, FD_ERR_ARG, "fd_verify_query: residue index outside the query structure");
            }
            for (int x = 0; x < 3; x++) {
                q_ca.push_back(Q.ca_xyz[3 * r + x]);
                q_cb.push_back(Q.cb_xyz[3 * r + x]);
            }
        }
        descs[q] = d;
    }
    for (uint64_t c = 0; c < n_cand; c++) {
        if (cand_query[c] >= nq || cand_nid[c] >= ctx->store.n_structs) {
            free(h_flags);
            return fd_fail(ctx, FD_ERR_ARG, "candidate query / structure id out of range");
        }
    }
    if (n_cand == 0) {
        *out_records = (fd_match_record *)malloc(sizeof(fd_match_record));
        *out_flags = h_flags;
        return FD_OK;
    }
    cudaStream_t s = ctx->stream;
    DevBuf<VQDesc> d_desc;
    DevBuf<VHash> d_hash;
    DevBuf<VAad> d_aad;
    DevBuf<uint8_t> d_idx, d_flags;
    DevBuf<uint32_t> d_cq, d_cn;
    DevBuf<float> d_qca, d_qcb;
    DevBuf<unsigned int> d_count;
    DevBuf<fd_match_record> d_out;
    FD_CUDA(ctx, d_desc.alloc(nq));
    FD_CUDA(ctx, d_hash.alloc(f_hash.size()));
    FD_CUDA(ctx, d_aad.alloc(f_aad.size()));
    FD_CUDA(ctx, d_idx.alloc(f_idx.size()));
    FD_CUDA(ctx, d_qca.alloc(q_ca.size()));
    FD_CUDA(ctx, d_qcb.alloc(q_cb.size()));
    FD_CUDA(ctx, d_cq.alloc(n_cand));
    FD_CUDA(ctx, d_cn.alloc(n_cand));
    FD_CUDA(ctx, d_flags.alloc(n_cand));
    FD_CUDA(ctx, d_count.alloc(1));
    FD_CUDA(ctx, cudaMemcpyAsync(d_desc.p, descs.data(), nq * sizeof(VQDesc), cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_hash.p, f_hash.data(), f_hash.size() * sizeof(VHash), cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_aad.p, f_aad.data(), f_aad.size() * sizeof(VAad), cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_idx.p, f_idx.data(), f_idx.size(), cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_qca.p, q_ca.data(), q_ca.size() * 4, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_qcb.p, q_cb.data(), q_cb.size() * 4, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_cq.p, cand_query, n_cand * 4, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_cn.p, cand_nid, n_cand * 4, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemsetAsync(d_flags.p, 0, n_cand, s));
    const FdDeviceStore &S = ctx->store;
    StoreView sv{S.row_offsets, S.n_xyz, S.ca_xyz, S.cb_xyz, S.aa, S.cb_valid};
    fdg::HashParams hp = fdg::make_params(params->nbin_dist, params->nbin_angle, params->dist_cutoff);
    uint64_t cap = std::max<uint64_t>(1024, 2 * n_cand);
    fd_match_record *h_out = nullptr;
    unsigned int produced = 0;
    StageTimer st(ctx, "verify");
    for (int attempt = 0; attempt < 2; attempt++) {
        FD_CUDA(ctx, d_out.alloc(cap));
        FD_CUDA(ctx, cudaMemsetAsync(d_count.p, 0, 4, s));
        FD_CUDA(ctx, cudaMemsetAsync(d_flags.p, 0, n_cand, s));
        FD_LAUNCH(ctx, k6_verify, (uint32_t)n_cand, V_THREADS, 0, sv, d_desc.p, d_hash.p, d_aad.p, d_idx.p,
                  d_qca.p, d_qcb.p, d_cq.p, d_cn.p, (uint32_t)n_cand, hp, ca_dist_cutoff, skip_ca_match, d_out.p,
                  d_count.p, (uint32_t)std::min<uint64_t>(cap, 0xffffffffu), d_flags.p);
        FD_CUDA(ctx, cudaMemcpyAsync(&produced, d_count.p, 4, cudaMemcpyDeviceToHost, s));
        FD_CUDA(ctx, cudaStreamSynchronize(s));
        if (produced <= cap) break;
        cap = produced; // the pool was too small: size it exactly and run once more
    }
    h_out = (fd_match_record *)malloc(std::max<uint64_t>(produced, 1) * sizeof(fd_match_record));
    if (!h_out) {
        free(h_flags);
        return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
    }
    FD_CUDA(ctx, cudaMemcpyAsync(h_out, d_out.p, (size_t)produced * sizeof(fd_match_record), cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, cudaMemcpyAsync(h_flags, d_flags.p, n_cand, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, st.finish());
    for (uint64_t c = 0; c < n_cand; c++)
        if (q_unfit[cand_query[c]]) h_flags[c] = 1;
    // records arrive in arbitrary CTA order (each candidate's block is contiguous and in component order):
    // counting sort by candidate
    if (produced) {
        std::vector<uint32_t> start(n_cand + 1, 0);
        for (unsigned int k = 0; k < produced; k++) start[h_out[k].cand + 1]++;
        for (uint64_t c = 0; c < n_cand; c++) start[c + 1] += start[c];
        fd_match_record *sorted = (fd_match_record *)malloc((size_t)produced * sizeof(fd_match_record));
        if (!sorted) {
            free(h_out);
            free(h_flags);
            return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
        }
        for (unsigned int k = 0; k < produced; k++) sorted[start[h_out[k].cand]++] = h_out[k];
        free(h_out);
        h_out = sorted;
    }
    *out_records = h_out;
    *out_n = produced;
    *out_flags = h_flags;
    return FD_OK;
}
