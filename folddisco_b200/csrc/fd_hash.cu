// fd_hash.cu -- K1: all-residue-pair geometric hashing (index build path (i)).
//
// Replaces the rayon loop of Folddisco::collect_and_count / add_entries
// (reference src/controller/mod.rs:274-441): get_geometric_hash_as_u32_from_structure
// (src/controller/feature.rs:198-231) over every structure, followed by sort+dedup (:343-344).
//
// Work decomposition.  A tile = 32 consecutive rows i of one structure against all of its columns j.
// One CTA (256 threads) per tile:
//   phase A  the CTA's warps sweep the tile's pairs in rounds of 8 rows x 512 columns and run only the cheap
//            C-alpha distance test; survivors (about 115 per row on real proteins) are compacted into a
//            shared-memory queue with one ballot + one shared atomic per warp;
//   phase B  the queue is drained densely, one survivor per thread: ~250 f32 flops plus three binary64
//            sincos, one acos and two atan2 (fd_geom.cuh), so no lane idles on pairs beyond the cutoff.
// The kernel is ALU bound (37 B of coordinates in, ~115 x ~1 kflop out per residue), so coordinates are read
// straight from global memory through L1/L2; nothing is staged by TMA.
// Output is a flat list of 64-bit keys in arbitrary order (one warp-aggregated global atomic per warp per
// drain round); the keys are sorted afterwards, which also performs the per-structure dedup.
#include <cub/cub.cuh>

#include "fd_common.cuh"
#include "fd_geom.cuh"
#include "fd_hashtypes.cuh"

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_TILE_ROWS = 32;
constexpr int K1_COL_CHUNK = 512;
constexpr int K1_QUEUE_CAP = K1_WARPS * K1_COL_CHUNK; // a round can never overflow the queue

struct Tile {
    uint32_t s;  // structure (row in the batch)
    uint32_t i0; // first row of the tile
};

struct BatchView {
    const uint64_t *row_offsets;
    const float *n_xyz, *ca_xyz, *cb_xyz;
    const uint8_t *aa, *cb_valid;
};

__device__ __forceinline__ fdg::V3 ld3(const float *p, uint64_t r) {
    return {p[3 * r], p[3 * r + 1], p[3 * r + 2]};
}

__device__ __forceinline__ bool residue_ok(const BatchView &b, uint64_t r) {
    return b.aa[r] != 255 && (b.cb_valid == nullptr || b.cb_valid[r] != 0);
}

// MODE 0: count survivors per tile (sizes the key buffer exactly).
// MODE 1: emit key = hash << 32 | (first_id + s)   (posting build)
// MODE 2: emit key = s << 32 | hash                (per-structure sorted unique output)
// MODE 3: emit key = (s - first_id) << 30 | hash, val = i << 16 | j   (pair table of the structure store)
// TYPED: any encoding of fd_hashtypes.cuh and `--multiple-bins` (tp): the screen uses the encoding's own cutoff
// distance, a survivor's feature is computed once and hashed once per (nbin_dist, nbin_angle) pair.  The default
// encoding with a single bin pair keeps the tuned route (TYPED = false: hp, pair_hash_auto).
template <int MODE, bool TYPED = false>
__global__ void __launch_bounds__(K1_THREADS)
    k1_pair_hash(BatchView b, const Tile *tiles, uint32_t n_tiles, fdg::HashParams hp, uint64_t first_id,
                 uint64_t hash_lo, uint64_t hash_hi, uint64_t *out_keys, unsigned long long *out_count,
                 uint32_t *out_vals = nullptr, fdg::TypedParams tp = fdg::TypedParams()) {
    __shared__ uint32_t q_ij[MODE == 0 ? 1 : K1_QUEUE_CAP]; // i_local << 16 | j
    __shared__ float q_d[MODE == 0 ? 1 : K1_QUEUE_CAP];     // ca_dist of the survivor
    __shared__ uint32_t q_n;
    __shared__ unsigned long long tile_count;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const Tile tile = tiles[t];
        const uint64_t base = b.row_offsets[tile.s];
        const uint32_t n = (uint32_t)(b.row_offsets[tile.s + 1] - base);
        const uint32_t rows = min((uint32_t)K1_TILE_ROWS, n - tile.i0);
        if (threadIdx.x == 0) {
            q_n = 0;
            tile_count = 0;
        }
        __syncthreads();
        unsigned long long my_count = 0;
        // TertiaryInteraction needs no CB; it and Hybrid exist only for residues with both sequence neighbours
        const bool need_cb = !TYPED || fdg::ht_needs_cb(tp.type);
        const bool need_nb = TYPED && fdg::ht_needs_neighbours(tp.type);
        for (uint32_t r0 = 0; r0 < rows; r0 += K1_WARPS) {
            const uint32_t il = r0 + warp; // this warp's row within the tile
            const bool row_live = il < rows;
            const uint32_t i = tile.i0 + il;
            fdg::V3 cai = {0.f, 0.f, 0.f}, cbi = {0.f, 0.f, 0.f};
            bool iok = false;
            if (row_live) {
                iok = need_cb ? residue_ok(b, base + i) : b.aa[base + i] != 255;
                if (need_nb) iok = iok && i > 0 && i + 1 < n;
                cai = ld3(b.ca_xyz, base + i);
                if (TYPED) cbi = ld3(b.cb_xyz, base + i);
            }
            for (uint32_t c0 = 0; c0 < n; c0 += K1_COL_CHUNK) {
                // ---- phase A: distance screen ----
                const uint32_t cend = min(n, c0 + K1_COL_CHUNK);
                for (uint32_t j = c0 + lane; j < c0 + K1_COL_CHUNK; j += 32) { // uniform trip count per warp
                    bool pass = false;
                    float d = 0.f;
                    if (iok && j < cend && j != i && (need_cb ? residue_ok(b, base + j) : b.aa[base + j] != 255) &&
                        (!need_nb || (j > 0 && j + 1 < n))) {
                        if (TYPED)
                            d = fdg::typed_screen_dist(tp.type, cai, cbi, ld3(b.ca_xyz, base + j), ld3(b.cb_xyz, base + j));
                        else
                            d = fdg::dist(cai, ld3(b.ca_xyz, base + j));
                        pass = !(d > hp.dist_cutoff); // reference: `if ca_dist > dist_cutoff { return None }`
                    }
                    if (MODE == 0) {
                        my_count += pass ? (TYPED ? tp.n_bins : 1u) : 0u;
                    } else {
                        const uint32_t m = __ballot_sync(0xffffffffu, pass);
                        if (m) {
                            uint32_t pos = 0;
                            if (lane == 0) pos = atomicAdd(&q_n, __popc(m));
                            pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1));
                            if (pass) {
                                q_ij[pos] = (il << 16) | j;
                                q_d[pos] = d;
                            }
                        }
                    }
                }
                if (MODE != 0) {
                    __syncthreads();
                    // ---- phase B: dense drain ----
                    const uint32_t qn = q_n;
                    for (uint32_t k0 = 0; TYPED && k0 < qn; k0 += K1_THREADS) {
                        const uint32_t k = k0 + threadIdx.x;
                        float f[9];
                        if (k < qn) {
                            const uint32_t ij = q_ij[k];
                            const uint64_t ri = base + tile.i0 + (ij >> 16), rj = base + (ij & 0xffffu);
                            fdg::Nbr nb;
                            if (need_nb) { // both residues are interior (screened in phase A)
                                nb.ca_pre1 = ld3(b.ca_xyz, ri - 1);
                                nb.ca_next1 = ld3(b.ca_xyz, ri + 1);
                                nb.ca_pre2 = ld3(b.ca_xyz, rj - 1);
                                nb.ca_next2 = ld3(b.ca_xyz, rj + 1);
                                nb.seq_dist = (float)(ij & 0xffffu) - (float)(tile.i0 + (ij >> 16)); // j as f32 - i as f32
                            }
                            fdg::typed_feature(tp.type, ld3(b.n_xyz, ri), ld3(b.ca_xyz, ri), ld3(b.cb_xyz, ri),
                                               ld3(b.n_xyz, rj), ld3(b.ca_xyz, rj), ld3(b.cb_xyz, rj),
                                               (float)(b.aa[ri] & 0x7Fu), (float)(b.aa[rj] & 0x7Fu), q_d[k], f, &nb);
                        }
                        for (uint32_t bi = 0; bi < tp.n_bins; bi++) { // uniform trip count
                            bool emit = false;
                            uint64_t key = 0;
                            if (k < qn) {
                                const uint32_t h = fdg::typed_hash(tp.type, f, tp.nbd[bi], tp.nba[bi]);
                                emit = (uint64_t)h >= hash_lo && (uint64_t)h < hash_hi;
                                key = MODE == 1 ? ((uint64_t)h << 32) | (first_id + tile.s) : ((uint64_t)tile.s << 32) | h;
                            }
                            const uint32_t m = __ballot_sync(0xffffffffu, emit);
                            if (m) {
                                unsigned long long pos = 0;
                                if (lane == 0) pos = atomicAdd(out_count, (unsigned long long)__popc(m));
                                pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1));
                                if (emit) out_keys[pos] = key;
                            }
                        }
                    }
                    for (uint32_t k0 = 0; !TYPED && k0 < qn; k0 += K1_THREADS) {
                        const uint32_t k = k0 + threadIdx.x;
                        bool emit = false;
                        uint64_t key = 0;
                        uint32_t val = 0;
                        if (k < qn) {
                            const uint32_t ij = q_ij[k];
                            const uint64_t ri = base + tile.i0 + (ij >> 16), rj = base + (ij & 0xffffu);
                            const uint32_t h = fdg::pair_hash_auto(ld3(b.n_xyz, ri), ld3(b.ca_xyz, ri), ld3(b.cb_xyz, ri),
                                                              ld3(b.n_xyz, rj), ld3(b.ca_xyz, rj), ld3(b.cb_xyz, rj),
                                                              b.aa[ri] & 0x7Fu, b.aa[rj] & 0x7Fu, q_d[k], hp);
                            emit = (uint64_t)h >= hash_lo && (uint64_t)h < hash_hi;
                            key = MODE == 1   ? ((uint64_t)h << 32) | (first_id + tile.s)
                                  : MODE == 3 ? (((uint64_t)tile.s - first_id) << 30) | h
                                              : ((uint64_t)tile.s << 32) | h;
                            val = ((tile.i0 + (ij >> 16)) << 16) | (ij & 0xffffu);
                        }
                        const uint32_t m = __ballot_sync(0xffffffffu, emit);
                        if (m) {
                            unsigned long long pos = 0;
                            if (lane == 0) pos = atomicAdd(out_count, (unsigned long long)__popc(m));
                            pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1));
                            if (emit) {
                                out_keys[pos] = key;
                                if (MODE == 3) out_vals[pos] = val;
                            }
                        }
                    }
                    __syncthreads();
                    if (threadIdx.x == 0) q_n = 0;
                    __syncthreads();
                }
            }
        }
        if (MODE == 0) {
            // block reduce of my_count -> one global atomic per tile
            for (int o = 16; o > 0; o >>= 1) my_count += __shfl_down_sync(0xffffffffu, my_count, o);
            if (lane == 0 && my_count) atomicAdd(&tile_count, my_count);
            __syncthreads();
            if (threadIdx.x == 0 && tile_count) atomicAdd(out_count, tile_count);
            __syncthreads();
        }
    }
}

__global__ void k1_split_keys(const uint64_t *keys, uint64_t n, uint32_t *hashes, uint64_t *row_counts) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint64_t key = keys[k];
    hashes[k] = (uint32_t)key;
    atomicAdd((unsigned long long *)&row_counts[key >> 32], 1ull);
}

int build_tiles(const fd_struct_batch *b, std::vector<Tile> &tiles) {
    for (uint64_t s = 0; s < b->n_structs; s++) {
        uint64_t n = b->row_offsets[s + 1] - b->row_offsets[s];
        if (n > 65535) return FD_ERR_LIMIT; // DEFAULT_MAX_RESIDUE of the reference (src/controller/mod.rs:40)
        for (uint64_t i0 = 0; i0 < n; i0 += K1_TILE_ROWS) tiles.push_back({(uint32_t)s, (uint32_t)i0});
    }
    return FD_OK;
}

// Runs the count pass and the emit pass; returns the device key array.
template <int MODE>
int run_pair_hash(fd_ctx *ctx, const fd_struct_batch *batch, const fd_hash_params *params, uint64_t first_id,
                  uint64_t hash_lo, uint64_t hash_hi, DevBuf<uint64_t> &keys, uint64_t *n_keys) {
    FdDeviceBatch d;
    FD_TRY(fd_upload_batch(ctx, batch, &d));
    std::vector<Tile> tiles;
    if (build_tiles(batch, tiles) != FD_OK)
        return fd_fail(ctx, FD_ERR_LIMIT, "structure with more than 65535 residues (reference max_residue)");
    *n_keys = 0;
    if (tiles.empty()) return FD_OK;
    if (tiles.size() > 0xffffffffull) return fd_fail(ctx, FD_ERR_LIMIT, "too many tiles in one batch; split it");
    DevBuf<Tile> d_tiles;
    DevBuf<unsigned long long> d_count;
    FD_CUDA(ctx, d_tiles.alloc(tiles.size()));
    FD_CUDA(ctx, d_count.alloc(1));
    FD_CUDA(ctx, cudaMemcpyAsync(d_tiles.p, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice,
                                 ctx->stream));
    FD_CUDA(ctx, cudaMemsetAsync(d_count.p, 0, sizeof(unsigned long long), ctx->stream));
    BatchView v{d.row_offsets.p, d.n_xyz.p, d.ca_xyz.p, d.cb_xyz.p, d.aa.p, batch->cb_valid ? d.cb_valid.p : nullptr};
    fdg::HashParams hp = fdg::make_params(params->nbin_dist, params->nbin_angle, params->dist_cutoff);
    fdg::TypedParams tp;
    if (const char *why = fdg::typed_params_from(params, &tp)) return fd_fail(ctx, FD_ERR_ARG, why);
    const bool typed = !fdg::ht_default_route(params);
    const uint32_t n_tiles = (uint32_t)tiles.size();
    const uint32_t grid = (uint32_t)std::min<uint64_t>(n_tiles, (uint64_t)ctx->num_sms * 8);
    unsigned long long total = 0;
    {
        StageTimer st(ctx, "hash");
        if (typed)
            FD_LAUNCH(ctx, (k1_pair_hash<0, true>), grid, K1_THREADS, 0, v, d_tiles.p, n_tiles, hp, first_id, hash_lo,
                      hash_hi, (uint64_t *)nullptr, d_count.p, (uint32_t *)nullptr, tp);
        else
            FD_LAUNCH(ctx, k1_pair_hash<0>, grid, K1_THREADS, 0, v, d_tiles.p, n_tiles, hp, first_id, hash_lo, hash_hi,
                      (uint64_t *)nullptr, d_count.p);
        FD_CUDA(ctx, cudaMemcpyAsync(&total, d_count.p, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
        FD_CUDA(ctx, st.finish());
    }
    {
        // the key buffer comes from the stream-ordered pool; growing the pool by gigabytes is host time, not kernel time
        HostTimer ht(ctx, "hash_alloc");
        FD_CUDA(ctx, keys.alloc(total));
        FD_CUDA(ctx, cudaMemsetAsync(d_count.p, 0, sizeof(unsigned long long), ctx->stream));
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    {
        StageTimer st(ctx, "hash");
        if (total && typed)
            FD_LAUNCH(ctx, (k1_pair_hash<MODE, true>), grid, K1_THREADS, 0, v, d_tiles.p, n_tiles, hp, first_id, hash_lo,
                      hash_hi, keys.p, d_count.p, (uint32_t *)nullptr, tp);
        else if (total)
            FD_LAUNCH(ctx, k1_pair_hash<MODE>, grid, K1_THREADS, 0, v, d_tiles.p, n_tiles, hp, first_id, hash_lo,
                      hash_hi, keys.p, d_count.p);
        FD_CUDA(ctx, cudaMemcpyAsync(&total, d_count.p, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
        FD_CUDA(ctx, st.finish());
    }
    *n_keys = total;
    return FD_OK;
}

} // namespace

int fd_upload_batch(fd_ctx *ctx, const fd_struct_batch *b, FdDeviceBatch *d) {
    if (!b || !b->row_offsets || (b->n_structs && (!b->n_xyz || !b->ca_xyz || !b->cb_xyz || !b->aa)))
        return fd_fail(ctx, FD_ERR_ARG, "fd_struct_batch: NULL array");
    const uint64_t S = b->n_structs, R = b->row_offsets[S];
    if (b->row_offsets[0] != 0) return fd_fail(ctx, FD_ERR_ARG, "fd_struct_batch: row_offsets[0] must be 0");
    for (uint64_t s = 0; s < S; s++)
        if (b->row_offsets[s + 1] < b->row_offsets[s])
            return fd_fail(ctx, FD_ERR_ARG, "fd_struct_batch: row_offsets not monotone");
    d->n_structs = S;
    d->n_res = R;
    FD_CUDA(ctx, d->row_offsets.alloc(S + 1));
    FD_CUDA(ctx, d->n_xyz.alloc(3 * R));
    FD_CUDA(ctx, d->ca_xyz.alloc(3 * R));
    FD_CUDA(ctx, d->cb_xyz.alloc(3 * R));
    FD_CUDA(ctx, d->aa.alloc(R));
    FD_CUDA(ctx, d->cb_valid.alloc(R));
    cudaStream_t st = ctx->stream;
    FD_CUDA(ctx, cudaMemcpyAsync(d->row_offsets.p, b->row_offsets, (S + 1) * 8, cudaMemcpyHostToDevice, st));
    if (R) {
        FD_CUDA(ctx, cudaMemcpyAsync(d->n_xyz.p, b->n_xyz, 12 * R, cudaMemcpyHostToDevice, st));
        FD_CUDA(ctx, cudaMemcpyAsync(d->ca_xyz.p, b->ca_xyz, 12 * R, cudaMemcpyHostToDevice, st));
        FD_CUDA(ctx, cudaMemcpyAsync(d->cb_xyz.p, b->cb_xyz, 12 * R, cudaMemcpyHostToDevice, st));
        FD_CUDA(ctx, cudaMemcpyAsync(d->aa.p, b->aa, R, cudaMemcpyHostToDevice, st));
        if (b->cb_valid) FD_CUDA(ctx, cudaMemcpyAsync(d->cb_valid.p, b->cb_valid, R, cudaMemcpyHostToDevice, st));
    }
    return FD_OK;
}

// ---- pair table of the attached structure store ----
namespace {
__global__ void k1_table_offsets(const uint64_t *sorted_keys, uint64_t n, uint32_t n_structs_chunk, uint64_t base,
                                 uint64_t *offsets /* [s0 ..] of the chunk, + 1 */) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_structs_chunk) return;
    const uint64_t want = (uint64_t)s << 30; // first key of structure s
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    offsets[s] = base + lo;
}
__global__ void k1_table_split(const uint64_t *sorted_keys, const uint32_t *sorted_vals, uint64_t n, uint32_t *hash,
                               uint32_t *ij) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    hash[k] = (uint32_t)(sorted_keys[k] & 0x3fffffffull);
    ij[k] = sorted_vals[k];
}
// dir[s][p] = first entry (relative to offsets[s]) of structure s with hash >> 20 >= p
__global__ void k1_table_dir(const uint32_t *hash, const uint64_t *offsets, uint64_t n_structs, uint32_t *dir) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_structs * FD_PT_DIR) return;
    const uint64_t s = t / FD_PT_DIR;
    const uint32_t p = (uint32_t)(t % FD_PT_DIR);
    const uint32_t *h = hash + offsets[s];
    uint32_t lo = 0, hi = (uint32_t)(offsets[s + 1] - offsets[s]);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((h[mid] >> 20) < p) lo = mid + 1;
        else hi = mid;
    }
    dir[t] = lo;
}
} // namespace

extern "C" int fd_store_build_pair_table(fd_ctx *ctx, const fd_hash_params *params, uint64_t max_bytes,
                                         uint64_t *out_bytes) {
    if (!ctx) return FD_ERR_ARG;
    if (!params) return fd_fail(ctx, FD_ERR_ARG, "fd_store_build_pair_table: NULL argument");
    if (!ctx->store.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_store_build_pair_table: no structure store attached");
    if (ctx->borrowed) return fd_fail(ctx, FD_ERR_STATE, "fd_store_build_pair_table: a forked context shares its parent's store");
    if (!fdg::ht_default_route(params))
        return fd_fail(ctx, FD_ERR_ARG, "fd_store_build_pair_table: the pair table holds 30-bit PDBTrRosetta hashes of one bin pair; other encodings / multiple bins are verified through fd_candidate_edges_batch");
    FD_ENTER(ctx);
    FdDeviceStore &S = ctx->store;
    cudaStream_t st = ctx->stream;
    if (out_bytes) *out_bytes = 0;
    FD_CUDA(ctx, cudaStreamSynchronize(st));
    cudaFree(S.pt.offsets);
    cudaFree(S.pt.hash);
    cudaFree(S.pt.ij);
    cudaFree(S.pt.dir);
    S.pt = FdPairTable();
    const uint64_t NS = S.n_structs;
    if (NS == 0) return FD_OK;
    const std::vector<uint64_t> &ro = S.h_row_offsets;
    BatchView v{S.row_offsets, S.n_xyz, S.ca_xyz, S.cb_xyz, S.aa, S.cb_valid};
    fdg::HashParams hp = fdg::make_params(params->nbin_dist, params->nbin_angle, params->dist_cutoff);
    // chunks of structures: at most ~2^28 residue-pairs-worth of tiles and 2^20 structures (key = s_local << 30 | hash)
    struct Chunk {
        uint64_t s0, s1, n_keys;
        std::vector<Tile> tiles;
    };
    std::vector<Chunk> chunks;
    const uint64_t chunk_res = 24ull * 1000 * 1000; // ~2.8 G table entries per chunk at 115 pairs per residue / 8 = sort scratch 3 x 12 B each
    {
        uint64_t s = 0;
        while (s < NS) {
            Chunk c;
            c.s0 = s;
            uint64_t r = 0;
            while (s < NS && (s == c.s0 || (r + (ro[s + 1] - ro[s]) <= chunk_res / 8 && s - c.s0 < (1u << 20)))) {
                const uint64_t n = ro[s + 1] - ro[s];
                for (uint64_t i0 = 0; i0 < n; i0 += K1_TILE_ROWS) c.tiles.push_back({(uint32_t)s, (uint32_t)i0});
                r += n;
                s++;
            }
            c.s1 = s;
            c.n_keys = 0;
            chunks.push_back(std::move(c));
        }
    }
    StageTimer stt(ctx, "pair_table");
    // pass 1: exact sizes
    DevBuf<unsigned long long> d_count;
    FD_CUDA(ctx, d_count.alloc(1));
    uint64_t total = 0;
    for (auto &c : chunks) {
        DevBuf<Tile> d_tiles;
        FD_CUDA(ctx, d_tiles.alloc(c.tiles.size()));
        FD_CUDA(ctx, cudaMemcpyAsync(d_tiles.p, c.tiles.data(), c.tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice, st));
        FD_CUDA(ctx, cudaMemsetAsync(d_count.p, 0, 8, st));
        const uint32_t n_tiles = (uint32_t)c.tiles.size();
        const uint32_t grid = (uint32_t)std::min<uint64_t>(n_tiles, (uint64_t)ctx->num_sms * 8);
        FD_LAUNCH(ctx, k1_pair_hash<0>, grid, K1_THREADS, 0, v, d_tiles.p, n_tiles, hp, 0, 0, 1ull << 32,
                  (uint64_t *)nullptr, d_count.p, (uint32_t *)nullptr);
        unsigned long long n = 0;
        FD_CUDA(ctx, cudaMemcpyAsync(&n, d_count.p, 8, cudaMemcpyDeviceToHost, st));
        FD_CUDA(ctx, cudaStreamSynchronize(st));
        c.n_keys = n;
        total += n;
    }
    const uint64_t bytes = total * 8 + (NS + 1) * 8 + NS * FD_PT_DIR * 4ull;
    if (out_bytes) *out_bytes = bytes;
    if (max_bytes && bytes > max_bytes) {
        FD_CUDA(ctx, stt.finish());
        return fd_fail(ctx, FD_ERR_LIMIT, "pair table larger than the caller's budget (verification keeps re-hashing candidates)");
    }
    FD_CUDA(ctx, cudaMalloc(&S.pt.offsets, (NS + 1) * 8));
    FD_CUDA(ctx, cudaMalloc(&S.pt.hash, std::max<uint64_t>(total, 1) * 4));
    FD_CUDA(ctx, cudaMalloc(&S.pt.ij, std::max<uint64_t>(total, 1) * 4));
    FD_CUDA(ctx, cudaMalloc(&S.pt.dir, NS * FD_PT_DIR * 4ull));
    uint64_t base = 0;
    for (auto &c : chunks) {
        const uint64_t n = c.n_keys;
        const uint32_t ns_chunk = (uint32_t)(c.s1 - c.s0);
        DevBuf<Tile> d_tiles;
        DevBuf<uint64_t> k_in, k_out;
        DevBuf<uint32_t> v_in, v_out;
        DevBuf<uint8_t> tmp;
        FD_CUDA(ctx, d_tiles.alloc(c.tiles.size()));
        FD_CUDA(ctx, k_in.alloc(n));
        FD_CUDA(ctx, k_out.alloc(n));
        FD_CUDA(ctx, v_in.alloc(n));
        FD_CUDA(ctx, v_out.alloc(n));
        FD_CUDA(ctx, cudaMemcpyAsync(d_tiles.p, c.tiles.data(), c.tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice, st));
        FD_CUDA(ctx, cudaMemsetAsync(d_count.p, 0, 8, st));
        const uint32_t n_tiles = (uint32_t)c.tiles.size();
        const uint32_t grid = (uint32_t)std::min<uint64_t>(n_tiles, (uint64_t)ctx->num_sms * 8);
        if (n)
            FD_LAUNCH(ctx, k1_pair_hash<3>, grid, K1_THREADS, 0, v, d_tiles.p, n_tiles, hp, c.s0, 0, 1ull << 32, k_in.p,
                      d_count.p, v_in.p);
        int end_bit = 31;
        for (uint32_t x = ns_chunk; x > 1; x >>= 1) end_bit++;
        end_bit = std::min(64, end_bit + 1);
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in.p, k_out.p, v_in.p, v_out.p, n, 0, end_bit, st);
        FD_CUDA(ctx, tmp.alloc(tb));
        if (n) {
            FD_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tb, k_in.p, k_out.p, v_in.p, v_out.p, n, 0, end_bit, st));
            ctx->launches += 4;
            FD_LAUNCH(ctx, k1_table_split, fd_div_up(n, 256), 256, 0, k_out.p, v_out.p, n, S.pt.hash + base, S.pt.ij + base);
        }
        FD_LAUNCH(ctx, k1_table_offsets, fd_div_up((uint64_t)ns_chunk + 1, 256), 256, 0, k_out.p, n, ns_chunk, base,
                  S.pt.offsets + c.s0);
        FD_CUDA(ctx, cudaStreamSynchronize(st));
        base += n;
    }
    FD_LAUNCH(ctx, k1_table_dir, fd_div_up(NS * FD_PT_DIR, 256), 256, 0, S.pt.hash, S.pt.offsets, NS, S.pt.dir);
    FD_CUDA(ctx, stt.finish());
    S.pt.n = total;
    S.pt.nbin_dist = params->nbin_dist;
    S.pt.nbin_angle = params->nbin_angle;
    S.pt.dist_cutoff = params->dist_cutoff;
    S.pt.built = true;
    return FD_OK;
}

extern "C" {

int fd_hash_structures(fd_ctx *ctx, const fd_struct_batch *batch, const fd_hash_params *params,
                       uint32_t **out_hashes, uint64_t **out_row_offsets) {
    if (!ctx) return FD_ERR_ARG;
    if (!batch || !params || !out_hashes || !out_row_offsets)
        return fd_fail(ctx, FD_ERR_ARG, "fd_hash_structures: NULL argument");
    FD_ENTER(ctx);
    *out_hashes = nullptr;
    *out_row_offsets = nullptr;
    const uint64_t S = batch->n_structs;
    DevBuf<uint64_t> keys;
    uint64_t n_keys = 0;
    FD_TRY(run_pair_hash<2>(ctx, batch, params, 0, 0, 1ull << 32, keys, &n_keys));
    uint64_t *h_rows = (uint64_t *)calloc(S + 1, sizeof(uint64_t));
    if (!h_rows) return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
    uint64_t n_unique = 0;
    uint32_t *h_hashes = nullptr;
    if (n_keys) {
        // sort by (structure, hash) then drop duplicates == per-structure sort_unstable + dedup
        StageTimer st(ctx, "hash");
        DevBuf<uint64_t> sorted, uniq, d_nsel, d_rows;
        DevBuf<uint32_t> d_hashes;
        DevBuf<uint8_t> tmp;
        FD_CUDA(ctx, sorted.alloc(n_keys));
        FD_CUDA(ctx, uniq.alloc(n_keys));
        FD_CUDA(ctx, d_nsel.alloc(1));
        FD_CUDA(ctx, d_rows.alloc(S + 1));
        int end_bit = 32;
        for (uint64_t v = S; v > 1; v >>= 1) end_bit++;
        end_bit = std::min(64, end_bit + 1);
        size_t tb1 = 0, tb2 = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tb1, keys.p, sorted.p, n_keys, 0, end_bit, ctx->stream);
        cub::DeviceSelect::Unique(nullptr, tb2, sorted.p, uniq.p, d_nsel.p, n_keys, ctx->stream);
        FD_CUDA(ctx, tmp.alloc(std::max(tb1, tb2)));
        size_t tb = std::max(tb1, tb2);
        FD_CUDA(ctx, cub::DeviceRadixSort::SortKeys(tmp.p, tb, keys.p, sorted.p, n_keys, 0, end_bit, ctx->stream));
        ctx->launches += 4;
        tb = std::max(tb1, tb2);
        FD_CUDA(ctx, cub::DeviceSelect::Unique(tmp.p, tb, sorted.p, uniq.p, d_nsel.p, n_keys, ctx->stream));
        ctx->launches += 2;
        FD_CUDA(ctx, cudaMemcpyAsync(&n_unique, d_nsel.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        FD_CUDA(ctx, d_hashes.alloc(n_unique));
        FD_CUDA(ctx, cudaMemsetAsync(d_rows.p, 0, (S + 1) * 8, ctx->stream));
        FD_LAUNCH(ctx, k1_split_keys, fd_div_up(n_unique, 256), 256, 0, uniq.p, n_unique, d_hashes.p, d_rows.p + 1);
        h_hashes = (uint32_t *)malloc(std::max<uint64_t>(n_unique, 1) * sizeof(uint32_t));
        if (!h_hashes) {
            free(h_rows);
            return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
        }
        FD_CUDA(ctx, cudaMemcpyAsync(h_hashes, d_hashes.p, n_unique * 4, cudaMemcpyDeviceToHost, ctx->stream));
        FD_CUDA(ctx, cudaMemcpyAsync(h_rows, d_rows.p, (S + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        FD_CUDA(ctx, st.finish());
        for (uint64_t s = 0; s < S; s++) h_rows[s + 1] += h_rows[s];
    } else {
        h_hashes = (uint32_t *)malloc(sizeof(uint32_t));
    }
    *out_hashes = h_hashes;
    *out_row_offsets = h_rows;
    return FD_OK;
}

int fd_build_index(fd_ctx *ctx, const fd_struct_batch *batch, const fd_hash_params *params, uint64_t first_id,
                   uint64_t hash_lo, uint64_t hash_hi, fd_index_buffers *out) {
    if (!ctx) return FD_ERR_ARG;
    if (!batch || !params || !out) return fd_fail(ctx, FD_ERR_ARG, "fd_build_index: NULL argument");
    if (first_id + batch->n_structs > 0xffffffffull)
        return fd_fail(ctx, FD_ERR_LIMIT, "structure ids must fit in 32 bits");
    FD_ENTER(ctx);
    memset(out, 0, sizeof(*out));
    DevBuf<uint64_t> keys, tmp;
    uint64_t n_keys = 0;
    FD_TRY(run_pair_hash<1>(ctx, batch, params, first_id, hash_lo, hash_hi, keys, &n_keys));
    FD_CUDA(ctx, tmp.alloc(n_keys));
    return fd_postings_from_keys(ctx, keys.p, n_keys, tmp.p, out);
}

} // extern "C"
