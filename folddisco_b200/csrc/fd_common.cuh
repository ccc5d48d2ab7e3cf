// fd_common.cuh -- context, error handling, device buffers, stage timers shared by the .cu files.
#pragma once
#include <functional>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <map>
#include <string>
#include <vector>

#include "../../include/folddisco_b200.h"

#define FD_NUM_SMS_FALLBACK 148

// persistent host worker pool (fd_ctx.cu): fn(0) .. fn(nt - 1) concurrently, fn(0) on the caller
void fd_parallel(int nt, const std::function<void(int)> &fn);

struct FdStage {
    double ms = 0.0;
    uint64_t launches = 0;
};

// Device copy of the attached inverted index (layout documented in DESIGN.md "HBM layout").
struct FdDeviceIndex {
    bool attached = false;
    uint64_t count = 0;       // number of distinct hashes
    uint64_t value_bytes = 0; // posting bytes
    uint64_t n_structs = 0;
    uint32_t *hashes = nullptr;   // [count] ascending
    uint64_t *offsets = nullptr;  // [count+1] byte offsets into values
    uint8_t *values = nullptr;    // [value_bytes + 16] delta+LEB128 postings, padded for vector loads
    uint32_t *counts = nullptr;   // [count] postings per list
    uint32_t *dir = nullptr;      // [FD_DIR_SIZE+1] hashes[] position of the first hash with (hash >> FD_DIR_SHIFT) >= b
    uint32_t *skip_id = nullptr;  // [n_skip] see fd_query.cu
    uint8_t *skip_off = nullptr;  // [n_skip]
    uint64_t n_skip = 0;
    uint32_t *nres = nullptr; // [n_structs]
    float *plddt = nullptr;   // [n_structs]
};

// Pair table of the structure store (fd_store_build_pair_table): for every stored structure the residue pairs that
// get a geometric hash -- exactly the pairs the index build emitted for it (K1) -- sorted by hash:
//   hash[offsets[s] .. offsets[s + 1])  ascending,   ij[...] = i << 16 | j of the same entry,
//   dir[s * FD_PT_DIR + p] = first entry of structure s whose hash >> 20 (the amino-acid pair) is >= p.
// This is the (structure, res_i, res_j) array of the task's north star: with it the verification looks a query hash
// up (one directory read + a short binary search) instead of re-screening and re-hashing n1 x n2 residue pairs of
// every candidate (retrieve.rs:52-156 recomputes them per candidate).  8 B per hashed pair, ~115 pairs per residue.
constexpr uint32_t FD_PT_DIR = 1025; // prefixes 0 .. 1023 and the end sentinel
constexpr uint32_t FD_AA_DIR = 42;   // 41 amino-acid buckets of a structure's residues and the end sentinel
struct FdPairTable {
    bool built = false;
    uint64_t n = 0;
    uint32_t nbin_dist = 0, nbin_angle = 0;
    float dist_cutoff = 0.f;
    uint64_t *offsets = nullptr; // [n_structs + 1]
    uint32_t *hash = nullptr;    // [n]
    uint32_t *ij = nullptr;      // [n]
    uint32_t *dir = nullptr;     // [n_structs * FD_PT_DIR] relative to offsets[s]
};

// Device copy of the compact-structure store used by candidate verification.
struct FdDeviceStore {
    bool attached = false;
    uint64_t n_structs = 0, n_res = 0;
    uint64_t *row_offsets = nullptr;
    float *n_xyz = nullptr, *ca_xyz = nullptr, *cb_xyz = nullptr;
    uint8_t *aa = nullptr, *cb_valid = nullptr;
    std::vector<uint64_t> h_row_offsets; // host copy (tile planning of the pair-table build)
    // residues of every structure grouped by amino acid (built at attach, fd_edges.cu: k_store_aa_rows):
    //   aa_rows[row_offsets[s] + k]  residue indices of structure s, ordered by bucket, ascending inside a bucket;
    //   aa_dir[s * FD_AA_DIR + b]    first k of bucket b (b = 0..19 canonical amino acid, 20..39 modified residue of
    //                                that code, 40 = cannot pair: unknown amino acid or no CB), [.. + 41] = n.
    // The rescue vote of the verification (retrieve.rs:453-516) visits the residues of ONE amino acid per unmapped
    // query residue; with the directory it reads those ~n/20 rows instead of scanning all n.
    uint16_t *aa_rows = nullptr;
    uint16_t *aa_dir = nullptr;
    // (chain, residue number) of every residue (fd_store_attach_labels; optional): the device writes result rows
    uint8_t *label_chain = nullptr;
    uint64_t *label_serial = nullptr;
    FdPairTable pt;
};

// Pinned host staging buffer (grow-only) for results that the host side consumes right after the call.
struct FdPinned {
    void *p = nullptr;
    size_t cap = 0;
};

// Match records of the last fd_verify_candidates_device call, kept on the device for fd_verify_rows (fd_verify.cu)
struct FdVerifyKeep {
    void *recs = nullptr; // fd_match_record[n], grouped by candidate in emission order
    uint64_t n = 0, cap = 0, n_cand = 0;
    uint32_t *d_cand_query = nullptr; // [n_cand] index into the prepared batch
    const void *prepared = nullptr;   // fd_verify_prepared of the call
};

struct FdComm; // fd_comm.cu: the rank's NCCL communicator

struct fd_ctx {
    FdComm *comm = nullptr;        // multi-GPU: set by fd_comm_init, released by fd_comm_destroy / fd_destroy
    bool borrowed = false;         // fd_fork child: idx / store point into the parent's device memory
    std::vector<fd_ctx *> lanes;   // children kept by the in-library host for overlapped call sequences
    int device = 0;
    int num_sms = FD_NUM_SMS_FALLBACK;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    uint64_t launches = 0;
    std::map<std::string, FdStage> stages;
    FdDeviceIndex idx;
    FdDeviceStore store;
    uint64_t last_posting_bytes = 0;
    uint64_t last_exchange_bytes = 0; // bytes this rank sent to other ranks in the last sharded count_query
    FdPinned pinned[10];
    cudaEvent_t ev_extra[4] = {nullptr, nullptr, nullptr, nullptr}; // finer stage timing inside one call
    cudaStream_t aux_stream[2] = {nullptr, nullptr}; // second compute stream + copy stream of the chunked verification
    std::vector<cudaEvent_t> ev_pool;                // events of the chunked verification, created on demand
    double verify_edges_per_cand = 0.0, verify_comps_per_cand = 0.0; // pool sizing of the chunked verification
    FdVerifyKeep vkeep;
    uint64_t idx_generation = 0;            // bumped by every fd_index_attach
    void *cq_cache = nullptr;               // fd_query.cu: the last batch that carried an id, with its lookup results
    void (*cq_cache_free)(void *) = nullptr;
    uint32_t *votes = nullptr; // dense partial-vote planes of the last fd_votes_scan (device, owned)
    uint64_t votes_cap = 0;    // capacity in u32 words
    uint32_t *merge = nullptr; // dense vote planes of this rank's slice of the batch (sparse merge), owned
    uint64_t merge_cap = 0;
};

extern thread_local std::string fd_g_create_error;
// pinned staging buffer `slot` of at least `bytes` bytes (contents not preserved when it grows)
int fd_pinned(fd_ctx *ctx, int slot, size_t bytes, void **out);
cudaError_t fd_ensure_events(fd_ctx *ctx);
// adds a fork's stage times and launch count to its parent's and clears them (in-library host, after joining a lane)
void fd_fold_stats(fd_ctx *parent, fd_ctx *child);

inline int fd_fail(fd_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    else fd_g_create_error = msg;
    return code;
}

#define FD_CUDA(ctx, call)                                                                             \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            char b__[512];                                                                             \
            snprintf(b__, sizeof(b__), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return fd_fail((ctx), e__ == cudaErrorMemoryAllocation ? FD_ERR_NOMEM : FD_ERR_CUDA, b__); \
        }                                                                                              \
    } while (0)

#define FD_TRY(expr)              \
    do {                          \
        int r__ = (expr);         \
        if (r__ != FD_OK) return r__; \
    } while (0)

// Stream of the API call currently executing on this host thread (set by FD_ENTER).  Temporaries are
// allocated and freed stream-ordered from the device's default memory pool, whose release threshold is
// raised at fd_create so that freed blocks stay cached: a call costs no cudaMalloc/cudaFree round trips.
extern thread_local cudaStream_t fd_tls_stream;

// Owning device buffer for per-call temporaries; freed (stream-ordered) on scope exit.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFreeAsync(p, fd_tls_stream);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) count = 1;
        return cudaMallocAsync((void **)&p, count * sizeof(T), fd_tls_stream);
    }
    T *take() { // ownership passes to the caller, who frees it with cudaFree (after a stream sync)
        T *q = p;
        p = nullptr;
        n = 0;
        return q;
    }
};

#define FD_ENTER(ctx)                                 \
    do {                                              \
        FD_CUDA((ctx), cudaSetDevice((ctx)->device)); \
        fd_tls_stream = (ctx)->stream;                \
    } while (0)

// Brackets one stage with CUDA events on the library stream and accumulates its device time.
struct StageTimer {
    fd_ctx *ctx;
    const char *name;
    uint64_t launches0;
    StageTimer(fd_ctx *c, const char *n) : ctx(c), name(n), launches0(c->launches) {
        cudaEventRecord(ctx->ev0, ctx->stream);
    }
    // call after the stage's last launch; synchronises the stream
    cudaError_t finish() {
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaError_t e = cudaEventSynchronize(ctx->ev1);
        if (e != cudaSuccess) return e;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        FdStage &s = ctx->stages[name];
        s.ms += ms;
        s.launches += ctx->launches - launches0;
        return cudaGetLastError();
    }
};

// Accumulates host wall time of a scope into ctx->stages[name] (same table as the device stages).
struct HostTimer {
    fd_ctx *ctx;
    const char *name;
    std::chrono::steady_clock::time_point t0;
    HostTimer(fd_ctx *c, const char *n) : ctx(c), name(n), t0(std::chrono::steady_clock::now()) {}
    ~HostTimer() {
        FdStage &s = ctx->stages[name];
        s.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        s.launches += 1;
    }
};

#define FD_LAUNCH_ON(ctx, strm, kernel, grid, block, smem, ...)                 \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__);               \
        (ctx)->launches++;                                                      \
    } while (0)

#define FD_LAUNCH(ctx, kernel, grid, block, smem, ...)                          \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);        \
        (ctx)->launches++;                                                      \
    } while (0)

inline uint32_t fd_div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// fd_postings.cu: large copies between pageable host memory and the device through two pinned staging buffers
int fd_copy_to_host_staged(fd_ctx *ctx, void *dst, const void *d_src, size_t bytes);
int fd_copy_to_device_staged(fd_ctx *ctx, void *d_dst, const void *src, size_t bytes);

// fd_comm.cu
void fd_comm_release(fd_ctx *ctx);
int fd_comm_world_of(const fd_ctx *ctx);
int fd_comm_rank_of(const fd_ctx *ctx);
int fd_comm_alltoallv_dev(fd_ctx *ctx, const uint8_t *d_send, const uint64_t *send_off, const uint64_t *send_bytes,
                          uint8_t *d_recv, const uint64_t *recv_off, const uint64_t *recv_bytes);

// Shared between fd_hash.cu (index build) and fd_postings.cu
int fd_postings_from_keys(fd_ctx *ctx, uint64_t *d_keys /* hash<<32|id, consumed */, uint64_t n_keys,
                          uint64_t *d_tmp /* same size scratch */, fd_index_buffers *out);

// device batch upload helper (fd_hash.cu)
struct FdDeviceBatch {
    DevBuf<uint64_t> row_offsets;
    DevBuf<float> n_xyz, ca_xyz, cb_xyz;
    DevBuf<uint8_t> aa, cb_valid;
    uint64_t n_structs = 0, n_res = 0;
};
int fd_upload_batch(fd_ctx *ctx, const fd_struct_batch *b, FdDeviceBatch *d);
