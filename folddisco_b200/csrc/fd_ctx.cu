// fd_ctx.cu -- lifecycle of the C ABI (include/folddisco_b200.h) and the math parity probes.
#include <algorithm>
#include <thread>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <vector>
#include "fd_common.cuh"
#include "fd_geom.cuh"
#include "fd_hashtypes.cuh"

thread_local std::string fd_g_create_error;
thread_local cudaStream_t fd_tls_stream = nullptr;

static void fd_release_index(FdDeviceIndex &ix) {
    cudaFree(ix.hashes);
    cudaFree(ix.offsets);
    cudaFree(ix.values);
    cudaFree(ix.counts);
    cudaFree(ix.dir);
    cudaFree(ix.skip_id);
    cudaFree(ix.skip_off);
    cudaFree(ix.nres);
    cudaFree(ix.plddt);
    ix = FdDeviceIndex();
}
static void fd_release_store(FdDeviceStore &st) {
    cudaFree(st.row_offsets);
    cudaFree(st.n_xyz);
    cudaFree(st.ca_xyz);
    cudaFree(st.cb_xyz);
    cudaFree(st.aa);
    cudaFree(st.cb_valid);
    cudaFree(st.aa_rows);
    cudaFree(st.aa_dir);
    cudaFree(st.label_chain);
    cudaFree(st.label_serial);
    cudaFree(st.pt.offsets);
    cudaFree(st.pt.hash);
    cudaFree(st.pt.ij);
    cudaFree(st.pt.dir);
    st = FdDeviceStore();
}
int fd_pinned(fd_ctx *ctx, int slot, size_t bytes, void **out) {
    FdPinned &b = ctx->pinned[slot];
    if (bytes > b.cap) {
        if (b.p) cudaFreeHost(b.p);
        b.p = nullptr;
        b.cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&b.p, want, cudaHostAllocDefault);
        if (e != cudaSuccess) return fd_fail(ctx, FD_ERR_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
        b.cap = want;
    }
    *out = b.p;
    return FD_OK;
}
void fd_fold_stats(fd_ctx *parent, fd_ctx *child) {
    for (auto &kv : child->stages) {
        FdStage &s = parent->stages[kv.first];
        s.ms += kv.second.ms;
        s.launches += kv.second.launches;
    }
    child->stages.clear();
    parent->launches += child->launches;
    child->launches = 0;
}

cudaError_t fd_ensure_events(fd_ctx *ctx) {
    for (auto &e : ctx->ev_extra)
        if (!e) {
            cudaError_t r = cudaEventCreate(&e);
            if (r != cudaSuccess) return r;
        }
    return cudaSuccess;
}
void fd_ctx_release_index(fd_ctx *ctx) { fd_release_index(ctx->idx); }
void fd_ctx_release_store(fd_ctx *ctx) { fd_release_store(ctx->store); }

__global__ void fd_math_probe_kernel(int op, const float *a, const float *b, uint64_t n, float *out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s, c;
    switch (op) {
        case 0:
            fdm::sincosf_exact(a[i], &s, &c);
            out[i] = s;
            break;
        case 1:
            fdm::sincosf_exact(a[i], &s, &c);
            out[i] = c;
            break;
        case 2: out[i] = fdm::acosf_exact(a[i]); break;
        default: out[i] = fdm::atan2f_exact(a[i], b[i]); break;
    }
}


// ------------------------------------------------------------------------------------------------
// Host worker pool.  The host side above the kernels runs many short parallel regions per batch (query maps,
// verification tables, row assembly); spawning std::threads for each costs more than the regions themselves
// (~20 us per thread).  fd_parallel(nt, fn) runs fn(0) .. fn(nt - 1) concurrently -- fn(0) on the caller -- on
// persistent workers that sleep on a condition variable between regions; concurrent regions share the workers.
// ------------------------------------------------------------------------------------------------
namespace {
struct PoolRegion {
    const std::function<void(int)> *fn;
    int nt;
    int next = 0; // next unclaimed worker index (under HostPool::m)
    int done = 0; // finished indices (under HostPool::m)
};
struct HostPool {
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    std::vector<PoolRegion *> active; // regions with unclaimed indices
    int n_workers = 0;

    // claims an index of some active region; nullptr if there is none (caller holds m)
    PoolRegion *claim(int *idx) {
        while (!active.empty()) {
            PoolRegion *r = active.back();
            if (r->next < r->nt) {
                *idx = r->next++;
                if (r->next == r->nt) active.pop_back();
                return r;
            }
            active.pop_back();
        }
        return nullptr;
    }
    void finish(PoolRegion *r) { // caller holds m
        if (++r->done == r->nt) cv_done.notify_all();
    }
    void worker() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            int idx = 0;
            PoolRegion *r = claim(&idx);
            if (!r) {
                cv_work.wait(lk);
                continue;
            }
            lk.unlock();
            (*r->fn)(idx);
            lk.lock();
            finish(r);
        }
    }
    void grow(int want) { // caller holds m
        while (n_workers < want) {
            std::thread([this] { worker(); }).detach();
            n_workers++;
        }
    }
};
HostPool *host_pool() {
    static HostPool *p = new HostPool(); // leaked on purpose: detached workers may outlive static destructors
    return p;
}
} // namespace

// Regions from several host threads (search lanes, a prepare thread next to a search thread) share the workers: a
// worker takes the next unclaimed index of any active region; the caller works on its own region too and then waits
// for the indices that others took.
void fd_parallel(int nt, const std::function<void(int)> &fn) {
    if (nt <= 1) {
        fn(0);
        return;
    }
    HostPool *P = host_pool();
    PoolRegion r{&fn, nt};
    std::unique_lock<std::mutex> lk(P->m);
    P->grow(std::min(nt - 1, 255));
    P->active.push_back(&r);
    P->cv_work.notify_all();
    for (;;) { // the caller takes indices of its own region only (it must return when the region is done)
        if (r.next >= r.nt) break;
        const int idx = r.next++;
        if (r.next == r.nt)
            for (size_t k = 0; k < P->active.size(); k++)
                if (P->active[k] == &r) {
                    P->active.erase(P->active.begin() + (long)k);
                    break;
                }
        lk.unlock();
        fn(idx);
        lk.lock();
        P->finish(&r);
    }
    P->cv_done.wait(lk, [&] { return r.done == r.nt; });
}

extern "C" {

int fd_default_host_threads(void) {
    if (const char *e = getenv("FD_HOST_THREADS")) {
        const int n = atoi(e);
        if (n > 0) return std::min(n, 256);
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

// parity probe (host build of fd_geom.cuh, no device needed): for residue pairs (i, j) of one SoA structure the exact
// hash, the fast-route hash (0 when the fast route declined) and whether it declined
void fd_pair_hash_host(const float *n_xyz, const float *ca_xyz, const float *cb_xyz, const uint8_t *aa,
                       const uint32_t *pi, const uint32_t *pj, uint64_t n_pairs, const fd_hash_params *params,
                       uint32_t *out_exact, uint32_t *out_fast, uint8_t *out_declined) {
    const fdg::HashParams hp = fdg::make_params(params->nbin_dist, params->nbin_angle, params->dist_cutoff);
    auto ld = [](const float *p, uint64_t r) { return fdg::V3{p[3 * r], p[3 * r + 1], p[3 * r + 2]}; };
    for (uint64_t k = 0; k < n_pairs; k++) {
        const uint64_t i = pi[k], j = pj[k];
        const float d = fdg::dist(ld(ca_xyz, i), ld(ca_xyz, j));
        out_exact[k] = fdg::pair_hash(ld(n_xyz, i), ld(ca_xyz, i), ld(cb_xyz, i), ld(n_xyz, j), ld(ca_xyz, j),
                                      ld(cb_xyz, j), aa[i], aa[j], d, hp);
        uint32_t h = 0;
        const bool ok = fdg::pair_hash_fast(ld(n_xyz, i), ld(ca_xyz, i), ld(cb_xyz, i), ld(n_xyz, j), ld(ca_xyz, j),
                                            ld(cb_xyz, j), aa[i], aa[j], d, hp, &h);
        out_fast[k] = ok ? h : 0u;
        out_declined[k] = ok ? 0 : 1;
    }
}

// parity probe of the other encodings: every ordered residue pair of one structure through fd_hashtypes.cuh on the host
int64_t fd_typed_hash_host(const float *n_xyz, const float *ca_xyz, const float *cb_xyz, const uint8_t *aa,
                           const uint8_t *cb_valid, uint64_t n_res, const fd_hash_params *params, uint32_t *out,
                           int64_t cap) {
    fdg::TypedParams tp;
    if (fdg::typed_params_from(params, &tp)) return -1;
    auto ld = [](const float *p, uint64_t r) { return fdg::V3{p[3 * r], p[3 * r + 1], p[3 * r + 2]}; };
    int64_t n = 0;
    float f[9];
    for (uint64_t i = 0; i < n_res; i++)
        for (uint64_t j = 0; j < n_res; j++) {
            if (i == j || aa[i] == 255 || aa[j] == 255) continue;
            if (fdg::ht_needs_cb(tp.type) && cb_valid && (!cb_valid[i] || !cb_valid[j])) continue;
            fdg::Nbr nb;
            if (fdg::ht_needs_neighbours(tp.type)) {
                if (i == 0 || j == 0 || i + 1 >= n_res || j + 1 >= n_res) continue;
                nb = fdg::Nbr{ld(ca_xyz, i - 1), ld(ca_xyz, i + 1), ld(ca_xyz, j - 1), ld(ca_xyz, j + 1), (float)j - (float)i};
            }
            const float d = fdg::typed_screen_dist(tp.type, ld(ca_xyz, i), ld(cb_xyz, i), ld(ca_xyz, j), ld(cb_xyz, j));
            if (d > tp.dist_cutoff) continue;
            fdg::typed_feature(tp.type, ld(n_xyz, i), ld(ca_xyz, i), ld(cb_xyz, i), ld(n_xyz, j), ld(ca_xyz, j),
                               ld(cb_xyz, j), (float)(aa[i] & 0x7Fu), (float)(aa[j] & 0x7Fu), d, f, &nb);
            for (uint32_t b = 0; b < tp.n_bins; b++) {
                const uint32_t h = fdg::typed_hash(tp.type, f, tp.nbd[b], tp.nba[b]);
                if (n < cap) out[n] = h;
                n++;
            }
        }
    return n;
}
int fd_typed_is_symmetric_host(uint32_t hash_type, uint32_t hash) {
    if (!fdg::ht_supported(hash_type)) return -1;
    return fdg::typed_is_symmetric(fdg::ht_canon(hash_type), hash) ? 1 : 0;
}

// parity / debug probe: runs a region of nt workers that each spin for `spin_us`; returns how many distinct host
// threads took part (1..nt: a fast thread may take several indices), or -1 if some index did not run exactly once
int fd_parallel_probe(int nt, int spin_us) {
    std::mutex m;
    std::vector<std::thread::id> ids;
    std::vector<int> seen((size_t)std::max(nt, 1), 0);
    fd_parallel(nt, [&](int idx) {
        const auto t0 = std::chrono::steady_clock::now();
        while (std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() < spin_us) {
        }
        std::lock_guard<std::mutex> lk(m);
        ids.push_back(std::this_thread::get_id());
        if (idx >= 0 && idx < (int)seen.size()) seen[(size_t)idx]++;
    });
    for (int k = 0; k < std::max(nt, 1); k++)
        if (seen[(size_t)k] != 1) return -1; // an index ran twice or not at all
    std::sort(ids.begin(), ids.end());
    return (int)(std::unique(ids.begin(), ids.end()) - ids.begin());
}

int fd_device(const fd_ctx *ctx) { return ctx ? ctx->device : -1; }

const char *fd_version(void) { return "folddisco_b200 0.1 (sm_100a)"; }

int fd_create(fd_ctx **out, int device) {
    if (!out) return fd_fail(nullptr, FD_ERR_ARG, "fd_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fd_fail(nullptr, FD_ERR_CUDA,
                       std::string("fd_create: no usable CUDA device (") + cudaGetErrorString(e) +
                           "); this library has no CPU fallback");
    if (device < 0 || device >= n) return fd_fail(nullptr, FD_ERR_ARG, "fd_create: device index out of range");
    fd_ctx *ctx = new fd_ctx();
    ctx->device = device;
    FD_CUDA(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    FD_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    FD_CUDA(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    FD_CUDA(nullptr, cudaEventCreate(&ctx->ev0));
    FD_CUDA(nullptr, cudaEventCreate(&ctx->ev1));
    { // keep freed temporaries cached in the default pool instead of returning them to the driver
        cudaMemPool_t pool;
        FD_CUDA(nullptr, cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thr = UINT64_MAX;
        FD_CUDA(nullptr, cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
    *out = ctx;
    return FD_OK;
}

int fd_fork(fd_ctx *parent, fd_ctx **out) {
    if (!parent || !out) return fd_fail(parent, FD_ERR_ARG, "fd_fork: NULL argument");
    *out = nullptr;
    fd_ctx *c = nullptr;
    const int rc = fd_create(&c, parent->device);
    if (rc != FD_OK) return fd_fail(parent, rc, fd_g_create_error);
    c->borrowed = true;
    c->idx = parent->idx;
    c->store = parent->store;
    *out = c;
    return FD_OK;
}

int fd_fork_refresh(fd_ctx *parent, fd_ctx *child) {
    if (!parent || !child || !child->borrowed) return fd_fail(parent, FD_ERR_ARG, "fd_fork_refresh: not a fork");
    child->idx = parent->idx;
    child->store = parent->store;
    return FD_OK;
}

int fd_lane(fd_ctx *ctx, int i, fd_ctx **out) {
    if (!ctx || !out || i < 0 || i >= 8) return fd_fail(ctx, FD_ERR_ARG, "fd_lane: bad argument");
    while ((int)ctx->lanes.size() <= i) {
        fd_ctx *child = nullptr;
        FD_TRY(fd_fork(ctx, &child));
        ctx->lanes.push_back(child);
    }
    FD_TRY(fd_fork_refresh(ctx, ctx->lanes[i]));
    *out = ctx->lanes[i];
    return FD_OK;
}

void fd_lanes_fold_stats(fd_ctx *ctx) {
    if (!ctx) return;
    for (fd_ctx *l : ctx->lanes) fd_fold_stats(ctx, l);
}

void fd_destroy(fd_ctx *ctx) {
    if (!ctx) return;
    for (fd_ctx *l : ctx->lanes) fd_destroy(l);
    ctx->lanes.clear();
    cudaSetDevice(ctx->device);
    fd_comm_release(ctx);
    if (!ctx->borrowed) {
        fd_release_index(ctx->idx);
        fd_release_store(ctx->store);
    }
    cudaFree(ctx->votes);
    cudaFree(ctx->merge);
    fd_tls_stream = ctx->stream; // the cached batch releases its device buffers stream-ordered
    if (ctx->cq_cache && ctx->cq_cache_free) ctx->cq_cache_free(ctx->cq_cache);
    cudaFree(ctx->vkeep.recs);
    cudaFree(ctx->vkeep.d_cand_query);
    for (auto &b : ctx->pinned)
        if (b.p) cudaFreeHost(b.p);
    for (auto &e : ctx->ev_extra)
        if (e) cudaEventDestroy(e);
    for (auto &e : ctx->ev_pool)
        if (e) cudaEventDestroy(e);
    for (auto &st : ctx->aux_stream)
        if (st) cudaStreamDestroy(st);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void fd_note_host_ms(fd_ctx *ctx, const char *stage, double ms) {
    if (!ctx || !stage) return;
    FdStage &s = ctx->stages[stage];
    s.ms += ms;
    s.launches += 1;
}

void fd_note_general_path(fd_ctx *ctx, uint64_t n) {
    if (ctx) ctx->stages["general_candidates"].launches += n;
}

const char *fd_last_error(const fd_ctx *ctx) { return ctx ? ctx->err.c_str() : fd_g_create_error.c_str(); }
void fd_free(void *p) { free(p); }
uint64_t fd_kernel_launches(const fd_ctx *ctx) { return ctx ? ctx->launches : 0; }
double fd_stage_ms(const fd_ctx *ctx, const char *stage) {
    if (!ctx || !stage) return -1.0;
    auto it = ctx->stages.find(stage);
    return it == ctx->stages.end() ? 0.0 : it->second.ms;
}
uint64_t fd_stage_launches(const fd_ctx *ctx, const char *stage) {
    if (!ctx || !stage) return 0;
    auto it = ctx->stages.find(stage);
    return it == ctx->stages.end() ? 0 : it->second.launches;
}
uint64_t fd_last_posting_bytes(const fd_ctx *ctx) { return ctx ? ctx->last_posting_bytes : 0; }
uint64_t fd_index_num_structs(const fd_ctx *ctx) { return ctx && ctx->idx.attached ? ctx->idx.n_structs : 0; }

int fd_math_probe(fd_ctx *ctx, int op, const float *a, const float *b, uint64_t n, float *out) {
    if (!ctx || !a || !out || (op == 3 && !b)) return fd_fail(ctx, FD_ERR_ARG, "fd_math_probe: bad argument");
    FD_ENTER(ctx);
    DevBuf<float> da, db, dout;
    FD_CUDA(ctx, da.alloc(n));
    FD_CUDA(ctx, db.alloc(n));
    FD_CUDA(ctx, dout.alloc(n));
    FD_CUDA(ctx, cudaMemcpyAsync(da.p, a, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (b) FD_CUDA(ctx, cudaMemcpyAsync(db.p, b, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (n) FD_LAUNCH(ctx, fd_math_probe_kernel, fd_div_up(n, 256), 256, 0, op, da.p, db.p, n, dout.p);
    FD_CUDA(ctx, cudaGetLastError());
    FD_CUDA(ctx, cudaMemcpyAsync(out, dout.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

void fd_math_host(int op, const float *a, const float *b, uint64_t n, float *out) {
    for (uint64_t i = 0; i < n; i++) {
        float s, c;
        switch (op) {
            case 0:
                fdm::sincosf_exact(a[i], &s, &c);
                out[i] = s;
                break;
            case 1:
                fdm::sincosf_exact(a[i], &s, &c);
                out[i] = c;
                break;
            case 2: out[i] = fdm::acosf_exact(a[i]); break;
            default: out[i] = fdm::atan2f_exact(a[i], b[i]); break;
        }
    }
}

} // extern "C"
