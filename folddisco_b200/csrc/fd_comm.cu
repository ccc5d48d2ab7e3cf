// fd_comm.cu -- NCCL inside the C ABI: one communicator per context (one context per rank / GPU), so that a host in
// any language can run the multi-GPU search through the library alone (a Rust host cannot use torch.distributed).
//
// The reference is a single-process program (rayon; no MPI / NCCL anywhere -- SURVEY section 2); the exchange step
// exists only here.  The id-range partition (SURVEY 8e ablation) needs three collectives per batch:
//   all-gather  of the ranks' flattened query descriptors (every rank scans its shard for the WHOLE batch),
//   all-reduce  (sum, u32) of the local posting counts -> global list lengths -> the reference's idf weights,
//   all-to-all  of fixed-size per-query top-n blocks to the rank that owns the query (grouped ncclSend / ncclRecv).
// All of them run on the context's stream; host buffers are staged through device memory.
#include <dlfcn.h>
#include <nccl.h>

#include "fd_common.cuh"

// NCCL is bound at run time (dlopen of libnccl.so.2 on the first fd_comm_* call) instead of at link time: a host
// process may already carry its own NCCL under the same soname (PyTorch bundles a newer one than the system's), and
// two different libnccl.so.2 cannot live in one process.  dlopen returns the copy that is already loaded, else the
// system's.  The entry points used are stable across NCCL 2.x.
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    std::string error;
};
NcclApi &nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) {
            a.error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return a;
        }
        bool ok = true;
        auto sym = [&](const char *name) {
            void *p = dlsym(a.handle, name);
            if (!p) {
                ok = false;
                a.error = std::string("libnccl.so.2 lacks ") + name;
            }
            return p;
        };
        a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
        a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
        a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
        a.Send = (decltype(a.Send))sym("ncclSend");
        a.Recv = (decltype(a.Recv))sym("ncclRecv");
        a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
        a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
        if (!ok) a.handle = nullptr;
        return a;
    }();
    return api;
}
} // namespace
#define ncclGetUniqueId nccl_api().GetUniqueId
#define ncclCommInitRank nccl_api().CommInitRank
#define ncclCommDestroy nccl_api().CommDestroy
#define ncclGetErrorString nccl_api().GetErrorString
#define ncclGroupStart nccl_api().GroupStart
#define ncclGroupEnd nccl_api().GroupEnd
#define ncclSend nccl_api().Send
#define ncclRecv nccl_api().Recv
#define ncclAllGather nccl_api().AllGather
#define ncclAllReduce nccl_api().AllReduce

struct FdComm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

static int nccl_fail(fd_ctx *ctx, ncclResult_t r, const char *what) {
    return fd_fail(ctx, FD_ERR_CUDA, std::string(what) + ": " + ncclGetErrorString(r));
}
#define FD_NCCL(ctx, call)                                       \
    do {                                                         \
        ncclResult_t r__ = (call);                               \
        if (r__ != ncclSuccess) return nccl_fail((ctx), r__, #call); \
    } while (0)

void fd_comm_release(fd_ctx *ctx) {
    if (ctx->comm) {
        if (ctx->comm->comm) ncclCommDestroy(ctx->comm->comm);
        delete ctx->comm;
        ctx->comm = nullptr;
    }
}

// device-level helpers for the other translation units (fd_query.cu)
int fd_comm_world_of(const fd_ctx *ctx) { return ctx->comm ? ctx->comm->world : 1; }
int fd_comm_rank_of(const fd_ctx *ctx) { return ctx->comm ? ctx->comm->rank : 0; }

// Every rank sends block r of `send` (blocks of block_bytes[r] bytes at send_off[r]) to rank r and receives rank r's
// block for itself at recv_off[r] (recv_bytes[r] bytes): one grouped send / recv = an all-to-all with per-peer sizes.
int fd_comm_alltoallv_dev(fd_ctx *ctx, const uint8_t *d_send, const uint64_t *send_off, const uint64_t *send_bytes,
                          uint8_t *d_recv, const uint64_t *recv_off, const uint64_t *recv_bytes) {
    if (!ctx->comm) return fd_fail(ctx, FD_ERR_STATE, "no communicator: call fd_comm_init first");
    FdComm &c = *ctx->comm;
    FD_NCCL(ctx, ncclGroupStart());
    for (int r = 0; r < c.world; r++) {
        if (send_bytes[r]) FD_NCCL(ctx, ncclSend(d_send + send_off[r], send_bytes[r], ncclUint8, r, c.comm, ctx->stream));
        if (recv_bytes[r]) FD_NCCL(ctx, ncclRecv(d_recv + recv_off[r], recv_bytes[r], ncclUint8, r, c.comm, ctx->stream));
    }
    FD_NCCL(ctx, ncclGroupEnd());
    return FD_OK;
}

extern "C" {

int fd_comm_unique_id(uint8_t *out_id) {
    if (!out_id) return FD_ERR_ARG;
    static_assert(sizeof(ncclUniqueId) == FD_COMM_ID_BYTES, "ncclUniqueId size");
    if (!nccl_api().handle) return FD_ERR_STATE;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return FD_ERR_CUDA;
    memcpy(out_id, &id, sizeof(id));
    return FD_OK;
}

int fd_comm_init(fd_ctx *ctx, const uint8_t *id_bytes, int rank, int world) {
    if (!ctx) return FD_ERR_ARG;
    if (!id_bytes || world < 1 || rank < 0 || rank >= world) return fd_fail(ctx, FD_ERR_ARG, "fd_comm_init: bad argument");
    if (!nccl_api().handle) return fd_fail(ctx, FD_ERR_STATE, "fd_comm_init: " + nccl_api().error);
    FD_ENTER(ctx);
    fd_comm_release(ctx);
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    FdComm *c = new FdComm();
    c->rank = rank;
    c->world = world;
    ncclResult_t r = ncclCommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return nccl_fail(ctx, r, "ncclCommInitRank");
    }
    ctx->comm = c;
    return FD_OK;
}

int fd_comm_rank(const fd_ctx *ctx) { return ctx ? fd_comm_rank_of(ctx) : 0; }
int fd_comm_world(const fd_ctx *ctx) { return ctx ? fd_comm_world_of(ctx) : 1; }

void fd_comm_destroy(fd_ctx *ctx) {
    if (ctx) fd_comm_release(ctx);
}

int fd_comm_allgather(fd_ctx *ctx, const void *send, uint64_t bytes, void *recv) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->comm) return fd_fail(ctx, FD_ERR_STATE, "no communicator: call fd_comm_init first");
    if (bytes && (!send || !recv)) return fd_fail(ctx, FD_ERR_ARG, "fd_comm_allgather: NULL buffer");
    FD_ENTER(ctx);
    if (bytes == 0) return FD_OK;
    FdComm &c = *ctx->comm;
    DevBuf<uint8_t> d_send, d_recv;
    FD_CUDA(ctx, d_send.alloc(bytes));
    FD_CUDA(ctx, d_recv.alloc(bytes * c.world));
    FD_CUDA(ctx, cudaMemcpyAsync(d_send.p, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
    FD_NCCL(ctx, ncclAllGather(d_send.p, d_recv.p, bytes, ncclUint8, c.comm, ctx->stream));
    FD_CUDA(ctx, cudaMemcpyAsync(recv, d_recv.p, bytes * c.world, cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

int fd_comm_allreduce_u32(fd_ctx *ctx, uint32_t *inout, uint64_t n) {
    if (!ctx) return FD_ERR_ARG;
    if (!ctx->comm) return fd_fail(ctx, FD_ERR_STATE, "no communicator: call fd_comm_init first");
    if (n && !inout) return fd_fail(ctx, FD_ERR_ARG, "fd_comm_allreduce_u32: NULL buffer");
    FD_ENTER(ctx);
    if (n == 0) return FD_OK;
    FdComm &c = *ctx->comm;
    DevBuf<uint32_t> d;
    FD_CUDA(ctx, d.alloc(n));
    FD_CUDA(ctx, cudaMemcpyAsync(d.p, inout, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    FD_NCCL(ctx, ncclAllReduce(d.p, d.p, n, ncclUint32, ncclSum, c.comm, ctx->stream));
    FD_CUDA(ctx, cudaMemcpyAsync(inout, d.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

int fd_comm_barrier(fd_ctx *ctx) {
    uint32_t one = 1;
    return fd_comm_allreduce_u32(ctx, &one, 1);
}

} // extern "C"
