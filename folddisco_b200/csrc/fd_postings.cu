// fd_postings.cu -- K2: (hash, id) pairs -> the reference's on-disk inverted index, byte for byte.
//
// Replaces FolddiscoIndex::count_single_entry / allocate_entries / add_single_entry / prune_to_sparse
// (reference src/index/indextable.rs:88-105, 171-237, 267-295).  The reference walks a dense 2^30-entry
// table twice (8 GB of counters, serial prefix sum, serial prune).  Here the multiset of pairs is sorted as
// 64-bit keys hash<<32|id (ids ascending inside each hash), and three streaming kernels produce the files:
//   k2_measure : per key, is it the head of a hash run?  LEB128 length of (id - previous id in the run)
//   (exclusive scans: byte offsets; compacted list index per head)
//   k2_encode  : writes the varint bytes at their final offset, and hashes[]/offsets[] at run heads
// Everything is HBM-bandwidth bound: 8 B read + ~1.1 B written per posting per pass.
#include <cub/cub.cuh>

#include "fd_common.cuh"

namespace {

__device__ __forceinline__ uint32_t varint_len(uint32_t v) {
    // 1 + ilog2(v)/7 for v > 0, 1 for v == 0   (indextable.rs:92-99)
    return v == 0 ? 1u : 1u + (31u - __clz(v)) / 7u;
}

__global__ void k2_measure(const uint64_t *keys, uint64_t n, uint32_t *byte_len, uint32_t *is_head) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t key = keys[k];
    const uint32_t h = (uint32_t)(key >> 32), id = (uint32_t)key;
    bool head = true;
    uint32_t delta = id;
    if (k > 0) {
        const uint64_t prev = keys[k - 1];
        if ((uint32_t)(prev >> 32) == h) {
            head = false;
            delta = id - (uint32_t)prev;
        }
    }
    byte_len[k] = varint_len(delta);
    is_head[k] = head ? 1u : 0u;
}

__global__ void k2_encode(const uint64_t *keys, uint64_t n, const uint64_t *byte_off, const uint64_t *list_idx,
                          const uint32_t *is_head, uint8_t *values, uint32_t *hashes, uint64_t *offsets) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t key = keys[k];
    const uint32_t h = (uint32_t)(key >> 32), id = (uint32_t)key;
    uint32_t delta = id;
    if (!is_head[k]) delta = id - (uint32_t)keys[k - 1];
    uint64_t o = byte_off[k];
    if (is_head[k]) {
        hashes[list_idx[k]] = h;
        offsets[list_idx[k]] = o;
    }
    // split_by_seven_bits (indextable.rs:397-418)
    while (delta >= 128u) {
        values[o++] = (uint8_t)((delta & 0x7Fu) | 0x80u);
        delta >>= 7;
    }
    values[o] = (uint8_t)delta;
}

__global__ void k2_widen(const uint32_t *in, uint64_t n, uint64_t *out) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[k];
}

__global__ void k2_make_keys(const uint32_t *hashes, const uint64_t *row_offsets, uint64_t n_structs,
                             uint64_t first_id, uint64_t *keys) {
    // one warp per structure row
    const uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_structs) return;
    for (uint64_t k = row_offsets[s] + lane; k < row_offsets[s + 1]; k += 32)
        keys[k] = ((uint64_t)hashes[k] << 32) | (first_id + s);
}

} // namespace


// Device -> pageable host copy through two pinned staging buffers: chunk i + 1 crosses PCIe while the host threads
// copy chunk i to its destination (a plain cudaMemcpy into pageable memory runs at ~1.5 GB/s: the driver stages it
// through one small bounce buffer; the finished index is ~35 B per residue, 0.8 GB for the human proteome).
int fd_copy_to_host_staged(fd_ctx *ctx, void *dst, const void *d_src, size_t bytes) {
    if (bytes == 0) return FD_OK;
    constexpr size_t CHUNK = 32u << 20;
    if (bytes <= (1u << 20)) {
        FD_CUDA(ctx, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return FD_OK;
    }
    void *stage[2] = {nullptr, nullptr};
    FD_TRY(fd_pinned(ctx, 4, CHUNK, &stage[0]));
    FD_TRY(fd_pinned(ctx, 5, CHUNK, &stage[1]));
    FD_CUDA(ctx, fd_ensure_events(ctx));
    const size_t n_chunks = (bytes + CHUNK - 1) / CHUNK;
    auto issue = [&](size_t c) -> cudaError_t {
        const size_t off = c * CHUNK, n = std::min(CHUNK, bytes - off);
        cudaError_t e = cudaMemcpyAsync(stage[c & 1], (const uint8_t *)d_src + off, n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(ctx->ev_extra[c & 1], ctx->stream);
    };
    FD_CUDA(ctx, issue(0));
    for (size_t c = 0; c < n_chunks; c++) {
        FD_CUDA(ctx, cudaEventSynchronize(ctx->ev_extra[c & 1]));
        if (c + 1 < n_chunks) FD_CUDA(ctx, issue(c + 1));
        const size_t off = c * CHUNK, n = std::min(CHUNK, bytes - off);
        const int nt = 4;
        const uint8_t *src = (const uint8_t *)stage[c & 1];
        uint8_t *out = (uint8_t *)dst + off;
        fd_parallel(nt, [&](int t) {
            const size_t b0 = n * t / nt, b1 = n * (t + 1) / nt;
            memcpy(out + b0, src + b0, b1 - b0);
        });
    }
    return FD_OK;
}

// Pageable host -> device through the same two pinned buffers (host threads fill chunk i + 1 while chunk i crosses
// PCIe): the attach of an index whose payload sits in the caller's (mmap'ed or malloc'ed) memory.
int fd_copy_to_device_staged(fd_ctx *ctx, void *d_dst, const void *src, size_t bytes) {
    if (bytes == 0) return FD_OK;
    constexpr size_t CHUNK = 32u << 20;
    if (bytes <= (1u << 20)) {
        FD_CUDA(ctx, cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return FD_OK;
    }
    void *stage[2] = {nullptr, nullptr};
    FD_TRY(fd_pinned(ctx, 4, CHUNK, &stage[0]));
    FD_TRY(fd_pinned(ctx, 5, CHUNK, &stage[1]));
    FD_CUDA(ctx, fd_ensure_events(ctx));
    const size_t n_chunks = (bytes + CHUNK - 1) / CHUNK;
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t off = c * CHUNK, n = std::min(CHUNK, bytes - off);
        if (c >= 2) FD_CUDA(ctx, cudaEventSynchronize(ctx->ev_extra[c & 1])); // the copy that last read this buffer
        const int nt = 4;
        const uint8_t *in = (const uint8_t *)src + off;
        uint8_t *st = (uint8_t *)stage[c & 1];
        fd_parallel(nt, [&](int t) {
            const size_t b0 = n * t / nt, b1 = n * (t + 1) / nt;
            memcpy(st + b0, in + b0, b1 - b0);
        });
        FD_CUDA(ctx, cudaMemcpyAsync((uint8_t *)d_dst + off, st, n, cudaMemcpyHostToDevice, ctx->stream));
        FD_CUDA(ctx, cudaEventRecord(ctx->ev_extra[c & 1], ctx->stream));
    }
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

// keys: hash<<32|id, any order, duplicates allowed (they are the per-structure duplicates).
int fd_postings_from_keys(fd_ctx *ctx, uint64_t *d_keys, uint64_t n_keys, uint64_t *d_tmp, fd_index_buffers *out) {
    memset(out, 0, sizeof(*out));
    if (n_keys == 0) {
        out->hashes = (uint32_t *)malloc(4);
        out->offsets = (uint64_t *)calloc(1, 8);
        out->values = (uint8_t *)malloc(1);
        if (!out->hashes || !out->offsets || !out->values) return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
        return FD_OK;
    }
    StageTimer st(ctx, "postings");
    cudaStream_t s = ctx->stream;
    DevBuf<uint8_t> tmp;
    DevBuf<uint64_t> d_nuniq;
    FD_CUDA(ctx, d_nuniq.alloc(1));
    size_t tb_sort = 0, tb_uniq = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tb_sort, d_keys, d_tmp, n_keys, 0, 64, s);
    cub::DeviceSelect::Unique(nullptr, tb_uniq, d_tmp, d_keys, d_nuniq.p, n_keys, s);
    FD_CUDA(ctx, tmp.alloc(std::max(tb_sort, tb_uniq)));
    size_t tb = tb_sort;
    FD_CUDA(ctx, cub::DeviceRadixSort::SortKeys(tmp.p, tb, d_keys, d_tmp, n_keys, 0, 64, s));
    ctx->launches += 8;
    tb = tb_uniq;
    FD_CUDA(ctx, cub::DeviceSelect::Unique(tmp.p, tb, d_tmp, d_keys, d_nuniq.p, n_keys, s)); // d_keys := unique sorted
    ctx->launches += 2;
    uint64_t n = 0;
    FD_CUDA(ctx, cudaMemcpyAsync(&n, d_nuniq.p, 8, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, cudaStreamSynchronize(s));

    DevBuf<uint32_t> byte_len, is_head;
    DevBuf<uint64_t> byte_off, list_idx;
    FD_CUDA(ctx, byte_len.alloc(n));
    FD_CUDA(ctx, is_head.alloc(n));
    FD_CUDA(ctx, byte_off.alloc(n + 1));
    FD_CUDA(ctx, list_idx.alloc(n + 1));
    const uint32_t grid = fd_div_up(n, 256);
    FD_LAUNCH(ctx, k2_measure, grid, 256, 0, d_keys, n, byte_len.p, is_head.p);
    size_t tb_scan = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb_scan, byte_off.p, byte_off.p, n + 1, s);
    if (tb_scan > tmp.n) FD_CUDA(ctx, tmp.alloc(tb_scan));
    // exclusive sums over n+1 elements (the last input is a zero pad) give the totals in slot n
    DevBuf<uint64_t> widep;
    FD_CUDA(ctx, widep.alloc(n + 1));
    FD_CUDA(ctx, cudaMemsetAsync(widep.p + n, 0, 8, s));
    FD_LAUNCH(ctx, k2_widen, grid, 256, 0, byte_len.p, n, widep.p);
    tb = tb_scan;
    FD_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb, widep.p, byte_off.p, n + 1, s));
    ctx->launches += 2;
    FD_LAUNCH(ctx, k2_widen, grid, 256, 0, is_head.p, n, widep.p);
    tb = tb_scan;
    FD_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb, widep.p, list_idx.p, n + 1, s));
    ctx->launches += 2;
    uint64_t totals[2] = {0, 0};
    FD_CUDA(ctx, cudaMemcpyAsync(&totals[0], byte_off.p + n, 8, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, cudaMemcpyAsync(&totals[1], list_idx.p + n, 8, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, cudaStreamSynchronize(s));
    const uint64_t value_bytes = totals[0], count = totals[1];

    DevBuf<uint8_t> d_values;
    DevBuf<uint32_t> d_hashes;
    DevBuf<uint64_t> d_offsets;
    FD_CUDA(ctx, d_values.alloc(value_bytes));
    FD_CUDA(ctx, d_hashes.alloc(count));
    FD_CUDA(ctx, d_offsets.alloc(count + 1));
    FD_LAUNCH(ctx, k2_encode, grid, 256, 0, d_keys, n, byte_off.p, list_idx.p, is_head.p, d_values.p, d_hashes.p,
              d_offsets.p);
    FD_CUDA(ctx, cudaMemcpyAsync(d_offsets.p + count, byte_off.p + n, 8, cudaMemcpyDeviceToDevice, s));

    out->count = count;
    out->value_bytes = value_bytes;
    out->hashes = (uint32_t *)malloc(std::max<uint64_t>(count, 1) * 4);
    out->offsets = (uint64_t *)malloc((count + 1) * 8);
    out->values = (uint8_t *)malloc(std::max<uint64_t>(value_bytes, 1));
    if (!out->hashes || !out->offsets || !out->values) {
        fd_free_index_buffers(out);
        return fd_fail(ctx, FD_ERR_NOMEM, "host allocation failed");
    }
    FD_CUDA(ctx, st.finish()); // "postings" = sort + encode kernels; the copy of the finished index is timed on its own
    {
        HostTimer stc(ctx, "postings_copy");
        FD_TRY(fd_copy_to_host_staged(ctx, out->hashes, d_hashes.p, count * 4));
        FD_TRY(fd_copy_to_host_staged(ctx, out->offsets, d_offsets.p, (count + 1) * 8));
        FD_TRY(fd_copy_to_host_staged(ctx, out->values, d_values.p, value_bytes));
    }
    return FD_OK;
}

extern "C" {

int fd_build_postings(fd_ctx *ctx, const uint32_t *hashes, const uint64_t *row_offsets, uint64_t n_structs,
                      uint64_t first_id, fd_index_buffers *out) {
    if (!ctx) return FD_ERR_ARG;
    if (!row_offsets || !out || (n_structs && row_offsets[n_structs] && !hashes))
        return fd_fail(ctx, FD_ERR_ARG, "fd_build_postings: NULL argument");
    if (first_id + n_structs > 0xffffffffull) return fd_fail(ctx, FD_ERR_LIMIT, "structure ids must fit in 32 bits");
    FD_ENTER(ctx);
    const uint64_t n = row_offsets[n_structs];
    DevBuf<uint32_t> d_h;
    DevBuf<uint64_t> d_ro, keys, tmp;
    FD_CUDA(ctx, d_h.alloc(n));
    FD_CUDA(ctx, d_ro.alloc(n_structs + 1));
    FD_CUDA(ctx, keys.alloc(n));
    FD_CUDA(ctx, tmp.alloc(n));
    FD_CUDA(ctx, cudaMemcpyAsync(d_h.p, hashes, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    FD_CUDA(ctx, cudaMemcpyAsync(d_ro.p, row_offsets, (n_structs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (n_structs)
        FD_LAUNCH(ctx, k2_make_keys, fd_div_up(n_structs * 32, 256), 256, 0, d_h.p, d_ro.p, n_structs, first_id,
                  keys.p);
    FD_CUDA(ctx, cudaGetLastError());
    return fd_postings_from_keys(ctx, keys.p, n, tmp.p, out);
}

void fd_free_index_buffers(fd_index_buffers *b) {
    if (!b) return;
    free(b->hashes);
    free(b->offsets);
    free(b->values);
    memset(b, 0, sizeof(*b));
}

} // extern "C"
