// fd_metrics.cu -- K7: similarity metrics of verified matches (TM-score, GDT-TS, GDT-HA, Chamfer, Hausdorff).
//
// Replaces StructureSimilarityMetrics::calculate_all as rmsd_with_calpha_and_rottran calls it for every match
// (reference src/controller/retrieve.rs:776-831, src/structure/metrics.rs:44-345).  One thread per match: a match is
// 2m points (CA, CB of the m matched residues), its superposition (U, t) is already known from K5 / K6c, so the work is
// n transforms and n^2 distances (n <= ~130) -- the batch dimension is the parallelism, as in K5.  Target coordinates
// are gathered from the attached structure store; nothing but five floats per match leaves HBM.
#include "fd_common.cuh"
#include "fd_metrics.cuh"

namespace {

__global__ void k7_metrics_store(const float *q_ca, const float *q_cb, const float *s_ca, const float *s_cb,
                                 const uint64_t *s_row_offsets, const uint32_t *nid, const uint32_t *pair_offsets,
                                 const uint32_t *pair_q, const uint32_t *pair_t, const float *U9, const float *t3,
                                 uint32_t n_align, float *out5) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_align) return;
    const uint32_t p0 = pair_offsets[a], p1 = pair_offsets[a + 1];
    fdmt::GatherF mov{s_ca, s_cb, pair_t + p0, s_row_offsets[nid[a]]};
    fdmt::GatherF ref{q_ca, q_cb, pair_q + p0, 0};
    fdmt::similarity_metrics(2 * (p1 - p0), ref, mov, U9 + 9 * (size_t)a, t3 + 3 * (size_t)a, out5 + 5 * (size_t)a);
}

} // namespace

extern "C" {

int fd_metrics_store_batch(fd_ctx *ctx, const float *q_ca_xyz, const float *q_cb_xyz, uint64_t n_q_res,
                           const uint32_t *align_nid, const uint32_t *pair_offsets, uint32_t n_align,
                           const uint32_t *pair_qres, const uint32_t *pair_tres, const float *U9, const float *t3,
                           float *out_metrics) {
    if (!ctx) return FD_ERR_ARG;
    if (n_align == 0) return FD_OK;
    if (!ctx->store.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_metrics_store_batch: no structure store attached");
    if (!q_ca_xyz || !q_cb_xyz || !align_nid || !pair_offsets || !pair_qres || !pair_tres || !U9 || !t3 || !out_metrics)
        return fd_fail(ctx, FD_ERR_ARG, "fd_metrics_store_batch: NULL argument");
    const uint64_t np = pair_offsets[n_align];
    const FdDeviceStore &S = ctx->store;
    for (uint32_t a = 0; a < n_align; a++) {
        if (align_nid[a] >= S.n_structs) return fd_fail(ctx, FD_ERR_ARG, "align_nid outside the store");
        if (pair_offsets[a + 1] < pair_offsets[a]) return fd_fail(ctx, FD_ERR_ARG, "pair_offsets not ascending");
        const uint64_t nres = S.h_row_offsets[align_nid[a] + 1] - S.h_row_offsets[align_nid[a]];
        for (uint32_t k = pair_offsets[a]; k < pair_offsets[a + 1]; k++)
            if (pair_tres[k] >= nres) return fd_fail(ctx, FD_ERR_ARG, "pair_tres outside its structure");
    }
    for (uint64_t k = 0; k < np; k++)
        if (pair_qres[k] >= n_q_res) return fd_fail(ctx, FD_ERR_ARG, "pair_qres out of range");
    FD_ENTER(ctx);
    DevBuf<float> d_qca, d_qcb, d_U, d_t, d_out;
    DevBuf<uint32_t> d_nid, d_off, d_pq, d_pt;
    FD_CUDA(ctx, d_qca.alloc(3 * n_q_res));
    FD_CUDA(ctx, d_qcb.alloc(3 * n_q_res));
    FD_CUDA(ctx, d_nid.alloc(n_align));
    FD_CUDA(ctx, d_off.alloc((size_t)n_align + 1));
    FD_CUDA(ctx, d_pq.alloc(np));
    FD_CUDA(ctx, d_pt.alloc(np));
    FD_CUDA(ctx, d_U.alloc(9 * (size_t)n_align));
    FD_CUDA(ctx, d_t.alloc(3 * (size_t)n_align));
    FD_CUDA(ctx, d_out.alloc(5 * (size_t)n_align));
    cudaStream_t s = ctx->stream;
    FD_CUDA(ctx, cudaMemcpyAsync(d_qca.p, q_ca_xyz, 12 * n_q_res, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_qcb.p, q_cb_xyz, 12 * n_q_res, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_nid.p, align_nid, 4 * (size_t)n_align, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_off.p, pair_offsets, 4 * ((size_t)n_align + 1), cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_pq.p, pair_qres, 4 * np, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_pt.p, pair_tres, 4 * np, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_U.p, U9, 36 * (size_t)n_align, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_t.p, t3, 12 * (size_t)n_align, cudaMemcpyHostToDevice, s));
    StageTimer st(ctx, "metrics");
    FD_LAUNCH(ctx, k7_metrics_store, fd_div_up(n_align, 128), 128, 0, d_qca.p, d_qcb.p, S.ca_xyz, S.cb_xyz, S.row_offsets,
              d_nid.p, d_off.p, d_pq.p, d_pt.p, d_U.p, d_t.p, n_align, d_out.p);
    FD_CUDA(ctx, cudaMemcpyAsync(out_metrics, d_out.p, 20 * (size_t)n_align, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, st.finish());
    return FD_OK;
}

// the same arithmetic on the host build of csrc/fd_metrics.cuh over explicit point lists (parity probe: no device)
void fd_metrics_host(const float *ref_xyz, const float *mov_xyz, uint32_t n_points, const float *U9, const float *t3,
                     float *out5) {
    fdmt::similarity_metrics(n_points, fdmt::FlatF{ref_xyz}, fdmt::FlatF{mov_xyz}, U9, t3, out5);
}

} // extern "C"
