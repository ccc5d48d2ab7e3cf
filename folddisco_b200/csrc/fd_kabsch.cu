// fd_kabsch.cu -- K5: batched Kabsch superposition + RMSD (query path (ii), verification tail).
//
// Replaces KabschSuperimposer::run -> kabsch(x = moving, y = reference, mode 2)
// (reference src/structure/kabsch.rs:73-95, 157-554; called from rmsd_with_calpha_and_rottran,
// src/controller/retrieve.rs:756-834) for a whole batch of alignments.  One thread per alignment: an alignment
// is 2m points (CA, CB of m <= ~10 matched residues), i.e. a few hundred bytes and ~2 kflop of binary64, so the
// batch dimension is the only parallelism worth having.  Same closed-form eigen decomposition of R^T R
// (TM-align) and the same operation order as the reference, binary64 inside, binary32 out; RMSD is computed
// from the explicitly transformed points like kabsch.rs:517-533.  Tolerance vs the reference: 1e-4 (f64 libm
// sin/cos/atan2 differ in the last ulp between CUDA and glibc).
#include "fd_common.cuh"
#include "fd_kabsch.cuh"
#include "fd_lmsqcp.cuh"

namespace {

using namespace fdk;

__global__ void k5_kabsch(const float *mov, const float *ref, const uint32_t *pt_offsets, uint32_t n_align,
                          float *rmsd, float *U9, float *t3) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_align) return;
    const uint32_t p0 = pt_offsets[a], p1 = pt_offsets[a + 1];
    kabsch_one(FlatPoints{mov + 3 * (size_t)p0}, FlatPoints{ref + 3 * (size_t)p0}, p1 - p0, U9 + 9 * (size_t)a,
               t3 + 3 * (size_t)a, rmsd + a);
}

// alignment a: target residues pair_t[k] of stored structure nid[a] (moving) onto query residues pair_q[k]
__global__ void k5_kabsch_store(const float *q_ca, const float *q_cb, const float *s_ca, const float *s_cb,
                                const uint64_t *s_row_offsets, const uint32_t *nid, const uint32_t *pair_offsets,
                                const uint32_t *pair_q, const uint32_t *pair_t, uint32_t n_align, float *rmsd,
                                float *U9, float *t3) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_align) return;
    const uint32_t p0 = pair_offsets[a], p1 = pair_offsets[a + 1];
    GatherPoints mov{s_ca, s_cb, pair_t + p0, s_row_offsets[nid[a]]};
    GatherPoints ref{q_ca, q_cb, pair_q + p0, 0};
    kabsch_one(mov, ref, 2 * (p1 - p0), U9 + 9 * (size_t)a, t3 + 3 * (size_t)a, rmsd + a);
}

// `--partial-fit` (retrieve.rs:773-814): more than three matched residues are superposed by LMS-QCP and report the RMSD of
// the inlier core, three or fewer by Kabsch like the default path
__global__ void k5_lmsqcp_store(const float *q_ca, const float *q_cb, const float *s_ca, const float *s_cb,
                                const uint64_t *s_row_offsets, const uint32_t *nid, const uint32_t *pair_offsets,
                                const uint32_t *pair_q, const uint32_t *pair_t, uint32_t n_align, float *rmsd,
                                float *U9, float *t3) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_align) return;
    const uint32_t p0 = pair_offsets[a], p1 = pair_offsets[a + 1];
    const uint64_t base = s_row_offsets[nid[a]];
    if (p1 - p0 <= 3) {
        GatherPoints mov{s_ca, s_cb, pair_t + p0, base};
        GatherPoints ref{q_ca, q_cb, pair_q + p0, 0};
        kabsch_one(mov, ref, 2 * (p1 - p0), U9 + 9 * (size_t)a, t3 + 3 * (size_t)a, rmsd + a);
        return;
    }
    fdmt::GatherF mov{s_ca, s_cb, pair_t + p0, base};
    fdmt::GatherF ref{q_ca, q_cb, pair_q + p0, 0};
    fdlq::lms_qcp_one(2 * (p1 - p0), ref, mov, U9 + 9 * (size_t)a, t3 + 3 * (size_t)a, rmsd + a);
}

} // namespace

extern "C" int fd_kabsch_batch(fd_ctx *ctx, const float *mov_xyz, const float *ref_xyz, const uint32_t *pt_offsets,
                               uint32_t n_align, float *rmsd, float *U9, float *t3) {
    if (!ctx) return FD_ERR_ARG;
    if (n_align == 0) return FD_OK;
    if (!mov_xyz || !ref_xyz || !pt_offsets || !rmsd || !U9 || !t3)
        return fd_fail(ctx, FD_ERR_ARG, "fd_kabsch_batch: NULL argument");
    FD_ENTER(ctx);
    const uint64_t npts = pt_offsets[n_align];
    DevBuf<float> d_mov, d_ref, d_rmsd, d_U, d_t;
    DevBuf<uint32_t> d_off;
    FD_CUDA(ctx, d_mov.alloc(3 * npts));
    FD_CUDA(ctx, d_ref.alloc(3 * npts));
    FD_CUDA(ctx, d_off.alloc((size_t)n_align + 1));
    FD_CUDA(ctx, d_rmsd.alloc(n_align));
    FD_CUDA(ctx, d_U.alloc(9 * (size_t)n_align));
    FD_CUDA(ctx, d_t.alloc(3 * (size_t)n_align));
    cudaStream_t s = ctx->stream;
    FD_CUDA(ctx, cudaMemcpyAsync(d_mov.p, mov_xyz, 12 * npts, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_ref.p, ref_xyz, 12 * npts, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_off.p, pt_offsets, ((size_t)n_align + 1) * 4, cudaMemcpyHostToDevice, s));
    StageTimer st(ctx, "kabsch");
    FD_LAUNCH(ctx, k5_kabsch, fd_div_up(n_align, 128), 128, 0, d_mov.p, d_ref.p, d_off.p, n_align, d_rmsd.p, d_U.p,
              d_t.p);
    FD_CUDA(ctx, cudaMemcpyAsync(rmsd, d_rmsd.p, 4 * (size_t)n_align, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, cudaMemcpyAsync(U9, d_U.p, 36 * (size_t)n_align, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, cudaMemcpyAsync(t3, d_t.p, 12 * (size_t)n_align, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, st.finish());
    return FD_OK;
}

// rmsd_with_calpha_and_rottran (src/controller/retrieve.rs:756-834) for a batch: coordinates of the matched
// target residues are gathered on the device from the attached store, CA and CB interleaved (:761-767).
static int superpose_store_batch(fd_ctx *ctx, bool partial_fit, const float *q_ca_xyz, const float *q_cb_xyz, uint64_t n_q_res,
                                 const uint32_t *align_nid, const uint32_t *pair_offsets, uint32_t n_align,
                                 const uint32_t *pair_qres, const uint32_t *pair_tres, float *rmsd, float *U9, float *t3) {
    if (!ctx) return FD_ERR_ARG;
    if (n_align == 0) return FD_OK;
    if (!ctx->store.attached) return fd_fail(ctx, FD_ERR_STATE, "fd_kabsch_store_batch: no structure store attached");
    if (!q_ca_xyz || !q_cb_xyz || !align_nid || !pair_offsets || !pair_qres || !pair_tres || !rmsd || !U9 || !t3)
        return fd_fail(ctx, FD_ERR_ARG, "fd_kabsch_store_batch: NULL argument");
    const uint64_t np = pair_offsets[n_align];
    for (uint32_t a = 0; a < n_align; a++)
        if (align_nid[a] >= ctx->store.n_structs) return fd_fail(ctx, FD_ERR_ARG, "align_nid outside the store");
    for (uint64_t k = 0; k < np; k++)
        if (pair_qres[k] >= n_q_res) return fd_fail(ctx, FD_ERR_ARG, "pair_qres out of range");
    if (partial_fit)
        for (uint32_t a = 0; a < n_align; a++)
            if (2 * (pair_offsets[a + 1] - pair_offsets[a]) > fdlq::LMS_MAX_POINTS)
                return fd_fail(ctx, FD_ERR_LIMIT, "fd_lmsqcp_store_batch: at most 64 matched residues per alignment");
    FD_ENTER(ctx);
    DevBuf<float> d_qca, d_qcb, d_rmsd, d_U, d_t;
    DevBuf<uint32_t> d_nid, d_off, d_pq, d_pt;
    FD_CUDA(ctx, d_qca.alloc(3 * n_q_res));
    FD_CUDA(ctx, d_qcb.alloc(3 * n_q_res));
    FD_CUDA(ctx, d_nid.alloc(n_align));
    FD_CUDA(ctx, d_off.alloc((size_t)n_align + 1));
    FD_CUDA(ctx, d_pq.alloc(np));
    FD_CUDA(ctx, d_pt.alloc(np));
    FD_CUDA(ctx, d_rmsd.alloc(n_align));
    FD_CUDA(ctx, d_U.alloc(9 * (size_t)n_align));
    FD_CUDA(ctx, d_t.alloc(3 * (size_t)n_align));
    cudaStream_t s = ctx->stream;
    FD_CUDA(ctx, cudaMemcpyAsync(d_qca.p, q_ca_xyz, 12 * n_q_res, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_qcb.p, q_cb_xyz, 12 * n_q_res, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_nid.p, align_nid, 4 * (size_t)n_align, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_off.p, pair_offsets, 4 * ((size_t)n_align + 1), cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_pq.p, pair_qres, 4 * np, cudaMemcpyHostToDevice, s));
    FD_CUDA(ctx, cudaMemcpyAsync(d_pt.p, pair_tres, 4 * np, cudaMemcpyHostToDevice, s));
    StageTimer st(ctx, "kabsch");
    const FdDeviceStore &S = ctx->store;
    if (partial_fit)
        FD_LAUNCH(ctx, k5_lmsqcp_store, fd_div_up(n_align, 64), 64, 0, d_qca.p, d_qcb.p, S.ca_xyz, S.cb_xyz,
                  S.row_offsets, d_nid.p, d_off.p, d_pq.p, d_pt.p, n_align, d_rmsd.p, d_U.p, d_t.p);
    else
        FD_LAUNCH(ctx, k5_kabsch_store, fd_div_up(n_align, 128), 128, 0, d_qca.p, d_qcb.p, S.ca_xyz, S.cb_xyz,
                  S.row_offsets, d_nid.p, d_off.p, d_pq.p, d_pt.p, n_align, d_rmsd.p, d_U.p, d_t.p);
    FD_CUDA(ctx, cudaMemcpyAsync(rmsd, d_rmsd.p, 4 * (size_t)n_align, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, cudaMemcpyAsync(U9, d_U.p, 36 * (size_t)n_align, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, cudaMemcpyAsync(t3, d_t.p, 12 * (size_t)n_align, cudaMemcpyDeviceToHost, s));
    FD_CUDA(ctx, st.finish());
    return FD_OK;
}

extern "C" int fd_kabsch_store_batch(fd_ctx *ctx, const float *q_ca_xyz, const float *q_cb_xyz, uint64_t n_q_res,
                                     const uint32_t *align_nid, const uint32_t *pair_offsets, uint32_t n_align,
                                     const uint32_t *pair_qres, const uint32_t *pair_tres, float *rmsd, float *U9,
                                     float *t3) {
    return superpose_store_batch(ctx, false, q_ca_xyz, q_cb_xyz, n_q_res, align_nid, pair_offsets, n_align, pair_qres,
                                 pair_tres, rmsd, U9, t3);
}
// rmsd_with_calpha_and_rottran with lms = true (`--partial-fit`, src/controller/retrieve.rs:773-814)
extern "C" int fd_lmsqcp_store_batch(fd_ctx *ctx, const float *q_ca_xyz, const float *q_cb_xyz, uint64_t n_q_res,
                                     const uint32_t *align_nid, const uint32_t *pair_offsets, uint32_t n_align,
                                     const uint32_t *pair_qres, const uint32_t *pair_tres, float *rmsd, float *U9,
                                     float *t3) {
    return superpose_store_batch(ctx, true, q_ca_xyz, q_cb_xyz, n_q_res, align_nid, pair_offsets, n_align, pair_qres,
                                 pair_tres, rmsd, U9, t3);
}
// the same LMS-QCP on the host build of csrc/fd_lmsqcp.cuh over explicit point lists (parity probe, no device);
// returns 0, or -1 when n_points is outside 3 .. 128
extern "C" int fd_lmsqcp_host(const float *ref_xyz, const float *mov_xyz, uint32_t n_points, float *U9, float *t3,
                              float *rms_inliers) {
    return fdlq::lms_qcp_one(n_points, fdmt::FlatF{ref_xyz}, fdmt::FlatF{mov_xyz}, U9, t3, rms_inliers) ? 0 : -1;
}
