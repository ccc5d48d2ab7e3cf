// fd_metrics.cuh -- similarity metrics of one superposed match, one implementation for host and device.
//
// Mirrors StructureSimilarityMetrics::calculate_all over PrecomputedDistances::new(reference, transformed)
// (reference src/structure/metrics.rs:44-92, 108-273; called from rmsd_with_calpha_and_rottran,
// src/controller/retrieve.rs:776-831) with the reference's arithmetic:
//   transformed point   = U * x + t in f32, row by row (src/structure/kabsch.rs:141-151)
//   get_distance(i, j)  = sqrt(f64 sum of squares) rounded to f32, between reference point j and transformed point i
//   tm_score / gdt_*    use get_distance(i, i) where their formulas want a SQUARED distance (metrics.rs:133-173: the
//                       value is compared with cutoff^2 and divided by d0^2 although it is a distance) -- reproduced
//   chamfer / hausdorff = mean / max over transformed points of the distance to the nearest reference point
// No n x n matrix is stored: every distance is evaluated where it is used (n = 2 x matched residues <= ~130 points).
#pragma once
#include "fd_geom.cuh"

namespace fdmt {

struct P3f {
    float x, y, z;
};

FD_HD float dist_f32(P3f a, P3f b) { // metrics.rs:97-111
    const double dx = (double)a.x - (double)b.x, dy = (double)a.y - (double)b.y, dz = (double)a.z - (double)b.z;
    return (float)sqrt(dx * dx + dy * dy + dz * dz);
}
FD_HD P3f transform(const float *U, const float *t, P3f v) { // matrix_vector_multiply + add_vec, f32
    P3f r;
    r.x = FD_ADD(FD_ADD(FD_ADD(FD_MUL(U[0], v.x), FD_MUL(U[1], v.y)), FD_MUL(U[2], v.z)), t[0]);
    r.y = FD_ADD(FD_ADD(FD_ADD(FD_MUL(U[3], v.x), FD_MUL(U[4], v.y)), FD_MUL(U[5], v.z)), t[1]);
    r.z = FD_ADD(FD_ADD(FD_ADD(FD_MUL(U[6], v.x), FD_MUL(U[7], v.y)), FD_MUL(U[8], v.z)), t[2]);
    return r;
}
FD_HD float d0_scale(uint32_t n) { // metrics.rs:116-122
    if (n > 21) return FD_SUB(FD_MUL(1.24f, powf(FD_SUB((float)n, 15.0f), 1.0f / 3.0f)), 1.8f);
    return 0.5f;
}

// out[5] = tm_score, gdt_ts, gdt_ha, chamfer_distance, hausdorff_distance.  ref(i) / mov(i): point i of the reference
// (query) and of the moving (target) set; U (row-major 3x3), t: the superposition of mov onto ref.
template <class RefAt, class MovAt>
FD_HD void similarity_metrics(uint32_t n, RefAt ref, MovAt mov, const float *U, const float *t, float *out) {
    if (n == 0) { // metrics.rs: tm / gdt 0, chamfer / hausdorff infinity
        out[0] = out[1] = out[2] = 0.0f;
        out[3] = out[4] = (float)HUGE_VAL;
        return;
    }
    const float d0 = d0_scale(n);
    const double d0_sq = (double)FD_MUL(d0, d0);
    const double ts_cut[4] = {1.0, 2.0, 4.0, 8.0}, ha_cut[4] = {0.5, 1.0, 2.0, 4.0};
    uint32_t ts_cnt[4] = {0, 0, 0, 0}, ha_cnt[4] = {0, 0, 0, 0};
    double tm_sum = 0.0, chamfer_sum = 0.0;
    float hausdorff = 0.0f;
    for (uint32_t i = 0; i < n; i++) {
        const P3f c = transform(U, t, mov(i));
        const double dii = (double)dist_f32(ref(i), c);
        tm_sum += 1.0 / (1.0 + dii / d0_sq);
        for (int k = 0; k < 4; k++) {
            ts_cnt[k] += dii <= ts_cut[k] * ts_cut[k] ? 1u : 0u;
            ha_cnt[k] += dii <= ha_cut[k] * ha_cut[k] ? 1u : 0u;
        }
        float mn = dist_f32(ref(0), c);
        for (uint32_t j = 1; j < n; j++) {
            const float d = dist_f32(ref(j), c);
            if (d < mn) mn = d;
        }
        chamfer_sum += (double)mn;
        if (i == 0 || mn > hausdorff) hausdorff = mn;
    }
    double ts = 0.0, ha = 0.0;
    for (int k = 0; k < 4; k++) {
        ts += (double)ts_cnt[k] / (double)n;
        ha += (double)ha_cnt[k] / (double)n;
    }
    out[0] = (float)(tm_sum / (double)n);
    out[1] = (float)(ts / 4.0);
    out[2] = (float)(ha / 4.0);
    out[3] = (float)(chamfer_sum / (double)n);
    out[4] = hausdorff;
}

// CA, CB interleaved, gathered by residue index (point 2k = CA(res[k]), 2k + 1 = CB(res[k])): retrieve.rs:761-767
struct GatherF {
    const float *ca, *cb;
    const uint32_t *res;
    uint64_t base;
    FD_HD P3f operator()(uint32_t i) const {
        const uint64_t r = base + res[i >> 1];
        const float *s = (i & 1) ? cb : ca;
        return {s[3 * r], s[3 * r + 1], s[3 * r + 2]};
    }
};
struct FlatF {
    const float *p;
    FD_HD P3f operator()(uint32_t i) const { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
};

} // namespace fdmt
