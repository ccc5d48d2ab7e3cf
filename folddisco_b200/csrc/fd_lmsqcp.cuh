// fd_lmsqcp.cuh -- LMS-QCP partial superposition (`--partial-fit`), one implementation for host and device.
//
// Mirrors LmsQcpSuperimposer::run with its default parameters (reference src/structure/lms_qcp.rs:28-40, 91-196):
//   1. 500 seed trials: three non-collinear point pairs drawn by the reference's own xorshift64* generator (fixed seed
//      0xC0FFEE005EED, :543-564), superposed by QCP, scored by the median squared residual of the other pairs; the
//      best-scoring triple wins (:104-137);
//   2. forward search: the core grows by the pair with the smallest residual under the core's current superposition
//      until it holds n / 2 pairs and the next residual exceeds r_max = 2 A, or every pair is in (:140-195);
//   3. rms over the core under the last superposition (:222-236).
// Incremental sums and the QCP characteristic polynomial / quaternion are binary64 (:245-447), the transform is rounded to
// f32 and applied in f32 (:479-505), like the reference: selection steps compare f32 values, so the arithmetic order is
// kept term by term.  Deterministic -- no host randomness.  At most LMS_MAX_POINTS point pairs (2 x 64 matched residues).
#pragma once
#include "fd_metrics.cuh"

namespace fdlq {

constexpr uint32_t LMS_MAX_POINTS = 128;
using fdmt::P3f;

struct Stats { // RunningStats (lms_qcp.rs:245-287)
    uint32_t n;
    double sum_x[3], sum_y[3], sxx, syy, syx[3][3];
};
FD_HD void stats_init(Stats &s) {
    s.n = 0;
    s.sxx = s.syy = 0.0;
    for (int a = 0; a < 3; a++) {
        s.sum_x[a] = s.sum_y[a] = 0.0;
        for (int b = 0; b < 3; b++) s.syx[a][b] = 0.0;
    }
}
FD_HD void stats_add(Stats &s, P3f mov, P3f ref) {
    const double x[3] = {(double)mov.x, (double)mov.y, (double)mov.z}, y[3] = {(double)ref.x, (double)ref.y, (double)ref.z};
    s.n += 1;
    for (int a = 0; a < 3; a++) {
        s.sum_x[a] += x[a];
        s.sum_y[a] += y[a];
    }
    s.sxx += x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    s.syy += y[0] * y[0] + y[1] * y[1] + y[2] * y[2];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) s.syx[a][b] += y[a] * x[b];
}

// qcp_from_a_e0 (lms_qcp.rs:330-447): largest root of the QCP quartic by Newton from e0, rotation from the quaternion
FD_HD void qcp_rotation(const double a[3][3], double e0, double rot[3][3]) {
    const double sxx = a[0][0], sxy = a[0][1], sxz = a[0][2], syx = a[1][0], syy = a[1][1], syz = a[1][2], szx = a[2][0],
                 szy = a[2][1], szz = a[2][2];
    const double sxx2 = sxx * sxx, syy2 = syy * syy, szz2 = szz * szz, sxy2 = sxy * sxy, syz2 = syz * syz, sxz2 = sxz * sxz,
                 syx2 = syx * syx, szy2 = szy * szy, szx2 = szx * szx;
    const double syz_szy_m_syy_szz2 = 2.0 * (syz * szy - syy * szz);
    const double sxx2_syy2_szz2_syz2_szy2 = syy2 + szz2 - sxx2 + syz2 + szy2;
    const double c2 = -2.0 * (sxx2 + syy2 + szz2 + sxy2 + syx2 + sxz2 + szx2 + syz2 + szy2);
    const double c1 = 8.0 * (sxx * syz * szy + syy * szx * sxz + szz * sxy * syx - sxx * syy * szz - syz * szx * sxy - szy * syx * sxz);
    const double sxz_p_szx = sxz + szx, syz_p_szy = syz + szy, sxy_p_syx = sxy + syx, syz_m_szy = syz - szy,
                 sxz_m_szx = sxz - szx, sxy_m_syx = sxy - syx, sxx_p_syy = sxx + syy, sxx_m_syy = sxx - syy;
    const double sxy2_sxz2_syx2_szx2 = sxy2 + sxz2 - syx2 - szx2;
    const double neg_sxz_p_szx = -sxz_p_szx, neg_sxz_m_szx = -sxz_m_szx, neg_sxy_m_syx = -sxy_m_syx;
    const double sxx_p_syy_p_szz = sxx_p_syy + szz;
    const double c0 = sxy2_sxz2_syx2_szx2 * sxy2_sxz2_syx2_szx2 +
                      (sxx2_syy2_szz2_syz2_szy2 + syz_szy_m_syy_szz2) * (sxx2_syy2_szz2_syz2_szy2 - syz_szy_m_syy_szz2) +
                      (neg_sxz_p_szx * (syz_m_szy) + (sxy_m_syx) * (sxx_m_syy - szz)) *
                          (neg_sxz_m_szx * (syz_p_szy) + (sxy_m_syx) * (sxx_m_syy + szz)) +
                      (neg_sxz_p_szx * (syz_p_szy) - (sxy_p_syx) * (sxx_p_syy - szz)) *
                          (neg_sxz_m_szx * (syz_m_szy) - (sxy_p_syx)*sxx_p_syy_p_szz) +
                      ((sxy_p_syx) * (syz_p_szy) + (sxz_p_szx) * (sxx_m_syy + szz)) *
                          (neg_sxy_m_syx * (syz_m_szy) + (sxz_p_szx)*sxx_p_syy_p_szz) +
                      ((sxy_p_syx) * (syz_m_szy) + (sxz_m_szx) * (sxx_m_syy - szz)) *
                          (neg_sxy_m_syx * (syz_p_szy) + (sxz_m_szx) * (sxx_p_syy - szz));
    double lam = e0 > 0.0 ? e0 : 0.0;
    const double eps = 1e-15;
    for (int it = 0; it < 10; it++) {
        const double x2 = lam * lam;
        const double b = (x2 + c2) * lam;
        const double aa = b + c1;
        const double f = aa * lam + c0;
        const double fp = 2.0 * x2 * lam + b + aa;
        const double delta = f / (fp + eps);
        const double nlam = fabs(lam - delta);
        if (fabs(nlam - lam) < eps * nlam) {
            lam = nlam;
            break;
        }
        lam = nlam;
    }
    const double a11 = sxx_p_syy + szz - lam, a12 = syz_m_szy, a13 = neg_sxz_m_szx, a14 = sxy_m_syx;
    const double a21 = a12, a22 = sxx_m_syy - szz - lam, a23 = sxy_p_syx, a24 = sxz_p_szx;
    const double a31 = a13, a32 = a23, a33 = syy - sxx - szz - lam, a34 = syz_p_szy;
    const double a41 = a14, a42 = a24, a43 = a34, a44 = szz - sxx_p_syy - lam;
    const double a3344_4334 = a33 * a44 - a43 * a34, a3244_4234 = a32 * a44 - a42 * a34, a3243_4233 = a32 * a43 - a42 * a33,
                 a3143_4133 = a31 * a43 - a41 * a33, a3144_4134 = a31 * a44 - a41 * a34, a3142_4132 = a31 * a42 - a41 * a32;
    double q1 = a22 * a3344_4334 - a23 * a3244_4234 + a24 * a3243_4233;
    double q2 = -a21 * a3344_4334 + a23 * a3144_4134 - a24 * a3143_4133;
    double q3 = a21 * a3244_4234 - a22 * a3144_4134 + a24 * a3142_4132;
    double q4 = -a21 * a3243_4233 + a22 * a3143_4133 - a23 * a3142_4132;
    double qsqr = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
    const double evec_prec = 1e-12;
    if (qsqr < evec_prec) {
        q1 = a12 * a3344_4334 - a13 * a3244_4234 + a14 * a3243_4233;
        q2 = -a11 * a3344_4334 + a13 * a3144_4134 - a14 * a3143_4133;
        q3 = a11 * a3244_4234 - a12 * a3144_4134 + a14 * a3142_4132;
        q4 = -a11 * a3243_4233 + a12 * a3143_4133 - a13 * a3142_4132;
        qsqr = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
        if (qsqr < evec_prec) {
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) rot[r][c] = r == c ? 1.0 : 0.0;
            return;
        }
    }
    const double inv = 1.0 / sqrt(qsqr);
    q1 *= inv;
    q2 *= inv;
    q3 *= inv;
    q4 *= inv;
    const double a2 = q1 * q1, x2 = q2 * q2, y2 = q3 * q3, z2 = q4 * q4, xy = q2 * q3, az = q1 * q4, zx = q4 * q2, ay = q1 * q3,
                 yz = q3 * q4, ax = q1 * q2;
    rot[0][0] = a2 + x2 - y2 - z2;
    rot[0][1] = 2.0 * (xy + az);
    rot[0][2] = 2.0 * (zx - ay);
    rot[1][0] = 2.0 * (xy - az);
    rot[1][1] = a2 - x2 + y2 - z2;
    rot[1][2] = 2.0 * (yz + ax);
    rot[2][0] = 2.0 * (zx + ay);
    rot[2][1] = 2.0 * (yz - ax);
    rot[2][2] = a2 - x2 - y2 + z2;
}

// qcp_from_stats (lms_qcp.rs:294-326): U (row-major f32) and t = mean_y - R mean_x
FD_HD void qcp_from_stats(const Stats &st, float *U, float *t) {
    const double inv = 1.0 / (double)st.n;
    const double mux[3] = {st.sum_x[0] * inv, st.sum_x[1] * inv, st.sum_x[2] * inv};
    const double muy[3] = {st.sum_y[0] * inv, st.sum_y[1] * inv, st.sum_y[2] * inv};
    double a[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) a[r][c] = st.syx[r][c] - (double)st.n * (muy[r] * mux[c]);
    const double mu2x = mux[0] * mux[0] + mux[1] * mux[1] + mux[2] * mux[2];
    const double mu2y = muy[0] * muy[0] + muy[1] * muy[1] + muy[2] * muy[2];
    double e0 = (st.syy - (double)st.n * mu2y) + (st.sxx - (double)st.n * mu2x);
    e0 = 0.5 * (e0 > 0.0 ? e0 : 0.0);
    double rot[3][3];
    qcp_rotation(a, e0, rot);
    for (int r = 0; r < 3; r++) {
        const double rx = rot[r][0] * mux[0] + rot[r][1] * mux[1] + rot[r][2] * mux[2];
        t[r] = (float)(muy[r] - rx);
        for (int c = 0; c < 3; c++) U[3 * r + c] = (float)rot[r][c];
    }
}

FD_HD float dist2_f32(P3f a, P3f b) { // lms_qcp.rs:501-505
    const float dx = FD_SUB(a.x, b.x), dy = FD_SUB(a.y, b.y), dz = FD_SUB(a.z, b.z);
    return FD_ADD(FD_ADD(FD_MUL(dx, dx), FD_MUL(dy, dy)), FD_MUL(dz, dz));
}

// value of the element a full sort would put at position round(q * (n - 1)) (select_quantile_squared, :451-463)
FD_HD float select_quantile(float *v, uint32_t n, float q) {
    if (n == 0) return 0.0f;
    if (n == 1) return v[0];
    const uint32_t pos = (uint32_t)roundf(FD_MUL(q, (float)(n - 1)));
    for (uint32_t i = 0; i <= pos; i++) { // partial selection sort: n <= 125
        uint32_t m = i;
        for (uint32_t j = i + 1; j < n; j++)
            if (v[j] < v[m]) m = j;
        const float tmp = v[i];
        v[i] = v[m];
        v[m] = tmp;
    }
    return v[pos];
}

struct Rng { // SmallRng (lms_qcp.rs:543-564)
    unsigned long long state;
    FD_HD explicit Rng(unsigned long long seed) {
        unsigned long long x = seed + 0x9E3779B97F4A7C15ull;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        state = x ^ (x >> 31);
    }
    FD_HD unsigned long long next() {
        unsigned long long x = state;
        x ^= x << 13;
        x ^= x >> 7;
        x ^= x << 17;
        state = x;
        return x;
    }
    FD_HD uint32_t below(uint32_t end) { return (uint32_t)(next() % (unsigned long long)end); }
};

// LmsQcpSuperimposer::run + finish.  ref(i) / mov(i): point i of the reference (query) and moving (target) set.
// Returns false when n is outside 3 .. LMS_MAX_POINTS (the caller uses Kabsch for three residues or fewer anyway).
template <class RefAt, class MovAt>
FD_HD bool lms_qcp_one(uint32_t n, RefAt ref, MovAt mov, float *U, float *t, float *rms_inliers) {
    if (n < 3 || n > LMS_MAX_POINTS) return false;
    Rng rng(0xC0FFEE005EEDull);
    uint32_t best_seed[3] = {0, 1, 2};
    float best_qval = (float)HUGE_VAL;
    float res_sq[LMS_MAX_POINTS];
    Stats st;
    for (int trial = 0; trial < 500; trial++) {
        uint32_t seed[3];
        bool found = false;
        for (int tries = 0; tries < 64 && !found; tries++) { // sample_three_non_collinear (:509-526)
            const uint32_t i = rng.below(n);
            uint32_t j = rng.below(n);
            if (j == i) j = (j + 1) % n;
            uint32_t k = rng.below(n);
            while (k == i || k == j) k = (k + 1) % n;
            const P3f pi = mov(i), pj = mov(j), pk = mov(k);
            const float v1[3] = {FD_SUB(pj.x, pi.x), FD_SUB(pj.y, pi.y), FD_SUB(pj.z, pi.z)};
            const float v2[3] = {FD_SUB(pk.x, pi.x), FD_SUB(pk.y, pi.y), FD_SUB(pk.z, pi.z)};
            const float cx = FD_SUB(FD_MUL(v1[1], v2[2]), FD_MUL(v1[2], v2[1]));
            const float cy = FD_SUB(FD_MUL(v1[2], v2[0]), FD_MUL(v1[0], v2[2]));
            const float cz = FD_SUB(FD_MUL(v1[0], v2[1]), FD_MUL(v1[1], v2[0]));
            const float area2 = FD_ADD(FD_ADD(FD_MUL(cx, cx), FD_MUL(cy, cy)), FD_MUL(cz, cz));
            if (area2 > 1e-6f) {
                seed[0] = i, seed[1] = j, seed[2] = k;
                found = true;
            }
        }
        if (!found) continue;
        stats_init(st);
        for (int s = 0; s < 3; s++) stats_add(st, mov(seed[s]), ref(seed[s]));
        float r[9], tt[3];
        qcp_from_stats(st, r, tt);
        uint32_t m = 0;
        for (uint32_t i = 0; i < n; i++) {
            if (i == seed[0] || i == seed[1] || i == seed[2]) continue;
            res_sq[m++] = dist2_f32(fdmt::transform(r, tt, mov(i)), ref(i));
        }
        const float qv = select_quantile(res_sq, m, 0.5f);
        if (qv < best_qval) {
            best_qval = qv;
            best_seed[0] = seed[0], best_seed[1] = seed[1], best_seed[2] = seed[2];
        }
    }
    // forward search
    uint32_t min_core = n / 2;
    if (min_core < 3) min_core = 3;
    const float r2_max = FD_MUL(2.0f, 2.0f);
    uint8_t in_core[LMS_MAX_POINTS];
    uint8_t core[LMS_MAX_POINTS];
    for (uint32_t i = 0; i < n; i++) in_core[i] = 0;
    stats_init(st);
    uint32_t n_core = 0;
    for (int s = 0; s < 3; s++) {
        stats_add(st, mov(best_seed[s]), ref(best_seed[s]));
        in_core[best_seed[s]] = 1;
        core[n_core++] = (uint8_t)best_seed[s];
    }
    for (;;) {
        qcp_from_stats(st, U, t);
        int best_i = -1;
        float best_r2 = (float)HUGE_VAL;
        for (uint32_t i = 0; i < n; i++) {
            if (in_core[i]) continue;
            const float d2 = dist2_f32(fdmt::transform(U, t, mov(i)), ref(i));
            if (d2 < best_r2) {
                best_r2 = d2;
                best_i = (int)i;
            }
        }
        if (best_i < 0) break;
        if (n_core >= min_core && best_r2 > r2_max) break;
        stats_add(st, mov((uint32_t)best_i), ref((uint32_t)best_i));
        in_core[best_i] = 1;
        core[n_core++] = (uint8_t)best_i;
        if (n_core == n) break; // finish(r, t) with the transform computed BEFORE the last pair joined (:186-189)
    }
    float sum = 0.0f;
    for (uint32_t c = 0; c < n_core; c++) sum = FD_ADD(sum, dist2_f32(fdmt::transform(U, t, mov(core[c])), ref(core[c])));
    *rms_inliers = FD_SQRT(FD_DIV(sum, (float)n_core));
    return true;
}

} // namespace fdlq
