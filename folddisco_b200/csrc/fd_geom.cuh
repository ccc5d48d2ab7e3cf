// fd_geom.cuh -- pair feature + PDBTrRosetta perfect hash, one implementation for host and device.
//
// Mirrors, operation by operation (f32, round-to-nearest, no FMA):
//   CompactStructure::get_pdb_tr_feature   reference src/structure/core.rs:378-403
//   Coordinate::calc_distance / calc_angle  src/structure/coordinate.rs:109-133
//   calc_torsion_radian                      src/structure/coordinate.rs:203-215 (normalize = 3 divisions, :69-76)
//   pdb_tr::HashValue::perfect_hash          src/geometry/pdb_tr.rs:21-75
//   discretize_f32_value_into_u32            src/utils/convert.rs:32-36 (Rust `as u32` saturates, NaN -> 0)
#pragma once
#include "fd_math.cuh"

#if defined(__CUDA_ARCH__)
#define FD_MUL(a, b) __fmul_rn((a), (b))
#define FD_ADD(a, b) __fadd_rn((a), (b))
#define FD_SUB(a, b) __fsub_rn((a), (b))
#define FD_DIV(a, b) __fdiv_rn((a), (b))
#define FD_SQRT(a) __fsqrt_rn((a))
#else
#define FD_MUL(a, b) ((a) * (b))
#define FD_ADD(a, b) ((a) + (b))
#define FD_SUB(a, b) ((a) - (b))
#define FD_DIV(a, b) ((a) / (b))
#define FD_SQRT(a) sqrtf((a))
#endif

namespace fdg {

struct V3 {
    float x, y, z;
};

FD_HD V3 sub(V3 a, V3 b) { return {FD_SUB(a.x, b.x), FD_SUB(a.y, b.y), FD_SUB(a.z, b.z)}; }
FD_HD float dot(V3 a, V3 b) { return FD_ADD(FD_ADD(FD_MUL(a.x, b.x), FD_MUL(a.y, b.y)), FD_MUL(a.z, b.z)); }
FD_HD V3 cross(V3 a, V3 b) {
    return {FD_SUB(FD_MUL(a.y, b.z), FD_MUL(a.z, b.y)), FD_SUB(FD_MUL(a.z, b.x), FD_MUL(a.x, b.z)),
            FD_SUB(FD_MUL(a.x, b.y), FD_MUL(a.y, b.x))};
}
FD_HD float norm(V3 a) { return FD_SQRT(dot(a, a)); }
FD_HD V3 normalize(V3 a) {
    float n = norm(a);
    return {FD_DIV(a.x, n), FD_DIV(a.y, n), FD_DIV(a.z, n)};
}
FD_HD float dist(V3 a, V3 b) { return norm(sub(a, b)); }

// squared distance with the same rounding as dist(); sqrt is monotone, so
// (dist <= cutoff) can be pre-screened on d2 but the final test uses dist() itself.
FD_HD float dist2(V3 a, V3 b) {
    V3 d = sub(a, b);
    return dot(d, d);
}

FD_HD float torsion(V3 a, V3 b, V3 c, V3 d) {
    V3 v1 = sub(b, a), v2 = sub(c, b), v3 = sub(d, c);
    V3 r = normalize(cross(v1, v2));
    V3 s = normalize(cross(v2, v3));
    V3 t = normalize(cross(r, normalize(v2)));
    float x = dot(r, s);
    float y = dot(s, t);
    return -fdm::atan2f_exact(y, x);
}

// Rust `f32 as u32`
FD_HD uint32_t sat_u32(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

struct HashParams {
    float nbin_dist;  // already resolved: 0 -> 16, >16 -> 16 (pdb_tr.rs:22-28)
    float nbin_angle; // 0 -> 4, >4 -> 4 (pdb_tr.rs:29-35)
    float dist_cutoff;
};

FD_HD HashParams make_params(uint32_t nbd, uint32_t nba, float cutoff) {
    HashParams p;
    // every call site of the reference takes perfect_hash_default when EITHER count is 0 (feature.rs:215-221,
    // query.rs:72-76, retrieve.rs:133-137)
    if (nbd == 0 || nba == 0) nbd = nba = 0;
    p.nbin_dist = nbd > 16 ? 16.0f : (nbd == 0 ? 16.0f : (float)nbd);
    p.nbin_angle = nba > 4 ? 4.0f : (nba == 0 ? 4.0f : (float)nba);
    p.dist_cutoff = cutoff;
    return p;
}

FD_HD uint32_t discretize(float val, float mn, float mx, float nbin) {
    float cont_f = FD_DIV(FD_SUB(mx, mn), FD_SUB(nbin, 1.0f));
    float disc_f = FD_DIV(1.0f, cont_f);
    return sat_u32(FD_ADD(FD_MUL(FD_SUB(val, mn), disc_f), 0.5f));
}

// feature = [res1, res2, ca_dist, cb_dist, ca_cb_angle, theta1, theta2]
FD_HD uint32_t perfect_hash(const float *f, const HashParams &p) {
    uint32_t res1 = sat_u32(f[0]), res2 = sat_u32(f[1]);
    uint32_t ca = discretize(f[2], 2.0f, 20.0f, p.nbin_dist);
    uint32_t cb = discretize(f[3], 2.0f, 20.0f, p.nbin_dist);
    float s, c;
    fdm::sincosf_exact(f[4], &s, &c);
    uint32_t s0 = discretize(s, -1.0f, 1.0f, p.nbin_angle), c0 = discretize(c, -1.0f, 1.0f, p.nbin_angle);
    fdm::sincosf_exact(f[5], &s, &c);
    uint32_t s1 = discretize(s, -1.0f, 1.0f, p.nbin_angle), c1 = discretize(c, -1.0f, 1.0f, p.nbin_angle);
    fdm::sincosf_exact(f[6], &s, &c);
    uint32_t s2 = discretize(s, -1.0f, 1.0f, p.nbin_angle), c2 = discretize(c, -1.0f, 1.0f, p.nbin_angle);
    return res1 << 25 | res2 << 20 | ca << 16 | cb << 12 | s0 << 10 | c0 << 8 | s1 << 6 | c1 << 4 | s2 << 2 | c2;
}

// Bits 12..31 of perfect_hash (amino acids, CA and CB distance bins; fields OR-ed in unmasked exactly as above).
// The six sin / cos bins are at most nbin_angle - 1 <= 3, so they never reach bit 12: a pair can only hash into a
// set that holds a hash with the same upper bits -- a trigonometry-free necessary condition for set membership.
FD_HD uint32_t hash_upper(V3 cb1, V3 cb2, uint8_t aa1, uint8_t aa2, float ca_dist, const HashParams &p) {
    uint32_t ca = discretize(ca_dist, 2.0f, 20.0f, p.nbin_dist);
    uint32_t cb = discretize(dist(cb1, cb2), 2.0f, 20.0f, p.nbin_dist);
    return ((uint32_t)aa1 << 25 | (uint32_t)aa2 << 20 | ca << 16 | cb << 12) >> 12;
}

// Geometry part of the feature for residues i -> j.  The caller has already checked aa != 255, cb_valid and
// ca_dist <= cutoff (ca_dist is passed in so it is computed once, with dist()).
FD_HD void pair_feature(V3 n1, V3 ca1, V3 cb1, V3 n2, V3 ca2, V3 cb2, float aa1, float aa2, float ca_dist,
                        float *f) {
    f[0] = aa1;
    f[1] = aa2;
    f[2] = ca_dist;
    f[3] = dist(cb1, cb2);
    { // calc_angle(ca1, cb1, ca2, cb2): v1 = cb1 - ca1, v2 = cb2 - ca2
        V3 v1 = sub(cb1, ca1), v2 = sub(cb2, ca2);
        float d = dot(v1, v2);
        float l1 = norm(v1), l2 = norm(v2);
        float cs = FD_DIV(d, FD_MUL(l1, l2));
        f[4] = fdm::acosf_exact(cs);
    }
    f[5] = torsion(n1, ca1, cb1, cb2);
    f[6] = torsion(cb1, cb2, ca2, n2);
}

FD_HD uint32_t pair_hash(V3 n1, V3 ca1, V3 cb1, V3 n2, V3 ca2, V3 cb2, uint8_t aa1, uint8_t aa2, float ca_dist,
                         const HashParams &p) {
    float f[7];
    pair_feature(n1, ca1, cb1, n2, ca2, cb2, (float)aa1, (float)aa2, ca_dist, f);
    return perfect_hash(f, p);
}

// ---- fast path of the angle bins ---------------------------------------------------------------------------------
// perfect_hash needs only the BINS of sin / cos of the three angles, and the angles are themselves acos / atan2 of
// quantities the feature already has:
//     ca_cb_angle = acos(c)       =>  cos = c,  sin = sqrt(1 - c^2)
//     torsion     = -atan2(y, x)  =>  sin = -y / sqrt(x^2 + y^2),  cos = x / sqrt(x^2 + y^2)
// The reference's f32 chain (acosf / atan2f rounded to f32, then sinf / cosf rounded to f32) stays within 4e-7 of
// these closed forms (half an ulp of an angle <= pi is 1.2e-7, sin and cos are 1-Lipschitz, one more rounding of
// 6e-8), the f32 evaluation below adds < 4e-6, and discretize() scales by at most 1.5: the value that is truncated
// to the bin differs by < 1e-5 between the two routes.  So when that value is at least 1e-4 away from an integer
// the bin is decided without any trigonometry; otherwise (about one pair in a thousand, and for degenerate
// geometry) the caller falls back to the exact route.  pair_hash_auto therefore returns exactly pair_hash's value.
FD_HD bool fast_bin(float val, float nbin, uint32_t *out) {
    float cont_f = FD_DIV(FD_SUB(1.0f, -1.0f), FD_SUB(nbin, 1.0f));
    float disc_f = FD_DIV(1.0f, cont_f);
    float t = FD_ADD(FD_MUL(FD_SUB(val, -1.0f), disc_f), 0.5f);
    float fr = FD_SUB(t, floorf(t));
    if (!(fr > 1.0e-4f && fr < 1.0f - 1.0e-4f)) return false; // near a bin boundary, or NaN
    *out = sat_u32(t);
    return true;
}

FD_HD bool torsion_bins_fast(V3 a, V3 b, V3 c, V3 d, float nbin, uint32_t *sbin, uint32_t *cbin) {
    V3 v1 = sub(b, a), v2 = sub(c, b), v3 = sub(d, c);
    V3 r = normalize(cross(v1, v2));
    V3 s = normalize(cross(v2, v3));
    V3 t = normalize(cross(r, normalize(v2)));
    float x = dot(r, s);
    float y = dot(s, t);
    float len = FD_SQRT(FD_ADD(FD_MUL(x, x), FD_MUL(y, y)));
    if (!(len > 1.0e-3f && len < 2.0f)) return false; // degenerate (atan2(0, 0), NaN): exact route
    return fast_bin(FD_DIV(-y, len), nbin, sbin) && fast_bin(FD_DIV(x, len), nbin, cbin);
}

FD_HD bool pair_hash_fast(V3 n1, V3 ca1, V3 cb1, V3 n2, V3 ca2, V3 cb2, uint8_t aa1, uint8_t aa2, float ca_dist,
                          const HashParams &p, uint32_t *out) {
    uint32_t s0, c0, s1, c1, s2, c2;
    {
        V3 v1 = sub(cb1, ca1), v2 = sub(cb2, ca2);
        float d = dot(v1, v2);
        float l1 = norm(v1), l2 = norm(v2);
        float cs = FD_DIV(d, FD_MUL(l1, l2));
        if (!(cs >= -1.0f && cs <= 1.0f)) return false;
        float sn = FD_SQRT(FD_SUB(1.0f, FD_MUL(cs, cs)));
        // sin(acos) >= 0: below 0.01 the closed form loses relative accuracy but the bin is the one of 0+ either way
        if (!fast_bin(sn, p.nbin_angle, &s0) || !fast_bin(cs, p.nbin_angle, &c0)) return false;
    }
    if (!torsion_bins_fast(n1, ca1, cb1, cb2, p.nbin_angle, &s1, &c1)) return false;
    if (!torsion_bins_fast(cb1, cb2, ca2, n2, p.nbin_angle, &s2, &c2)) return false;
    uint32_t res1 = sat_u32((float)aa1), res2 = sat_u32((float)aa2);
    uint32_t ca = discretize(ca_dist, 2.0f, 20.0f, p.nbin_dist);
    uint32_t cb = discretize(dist(cb1, cb2), 2.0f, 20.0f, p.nbin_dist);
    *out = res1 << 25 | res2 << 20 | ca << 16 | cb << 12 | s0 << 10 | c0 << 8 | s1 << 6 | c1 << 4 | s2 << 2 | c2;
    return true;
}

// pair_hash, through the fast route when every bin is decided by it
FD_HD uint32_t pair_hash_auto(V3 n1, V3 ca1, V3 cb1, V3 n2, V3 ca2, V3 cb2, uint8_t aa1, uint8_t aa2, float ca_dist,
                              const HashParams &p) {
    uint32_t h;
    if (pair_hash_fast(n1, ca1, cb1, n2, ca2, cb2, aa1, aa2, ca_dist, p, &h)) return h;
    return pair_hash(n1, ca1, cb1, n2, ca2, cb2, aa1, aa2, ca_dist, p);
}

} // namespace fdg
