// fd_hashtypes.cuh -- the reference's other geometric-hash encodings (`--type`) and `--multiple-bins`, one
// implementation for host and device.
//
// Mirrors, operation by operation (f32, round-to-nearest, no FMA; trigonometry through fd_math.cuh):
//   get_single_feature                     reference src/controller/feature.rs:11-190
//   CompactStructure::get_ca_cb_angle /    src/structure/core.rs:283-300, 319-367
//     get_ppf / get_trrosetta_feature
//   Coordinate::get_ppf, calc_angle,       src/structure/coordinate.rs:93-102, 117-133, 151-162
//     calc_angle_radian
//   perfect_hash / reverse_hash /          src/geometry/pdb_motif.rs:26-103, pdb_motif_sincos.rs:17-150,
//     is_symmetric per encoding              trrosetta.rs:22-162, ppf.rs:15-127, folddisco_angle.rs:24-137,
//                                            folddisco_dist.rs:22-130
//                                            tertiary_interaction.rs:19-151, hybrid.rs:19-189
// The default encoding (PDBTrRosetta, pdb_tr.rs) keeps its own tuned route in fd_geom.cuh; typed_* below covers it
// too so that one code path can serve any `--type`.  TertiaryInteraction and Hybrid also read the C-alpha atoms of the
// neighbouring residues i-1 / i+1 (Nbr) and exist only for residues that are neither first nor last in their structure.
#pragma once
#include "../../include/folddisco_b200.h"
#include "fd_geom.cuh"

namespace fdg {

// values of fd_hash_params.hash_type: the reference's HashType index (src/geometry/core.rs:25-38) + 1; 0 = default
enum : uint32_t {
    HT_DEFAULT = 0,
    HT_PDBMOTIF = 1,
    HT_PDBMOTIFSINCOS = 2,
    HT_TRROSETTA = 3,
    HT_PDBTR = 4,
    HT_PPF = 5,
    HT_TERTIARY = 6,
    HT_HYBRID = 7,
    HT_FDANGLE = 8,
    HT_FDDIST = 9,
};
constexpr int HT_MAX_BINS = 8;

FD_HD bool ht_supported(uint32_t t) { return t <= HT_FDDIST; }
// encodings whose feature reads CA(i-1), CA(i+1), CA(j-1), CA(j+1) and is undefined for the first / last residue
FD_HD bool ht_needs_neighbours(uint32_t t) { return t == HT_TERTIARY || t == HT_HYBRID; }
// TertiaryInteraction is built from C-alpha atoms only: a residue without CB still gets features (feature.rs:112-131)
FD_HD bool ht_needs_cb(uint32_t t) { return t != HT_TERTIARY; }
// HashType::amino_acid_index (feature.rs:260-267): None for the two neighbour encodings -- no amino-acid prefilter
// (retrieve.rs:385-390) and no substitutions (query.rs:301-306) with them
FD_HD bool ht_has_aa_index(uint32_t t) { return !ht_needs_neighbours(t); }
FD_HD uint32_t ht_canon(uint32_t t) { return t == HT_DEFAULT ? (uint32_t)HT_PDBTR : t; }

// HashType::default_dist_bin / default_angle_bin (src/geometry/core.rs:118-147)
FD_HD uint32_t ht_default_dist_bin(uint32_t t) {
    switch (ht_canon(t)) {
        case HT_PDBMOTIF: return 18;
        case HT_PDBTR: return 16;
        case HT_HYBRID: return 16;
        case HT_FDANGLE: return 8;
        case HT_FDDIST: return 32;
        default: return 8; // PDBMotifSinCos, TrRosetta, PointPairFeature, TertiaryInteraction: utils/convert.rs NBIN_DIST
    }
}
FD_HD uint32_t ht_default_angle_bin(uint32_t t) {
    switch (ht_canon(t)) {
        case HT_PDBMOTIF: return 9;
        case HT_PDBTR: return 4;
        case HT_HYBRID: return 4;
        case HT_FDANGLE: return 32;
        case HT_FDDIST: return 16;
        default: return 3; // NBIN_SIN_COS
    }
}

// The whole parameter block of one index: encoding, the (nbin_dist, nbin_angle) pairs a residue pair is hashed
// with (one pair, or the `--multiple-bins` list) and the distance cutoff.
struct TypedParams {
    uint32_t type;   // canonical (never HT_DEFAULT)
    uint32_t n_bins; // >= 1
    uint32_t nbd[HT_MAX_BINS], nba[HT_MAX_BINS];
    float dist_cutoff;
};

// `if nbin_dist == 0 || nbin_angle == 0 { perfect_hash_default } else { perfect_hash(nbin_dist, nbin_angle) }`
// (feature.rs:215-221, query.rs:72-76, retrieve.rs:133-137): either zero selects BOTH defaults.
FD_HD void ht_resolve_single(uint32_t type, uint32_t nbd, uint32_t nba, uint32_t *od, uint32_t *oa) {
    if (nbd == 0 || nba == 0) {
        *od = ht_default_dist_bin(type);
        *oa = ht_default_angle_bin(type);
    } else {
        *od = nbd;
        *oa = nba;
    }
}

// fd_hash_params -> TypedParams.  Returns nullptr, or the reason the parameters are refused.
inline const char *typed_params_from(const fd_hash_params *p, TypedParams *tp) {
    if (!ht_supported(p->hash_type)) return "hash_type: unknown encoding (FD_HASH_* of folddisco_b200.h)";
    if (p->n_multiple_bins > (uint32_t)HT_MAX_BINS) return "at most 8 (dist, angle) pairs in multiple_bins";
    tp->type = ht_canon(p->hash_type);
    tp->dist_cutoff = p->dist_cutoff;
    for (int k = 0; k < HT_MAX_BINS; k++) tp->nbd[k] = tp->nba[k] = 0;
    if (p->n_multiple_bins) {
        tp->n_bins = p->n_multiple_bins;
        for (uint32_t k = 0; k < p->n_multiple_bins; k++) {
            tp->nbd[k] = p->multiple_bins[2 * k];
            tp->nba[k] = p->multiple_bins[2 * k + 1];
        }
    } else {
        tp->n_bins = 1;
        ht_resolve_single(tp->type, p->nbin_dist, p->nbin_angle, &tp->nbd[0], &tp->nba[0]);
    }
    return nullptr;
}
// the tuned PDBTrRosetta single-bin route of fd_geom.cuh (pair_hash_auto, pair table, fused verification kernels)
inline bool ht_default_route(const fd_hash_params *p) {
    return (p->hash_type == HT_DEFAULT || p->hash_type == HT_PDBTR) && p->n_multiple_bins == 0;
}

FD_HD float ht_deg(float rad) { return FD_MUL(rad, 57.2957795130823208767981548141051703f); } // f32::to_degrees

// calc_angle(ca1, cb1, ca2, cb2) in radians (coordinate.rs:117-133)
FD_HD float ht_ca_cb_angle(V3 ca1, V3 cb1, V3 ca2, V3 cb2) {
    V3 v1 = sub(cb1, ca1), v2 = sub(cb2, ca2);
    float d = dot(v1, v2);
    float l1 = norm(v1), l2 = norm(v2);
    return fdm::acosf_exact(FD_DIV(d, FD_MUL(l1, l2)));
}
// calc_angle_radian(a, b, c): angle at b (coordinate.rs:151-162)
FD_HD float ht_angle3(V3 a, V3 b, V3 c) {
    V3 v1 = sub(a, b), v2 = sub(c, b);
    float d = dot(v1, v2);
    float l1 = norm(v1), l2 = norm(v2);
    return fdm::acosf_exact(FD_DIV(d, FD_MUL(l1, l2)));
}

// The distance get_single_feature tests against dist_cutoff for this encoding, with the reference's rounding:
// CA-CA for the PDB* / Folddisco* encodings, CB-CB for TrRosetta, |(cb2 - ca1) - (cb1 - ca1)| for PointPairFeature.
FD_HD float typed_screen_dist(uint32_t type, V3 ca1, V3 cb1, V3 ca2, V3 cb2) {
    switch (type) {
        case HT_TRROSETTA: return dist(cb1, cb2);
        case HT_PPF: return norm(sub(sub(cb2, ca1), sub(cb1, ca1)));
        default: return dist(ca1, ca2);
    }
}

// C-alpha atoms around the two residues and j - i, for the encodings of ht_needs_neighbours (zero otherwise)
struct Nbr {
    V3 ca_pre1, ca_next1, ca_pre2, ca_next2;
    float seq_dist;
};
// map_aa_to_u8_group (utils/convert.rs:85-131) as a function of the amino-acid code: 0 small / aliphatic (A C G P S),
// 1 hydrophobic (I L M F W V), 2 polar (N Q T Y), 3 charged (R D E H K)
FD_HD float ht_aa_group(float aa) {
    const uint32_t a = sat_u32(aa);
    //                         A  R  N  D  C  Q  E  G  H  I  L  K  M  F  P  S  T  W  Y  V
    const uint8_t group[20] = {0, 3, 2, 3, 0, 2, 3, 0, 3, 1, 1, 3, 1, 1, 0, 0, 2, 1, 2, 1};
    return a < 20 ? (float)group[a] : 255.0f;
}

// get_single_feature for residues i -> j whose amino acids are known and whose CB exist (where the encoding needs
// them); d = typed_screen_dist of the pair, already tested (`> dist_cutoff` rejects).  Fills f[0..9).
FD_HD void typed_feature(uint32_t type, V3 n1, V3 ca1, V3 cb1, V3 n2, V3 ca2, V3 cb2, float aa1, float aa2, float d,
                         float *f, const Nbr *nb = nullptr) {
    for (int k = 0; k < 9; k++) f[k] = 0.0f;
    f[0] = aa1;
    f[1] = aa2;
    switch (type) {
        case HT_TERTIARY: { // feature.rs:112-161: seven angles between the CA-trace directions around i and j
            const V3 u1 = normalize(sub(ca1, nb->ca_pre1)), u2 = normalize(sub(nb->ca_next1, ca1));
            const V3 u3 = normalize(sub(ca2, nb->ca_pre2)), u4 = normalize(sub(nb->ca_next2, ca2));
            const V3 u5 = normalize(sub(ca2, ca1));
            f[0] = fdm::acosf_exact(dot(u1, u2));
            f[1] = fdm::acosf_exact(dot(u3, u4));
            f[2] = fdm::acosf_exact(dot(u1, u5));
            f[3] = fdm::acosf_exact(dot(u3, u5));
            f[4] = fdm::acosf_exact(dot(u1, u4));
            f[5] = fdm::acosf_exact(dot(u2, u3));
            f[6] = fdm::acosf_exact(dot(u1, u3));
            f[7] = d;
            f[8] = nb->seq_dist;
            break;
        }
        case HT_HYBRID: // feature.rs:162-186 + core.rs:405-436: amino-acid groups, the PDBTrRosetta geometry, two CA-trace torsions
            pair_feature(n1, ca1, cb1, n2, ca2, cb2, ht_aa_group(aa1), ht_aa_group(aa2), d, f);
            f[7] = torsion(nb->ca_pre1, n1, ca1, nb->ca_next1);
            f[8] = torsion(nb->ca_pre2, n2, ca2, nb->ca_next2);
            break;
        case HT_PDBMOTIF:
        case HT_PDBMOTIFSINCOS: {
            f[2] = d;
            f[3] = dist(cb1, cb2);
            const float a = ht_ca_cb_angle(ca1, cb1, ca2, cb2);
            f[4] = type == HT_PDBMOTIF ? ht_deg(a) : a;
            break;
        }
        case HT_TRROSETTA:
            f[2] = d;
            f[3] = torsion(ca1, cb1, cb2, ca2);
            f[4] = torsion(n1, ca1, cb1, cb2);
            f[5] = torsion(cb1, cb2, ca2, n2);
            f[6] = ht_angle3(ca1, cb1, cb2);
            f[7] = ht_angle3(cb1, cb2, ca2);
            break;
        case HT_PPF: {
            const V3 r1 = sub(cb1, ca1), r2 = sub(cb2, ca1);
            const V3 u1 = normalize(r1), u2 = normalize(r2);
            const V3 nd = normalize(sub(r2, r1));
            f[2] = d;
            f[3] = fdm::acosf_exact(dot(u1, nd));
            f[4] = fdm::acosf_exact(dot(u2, nd));
            f[5] = fdm::acosf_exact(dot(u1, u2));
            break;
        }
        default: // PDBTrRosetta, FolddiscoAngle, FolddiscoDist share get_pdb_tr_feature
            pair_feature(n1, ca1, cb1, n2, ca2, cb2, aa1, aa2, d, f);
            break;
    }
}

FD_HD uint32_t ht_sin_bin(float a, float nb) {
    float s, c;
    fdm::sincosf_exact(a, &s, &c);
    return discretize(s, -1.0f, 1.0f, nb);
}
FD_HD uint32_t ht_cos_bin(float a, float nb) {
    float s, c;
    fdm::sincosf_exact(a, &s, &c);
    return discretize(c, -1.0f, 1.0f, nb);
}
FD_HD float ht_clamp(uint32_t n, uint32_t mx, float dflt) { return n > mx ? (float)mx : (n == 0 ? dflt : (float)n); }

// HashValue::perfect_hash(feature, nbin_dist, nbin_angle) of the encoding, with each file's own clamping.  Shifted
// fields are OR-ed in unmasked, like the reference.
FD_HD uint32_t typed_hash(uint32_t type, const float *f, uint32_t nbd_u, uint32_t nba_u) {
    const float PI_F = 3.14159274101257324f;
    const uint32_t r1 = sat_u32(f[0]), r2 = sat_u32(f[1]);
    switch (type) {
        case HT_PDBMOTIF: {
            const float nbd = ht_clamp(nbd_u, 32, 18.0f), nba = ht_clamp(nba_u, 32, 9.0f);
            const uint32_t ca = discretize(f[2], 2.0f, 20.0f, nbd), cb = discretize(f[3], 2.0f, 20.0f, nbd);
            const uint32_t an = discretize(f[4], 0.0f, 180.0f, nba);
            return r1 << 20 | r2 << 15 | ca << 10 | cb << 5 | an;
        }
        case HT_PDBMOTIFSINCOS: {
            const float nbd = ht_clamp(nbd_u, 16, 8.0f), nba = ht_clamp(nba_u, 16, 3.0f);
            const uint32_t ca = discretize(f[2], 2.0f, 20.0f, nbd), cb = discretize(f[3], 2.0f, 20.0f, nbd);
            return r1 << 21 | r2 << 16 | ca << 12 | cb << 8 | ht_sin_bin(f[4], nba) << 4 | ht_cos_bin(f[4], nba);
        }
        case HT_TRROSETTA: { // no zero handling inside _perfect_hash (trrosetta.rs:56-58): callers resolve zeros
            const float nbd = nbd_u > 8 ? 8.0f : (float)nbd_u, nba = nba_u > 4 ? 4.0f : (float)nba_u;
            uint32_t h = (r1 * 20u + r2) << 23 | discretize(f[2], 2.0f, 20.0f, nbd) << 20;
            for (int k = 0; k < 5; k++)
                h |= ht_sin_bin(f[3 + k], nba) << (18 - 4 * k) | ht_cos_bin(f[3 + k], nba) << (16 - 4 * k);
            return h;
        }
        case HT_PPF: {
            const float nbd = ht_clamp(nbd_u, 16, 8.0f), nba = ht_clamp(nba_u, 8, 3.0f);
            uint32_t h = r1 << 27 | r2 << 22 | discretize(f[2], 2.0f, 20.0f, nbd) << 18;
            for (int k = 0; k < 3; k++)
                h |= ht_sin_bin(f[3 + k], nba) << (15 - 6 * k) | ht_cos_bin(f[3 + k], nba) << (12 - 6 * k);
            return h;
        }
        case HT_TERTIARY: { // tertiary_interaction.rs:21-83: seven cos bins, the CA distance, the clamped sequence distance
            const float nbd = ht_clamp(nbd_u, 16, 8.0f), nba = ht_clamp(nba_u, 8, 3.0f);
            uint32_t h = 0;
            for (int k = 0; k < 7; k++) h |= ht_cos_bin(f[k], nba) << (26 - 3 * k);
            const uint32_t sd = f[8] < -4.0f ? 0u : (f[8] > 4.0f ? 8u : sat_u32(f[8]) + 4u);
            return h | discretize(f[7], 2.0f, 20.0f, nbd) << 4 | sd;
        }
        case HT_HYBRID: { // hybrid.rs:21-96
            const float nbd = ht_clamp(nbd_u, 16, 16.0f), nba = ht_clamp(nba_u, 4, 4.0f);
            uint32_t h = r1 << 30 | r2 << 28 | discretize(f[2], 2.0f, 20.0f, nbd) << 24 | discretize(f[3], 2.0f, 20.0f, nbd) << 20;
            for (int k = 0; k < 5; k++)
                h |= ht_sin_bin(f[4 + k], nba) << (18 - 4 * k) | ht_cos_bin(f[4 + k], nba) << (16 - 4 * k);
            return h;
        }
        case HT_FDANGLE: {
            const float nbd = ht_clamp(nbd_u, 8, 8.0f), nba = ht_clamp(nba_u, 32, 32.0f);
            const uint32_t ca = discretize(f[2], 2.0f, 20.0f, nbd), cb = discretize(f[3], 2.0f, 20.0f, nbd);
            const uint32_t an = discretize(f[4], 0.0f, PI_F, fminf(nba, 32.0f));
            const uint32_t p1 = discretize(f[5], -PI_F, PI_F, nba), p2 = discretize(f[6], -PI_F, PI_F, nba);
            return (r1 * 20u + r2) << 21 | ca << 18 | cb << 15 | an << 10 | p1 << 5 | p2;
        }
        case HT_FDDIST: {
            const float nbd = ht_clamp(nbd_u, 32, 32.0f), nba = ht_clamp(nba_u, 16, 16.0f);
            const uint32_t ca = discretize(f[2], 2.0f, 20.0f, nbd), cb = discretize(f[3], 2.0f, 20.0f, nbd);
            const uint32_t an = discretize(f[4], 0.0f, PI_F, fminf(nba, 8.0f));
            const uint32_t p1 = discretize(f[5], -PI_F, PI_F, nba), p2 = discretize(f[6], -PI_F, PI_F, nba);
            return (r1 * 20u + r2) << 21 | ca << 16 | cb << 11 | an << 8 | p1 << 4 | p2;
        }
        default: {
            HashParams p;
            p.nbin_dist = ht_clamp(nbd_u, 16, 16.0f);
            p.nbin_angle = ht_clamp(nba_u, 4, 4.0f);
            p.dist_cutoff = 0.f;
            return perfect_hash(f, p);
        }
    }
}

FD_HD float ht_cont(uint32_t v, float mn, float mx, float nb) { // continuize_u32_value_into_f32 (convert.rs:38-42)
    return FD_ADD(FD_MUL((float)v, FD_DIV(FD_SUB(mx, mn), FD_SUB(nb, 1.0f))), mn);
}

// (res1, res2) as HashValue::reverse_hash_default decodes them (prefilter_amino_acid, retrieve.rs:563-602)
FD_HD void typed_hash_aa(uint32_t type, uint32_t h, uint32_t *aa1, uint32_t *aa2) {
    switch (type) {
        case HT_PDBMOTIF: *aa1 = (h >> 20) & 31u, *aa2 = (h >> 15) & 31u; break;
        case HT_PDBMOTIFSINCOS: *aa1 = (h >> 21) & 31u, *aa2 = (h >> 16) & 31u; break;
        case HT_TRROSETTA: *aa1 = ((h >> 23) & 511u) / 20u, *aa2 = ((h >> 23) & 511u) % 20u; break;
        case HT_PPF: *aa1 = (h >> 27) & 31u, *aa2 = (h >> 22) & 31u; break;
        case HT_FDANGLE:
        case HT_FDDIST: *aa1 = ((h >> 21) & 511u) / 20u, *aa2 = ((h >> 21) & 511u) % 20u; break;
        default: *aa1 = (h >> 25) & 31u, *aa2 = (h >> 20) & 31u; break;
    }
}

// HashValue::is_symmetric of the encoding: always through reverse_hash_default, i.e. with the DEFAULT bin counts
// whatever the index was built with.  Host only (the verification receives the verdicts as a per-hash table).
inline bool typed_is_symmetric(uint32_t type, uint32_t h) {
    auto ang = [](uint32_t sb, uint32_t cb, float nb) {
        return ht_deg(fdm::atan2f_exact(ht_cont(sb, -1.0f, 1.0f, nb), ht_cont(cb, -1.0f, 1.0f, nb)));
    };
    if (type == HT_TERTIARY) return false; // tertiary_interaction.rs:145-150
    if (type == HT_HYBRID)                 // hybrid.rs:185-189: equal groups and phi1 == phi2 (default 4 bins)
        return ((h >> 30) & 3u) == ((h >> 28) & 3u) &&
               ang((h >> 14) & 3u, (h >> 12) & 3u, 4.0f) == ang((h >> 10) & 3u, (h >> 8) & 3u, 4.0f);
    uint32_t a1, a2;
    typed_hash_aa(type, h, &a1, &a2);
    if (a1 != a2) return false;
    const float PI_F = 3.14159274101257324f;
    switch (type) {
        case HT_PDBMOTIF:
        case HT_PDBMOTIFSINCOS: return true; // res1 == res2 only
        case HT_TRROSETTA: // theta1 == theta2 and phi1 == phi2 (default 3 sin / cos bins, 2-bit fields)
            return ang((h >> 14) & 3u, (h >> 12) & 3u, 3.0f) == ang((h >> 10) & 3u, (h >> 8) & 3u, 3.0f) &&
                   ang((h >> 6) & 3u, (h >> 4) & 3u, 3.0f) == ang((h >> 2) & 3u, h & 3u, 3.0f);
        case HT_PPF: // values[3] == values[4]: the first two angles
            return ang((h >> 15) & 7u, (h >> 12) & 7u, 3.0f) == ang((h >> 9) & 7u, (h >> 6) & 7u, 3.0f);
        case HT_FDANGLE:
            return ht_deg(ht_cont((h >> 5) & 31u, -PI_F, PI_F, 32.0f)) == ht_deg(ht_cont(h & 31u, -PI_F, PI_F, 32.0f));
        case HT_FDDIST:
            return ht_deg(ht_cont((h >> 4) & 15u, -PI_F, PI_F, 16.0f)) == ht_deg(ht_cont(h & 15u, -PI_F, PI_F, 16.0f));
        default: // pdb_tr.rs:158-162
            return ang((h >> 6) & 3u, (h >> 4) & 3u, 4.0f) == ang((h >> 2) & 3u, h & 3u, 4.0f);
    }
}

// which feature slots `-d` / `-a` perturb (HashType::dist_index / angle_index, feature.rs:269-289)
inline int typed_dist_index(uint32_t type, int *idx) {
    switch (type) {
        case HT_TRROSETTA:
        case HT_PPF: idx[0] = 2; return 1;
        case HT_TERTIARY: idx[0] = 7; return 1;
        default: idx[0] = 2, idx[1] = 3; return 2;
    }
}
inline int typed_angle_index(uint32_t type, int *idx) {
    switch (type) {
        case HT_PDBMOTIF:
        case HT_PDBMOTIFSINCOS: idx[0] = 4; return 1;
        case HT_TRROSETTA:
            for (int k = 0; k < 5; k++) idx[k] = 3 + k;
            return 5;
        case HT_PPF: idx[0] = 3, idx[1] = 4, idx[2] = 5; return 3;
        case HT_TERTIARY:
            for (int k = 0; k < 7; k++) idx[k] = k;
            return 7;
        case HT_HYBRID:
            for (int k = 0; k < 5; k++) idx[k] = 4 + k;
            return 5;
        default: idx[0] = 4, idx[1] = 5, idx[2] = 6; return 3;
    }
}

} // namespace fdg
