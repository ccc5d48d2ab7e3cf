// fd_oracle.cpp -- TEST INFRASTRUCTURE.  CPU restatement of the reference algorithm
// (steineggerlab/folddisco @ 9375a2d) for the hot path; see fd_oracle.h.
// Every function cites the reference file:line it follows.  Compile with
//   g++ -O2 -std=c++17 -ffp-contract=off   (Rust never contracts a*b+c into an FMA)
// Parity status: pinned by the reference's own KATs (tests/test_oracle_golden.py);
// last-ulp libm behaviour and FxHashMap iteration order are unpinned.
#include "fd_oracle.h"
#include "fd_oracle_math.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

int fdo_math_use_libm = 0;

namespace {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------

// Rust `f32 as u32`: saturating, NaN -> 0.
inline uint32_t sat_u32(float v) {
    if (!(v > 0.0f)) return 0u; // negative, -0, NaN
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

struct V3 {
    float x, y, z;
};
// src/structure/coordinate.rs:23-83
inline V3 vsub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 vadd(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 vscale(V3 a, float f) { return {a.x * f, a.y * f, a.z * f}; }
inline float vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 vcross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float vnorm(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V3 vnormalize(V3 a) {
    float n = vnorm(a);
    return {a.x / n, a.y / n, a.z / n};
}
// coordinate.rs:109-115
inline float calc_distance(V3 a, V3 b) {
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return sqrtf(dx * dx + dy * dy + dz * dz);
}
// coordinate.rs:117-133 (powf(2.0) == x*x exactly)
inline float calc_angle(V3 a, V3 b, V3 c, V3 d) {
    float v1x = b.x - a.x, v1y = b.y - a.y, v1z = b.z - a.z;
    float v2x = d.x - c.x, v2y = d.y - c.y, v2z = d.z - c.z;
    float dot = v1x * v2x + v1y * v2y + v1z * v2z;
    float l1 = sqrtf(v1x * v1x + v1y * v1y + v1z * v1z);
    float l2 = sqrtf(v2x * v2x + v2y * v2y + v2z * v2z);
    float cs = dot / (l1 * l2);
    return fdo_acosf(cs);
}
// coordinate.rs:203-215
inline float calc_torsion_radian(V3 a, V3 b, V3 c, V3 d) {
    V3 v1 = vsub(b, a), v2 = vsub(c, b), v3 = vsub(d, c);
    V3 r = vnormalize(vcross(v1, v2));
    V3 s = vnormalize(vcross(v2, v3));
    V3 t = vnormalize(vcross(r, vnormalize(v2)));
    float x = vdot(r, s);
    float y = vdot(s, t);
    return -fdo_atan2f(y, x);
}
// coordinate.rs:167-186
inline V3 approx_cb(V3 ca, V3 n, V3 c) {
    const float CA_CB_DIST = 1.5336f;
    V3 v1 = vnormalize(vsub(c, ca));
    V3 v2 = vnormalize(vsub(n, ca));
    V3 b1 = vadd(v2, vscale(v1, 1.0f / 3.0f));
    V3 b2 = vcross(v1, b1);
    V3 u1 = vnormalize(b1);
    V3 u2 = vnormalize(b2);
    V3 v4 = vsub(vscale(u1, -1.0f / 2.0f), vscale(u2, sqrtf(3.0f) / 2.0f));
    v4 = vscale(v4, sqrtf(8.0f) / 3.0f);
    v4 = vadd(v4, vscale(v1, -1.0f / 3.0f));
    return vadd(ca, vscale(v4, CA_CB_DIST));
}

// src/utils/convert.rs:53-81
uint8_t map_aa_to_u8(const uint8_t *aa) {
    struct E {
        const char *n;
        uint8_t v;
    };
    static const E tab[] = {
        {"ALA", 0},  {"ABA", 0},  {"ORN", 0},  {"DAL", 0},  {"AIB", 0},  {"ALC", 0},  {"MDO", 0},  {"MAA", 0},
        {"DAB", 0},  {"ARG", 1},  {"DAR", 1},  {"CIR", 1},  {"AGM", 1},  {"ASN", 2},  {"DSG", 2},  {"MEN", 2},
        {"SNN", 2},  {"ASP", 3},  {"0TD", 3},  {"DAS", 3},  {"IAS", 3},  {"PHD", 3},  {"BFD", 3},  {"ASX", 3},
        {"CYS", 4},  {"CSO", 4},  {"CSD", 4},  {"CME", 4},  {"OCS", 4},  {"CAS", 4},  {"CSX", 4},  {"CSS", 4},
        {"YCM", 4},  {"DCY", 4},  {"SMC", 4},  {"SCH", 4},  {"SCY", 4},  {"CAF", 4},  {"SNC", 4},  {"SEC", 4},
        {"GLN", 5},  {"DGN", 5},  {"CRQ", 5},  {"MEQ", 5},  {"GLU", 6},  {"PCA", 6},  {"DGL", 6},  {"CGU", 6},
        {"FGA", 6},  {"B3E", 6},  {"GLX", 6},  {"GLY", 7},  {"CR2", 7},  {"SAR", 7},  {"GHP", 7},  {"GL3", 7},
        {"HIS", 8},  {"HIC", 8},  {"DHI", 8},  {"NEP", 8},  {"CR8", 8},  {"MHS", 8},  {"ILE", 9},  {"DIL", 9},
        {"LEU", 10}, {"DLE", 10}, {"NLE", 10}, {"MLE", 10}, {"MK8", 10}, {"LYS", 11}, {"KCX", 11}, {"LLP", 11},
        {"MLY", 11}, {"M3L", 11}, {"ALY", 11}, {"MLZ", 11}, {"DLY", 11}, {"KPI", 11}, {"PYL", 11}, {"MET", 12},
        {"MSE", 12}, {"FME", 12}, {"NRQ", 12}, {"CXM", 12}, {"SME", 12}, {"MHO", 12}, {"MED", 12}, {"PHE", 13},
        {"DPN", 13}, {"PHI", 13}, {"MEA", 13}, {"PHL", 13}, {"PRO", 14}, {"HYP", 14}, {"DPR", 14}, {"SER", 15},
        {"CSH", 15}, {"SEP", 15}, {"DSN", 15}, {"SAC", 15}, {"GYS", 15}, {"DHA", 15}, {"OAS", 15}, {"THR", 16},
        {"TPO", 16}, {"CRO", 16}, {"DTH", 16}, {"BMT", 16}, {"CRF", 16}, {"TRP", 17}, {"DTR", 17}, {"TRQ", 17},
        {"TOX", 17}, {"0AF", 17}, {"TYR", 18}, {"PTR", 18}, {"TYS", 18}, {"TPQ", 18}, {"DTY", 18}, {"OMY", 18},
        {"VAL", 19}, {"DVA", 19}, {"MVA", 19}, {"FVA", 19},
    };
    for (const E &e : tab)
        if (e.n[0] == (char)aa[0] && e.n[1] == (char)aa[1] && e.n[2] == (char)aa[2]) return e.v;
    return 255;
}
// convert.rs:166-190
const char *map_u8_to_aa(uint8_t aa) {
    static const char *names[20] = {"ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE",
                                    "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL"};
    return aa < 20 ? names[aa] : "UNK";
}

} // namespace

// ---------------------------------------------------------------------------------------------
// Structure / CompactStructure
// ---------------------------------------------------------------------------------------------

struct fdo_structure { // src/structure/core.rs:9-16 + atom.rs:73-81
    std::vector<float> x, y, z, b_factor;
    std::vector<std::array<uint8_t, 4>> atom_name;
    std::vector<std::array<uint8_t, 3>> res_name;
    std::vector<uint64_t> res_serial;
    std::vector<uint8_t> chain;
    std::vector<uint8_t> chains;
    uint64_t num_residues = 0;
    uint8_t rec_chain = ' ';
    uint64_t rec_serial = 0;
    // core.rs:29-43
    void update(float ax, float ay, float az, const uint8_t *an, uint8_t ch, const uint8_t *rn, uint64_t rs,
                float bf) {
        if (rec_chain != ch) {
            chains.push_back(ch);
            rec_chain = ch;
        }
        if (rec_serial != rs) {
            num_residues++;
            rec_serial = rs;
        }
        x.push_back(ax);
        y.push_back(ay);
        z.push_back(az);
        atom_name.push_back({an[0], an[1], an[2], an[3]});
        chain.push_back(ch);
        res_name.push_back({rn[0], rn[1], rn[2]});
        res_serial.push_back(rs);
        b_factor.push_back(bf);
    }
};

struct fdo_compact { // core.rs:55-67
    std::vector<V3> n, ca, cb;
    std::vector<uint8_t> cb_valid;
    std::vector<std::array<uint8_t, 3>> res_name;
    std::vector<uint8_t> aa; // map_aa_to_u8(res_name), cached
    std::vector<uint8_t> chain;
    std::vector<uint64_t> serial;
    std::vector<float> b_factor;
    std::vector<uint8_t> chains;
    size_t nres() const { return serial.size(); }
};

namespace {

std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
// Rust str::parse::<f32>() on a trimmed field: decimal, correctly rounded. strtof is too.
bool parse_f32(const std::string &field, float *out) {
    std::string t = trim(field);
    if (t.empty()) return false;
    // Rust accepts [+-]digits[.digits][e[+-]digits], "inf", "infinity", "nan" (any case); no hex, no spaces.
    for (char ch : t)
        if (ch == 'x' || ch == 'X' || ch == '(' || isspace((unsigned char)ch)) return false;
    char *end = nullptr;
    float v = strtof(t.c_str(), &end);
    if (end == t.c_str() || *end != '\0') return false;
    *out = v;
    return true;
}
bool parse_u64(const std::string &field, uint64_t *out) {
    std::string t = trim(field);
    if (t.empty()) return false;
    size_t i = 0;
    if (t[0] == '+') i = 1;
    if (i >= t.size()) return false;
    uint64_t v = 0;
    for (; i < t.size(); i++) {
        if (t[i] < '0' || t[i] > '9') return false;
        uint64_t nv = v * 10 + (uint64_t)(t[i] - '0');
        if (nv / 10 != v) return false; // overflow
        v = nv;
    }
    *out = v;
    return true;
}

// src/structure/io/parser.rs:3-55 + pdb.rs:37-76
bool read_pdb(const char *path, fdo_structure *st) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    std::string line;
    int model = 0;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (model > 1) break;
        if (line.size() < 6) continue;
        if (line.compare(0, 6, "MODEL ") == 0) {
            model++;
        } else if (line.compare(0, 6, "ATOM  ") == 0) {
            if (line.size() < 54) continue; // the reference would panic on the slice; treat as unparsable
            float x, y, z, bf = 1.0f;
            uint64_t aser, rser;
            if (!parse_f32(line.substr(30, 8), &x)) continue;
            if (!parse_f32(line.substr(38, 8), &y)) continue;
            if (!parse_f32(line.substr(46, 8), &z)) continue;
            if (!parse_u64(line.substr(6, 5), &aser)) continue;
            if (!parse_u64(line.substr(22, 4), &rser)) continue;
            if (line.size() >= 66) {
                if (!parse_f32(line.substr(60, 6), &bf)) continue;
            }
            st->update(x, y, z, (const uint8_t *)line.data() + 12, (uint8_t)line[21],
                       (const uint8_t *)line.data() + 17, rser, bf);
        }
    }
    return true;
}

// src/structure/core.rs:70-214 with quirks Q1..Q5 of SURVEY 8a.
fdo_compact *build_compact(const fdo_structure &o) {
    fdo_compact *c = new fdo_compact();
    c->chains = o.chains;
    const size_t na = o.x.size();
    bool has_prev = false;
    uint64_t prev_serial = 0;
    const std::array<uint8_t, 3> *prev_name = nullptr;
    bool has_n = false, has_ca = false, has_cb = false, has_c = false;
    V3 n{}, ca{}, cb{}, cc{};
    auto is = [&](size_t i, const char *nm) { return memcmp(o.atom_name[i].data(), nm, 4) == 0; };
    for (size_t idx = 0; idx < na; idx++) {
        if (!has_prev || prev_serial != o.res_serial[idx] || idx == na - 1) {
            if (has_n && has_ca) {
                bool push = true;
                V3 cbv{};
                uint8_t cbok = 1;
                if (has_cb) {
                    cbv = cb;
                } else {
                    // gly_c is never set in the reference (the is_c branch shadows it, core.rs:173-189),
                    // so the GLY arm (core.rs:138-146) is dead and every CB-less residue uses the
                    // most recent backbone C seen so far.
                    if (has_c) cbv = approx_cb(ca, n, cc);
                    else cbok = 0;
                }
                if (push) {
                    c->n.push_back(n);
                    c->ca.push_back(ca);
                    c->cb.push_back(cbv);
                    c->cb_valid.push_back(cbok);
                    c->serial.push_back(prev_serial);
                    c->res_name.push_back(*prev_name);
                    c->aa.push_back(map_aa_to_u8(prev_name->data()));
                    c->chain.push_back(o.chain[idx]);       // Q2: atom idx, not the residue's own atom
                    c->b_factor.push_back(o.b_factor[idx]); // Q2
                }
            }
            has_ca = has_cb = has_n = false; // c is NOT reset (Q3)
            has_prev = true;
            prev_serial = o.res_serial[idx];
            prev_name = &o.res_name[idx];
        }
        V3 p{o.x[idx], o.y[idx], o.z[idx]};
        bool gly = memcmp(o.res_name[idx].data(), "GLY", 3) == 0;
        if (is(idx, " CA ")) {
            ca = p;
            has_ca = true;
        } else if (is(idx, " CB ")) {
            cb = p;
            has_cb = true;
        } else if (is(idx, " C  ")) {
            cc = p;
            has_c = true;
        } else if (is(idx, " N  ") && !gly) {
            n = p;
            has_n = true;
        } else if (gly) {
            if (is(idx, " N  ")) {
                n = p;
                has_n = true;
            }
        }
    }
    return c;
}

// Encoding selected for the whole oracle (fdo_set_hash_type): the reference's HashType index (geometry/core.rs:25-38)
// 0 PDBMotif, 1 PDBMotifSinCos, 2 TrRosetta, 3 PDBTrRosetta (default), 4 PointPairFeature, 7 FolddiscoAngle,
// 8 FolddiscoDist; and the --multiple-bins list (empty = none).
int g_hash_type = 3;
std::vector<std::pair<uint32_t, uint32_t>> g_multi_bins;

// coordinate.rs:151-162
inline float calc_angle_radian(V3 a, V3 b, V3 c) {
    float v1x = a.x - b.x, v1y = a.y - b.y, v1z = a.z - b.z;
    float v2x = c.x - b.x, v2y = c.y - b.y, v2z = c.z - b.z;
    float dot = v1x * v2x + v1y * v2y + v1z * v2z;
    float l1 = sqrtf(v1x * v1x + v1y * v1y + v1z * v1z);
    float l2 = sqrtf(v2x * v2x + v2y * v2y + v2z * v2z);
    return fdo_acosf(dot / (l1 * l2));
}
// f32::to_degrees multiplies by the literal 57.29577951...f32 (not by 180 / PI evaluated in f32)
inline float to_degrees(float r) { return r * 57.2957795130823208767981548141051703f; }

// get_single_feature (feature.rs:11-190) for the encodings over the N / CA / CB atoms of the two residues
// map_aa_to_u8_group (convert.rs:85-131), by amino-acid code
uint8_t aa_group(uint8_t aa) {
    switch (aa) {
        case 0: case 4: case 7: case 14: case 15: return 0; // ALA CYS GLY PRO SER (small / aliphatic)
        case 9: case 10: case 12: case 13: case 17: case 19: return 1; // ILE LEU MET PHE TRP VAL (hydrophobic)
        case 2: case 5: case 16: case 18: return 2; // ASN GLN THR TYR (polar)
        case 1: case 3: case 6: case 8: case 11: return 3; // ARG ASP GLU HIS LYS (charged)
        default: return 255;
    }
}

bool pair_feature(const fdo_compact &c, size_t i, size_t j, float cutoff, float *f) {
    if (i == j) return false;
    uint8_t r1 = c.aa[i], r2 = c.aa[j];
    if (r1 == 255 || r2 == 255) return false;
    if (g_hash_type != 5 && (!c.cb_valid[i] || !c.cb_valid[j])) return false; // TertiaryInteraction reads CA only
    for (int k = 0; k < 9; k++) f[k] = 0.0f;
    f[0] = (float)r1;
    f[1] = (float)r2;
    switch (g_hash_type) {
        case 5: { // TertiaryInteraction (feature.rs:112-161)
            const size_t n = c.nres();
            if (i == 0 || j == 0 || i == n - 1 || j == n - 1) return false;
            float ca_dist = calc_distance(c.ca[i], c.ca[j]);
            if (ca_dist > cutoff) return false;
            V3 u1 = vnormalize(vsub(c.ca[i], c.ca[i - 1])), u2 = vnormalize(vsub(c.ca[i + 1], c.ca[i]));
            V3 u3 = vnormalize(vsub(c.ca[j], c.ca[j - 1])), u4 = vnormalize(vsub(c.ca[j + 1], c.ca[j]));
            V3 u5 = vnormalize(vsub(c.ca[j], c.ca[i]));
            f[0] = fdo_acosf(vdot(u1, u2)); // phi_12
            f[1] = fdo_acosf(vdot(u3, u4)); // phi_34
            f[2] = fdo_acosf(vdot(u1, u5)); // phi_15
            f[3] = fdo_acosf(vdot(u3, u5)); // phi_35
            f[4] = fdo_acosf(vdot(u1, u4)); // phi_14
            f[5] = fdo_acosf(vdot(u2, u3)); // phi_23
            f[6] = fdo_acosf(vdot(u1, u3)); // phi_13
            f[7] = ca_dist;
            f[8] = (float)j - (float)i;
            return true;
        }
        case 6: { // Hybrid (feature.rs:162-186, core.rs:405-436)
            const size_t n = c.nres();
            if (i == 0 || j == 0 || i == n - 1 || j == n - 1) return false;
            float ca_dist = calc_distance(c.ca[i], c.ca[j]);
            if (ca_dist > cutoff) return false;
            f[0] = (float)aa_group(r1);
            f[1] = (float)aa_group(r2);
            f[2] = ca_dist;
            f[3] = calc_distance(c.cb[i], c.cb[j]);
            f[4] = calc_angle(c.ca[i], c.cb[i], c.ca[j], c.cb[j]);
            f[5] = calc_torsion_radian(c.n[i], c.ca[i], c.cb[i], c.cb[j]);
            f[6] = calc_torsion_radian(c.cb[i], c.cb[j], c.ca[j], c.n[j]);
            f[7] = calc_torsion_radian(c.ca[i - 1], c.n[i], c.ca[i], c.ca[i + 1]);
            f[8] = calc_torsion_radian(c.ca[j - 1], c.n[j], c.ca[j], c.ca[j + 1]);
            return true;
        }
        case 0:   // PDBMotif: ca_dist, cb_dist, CA-CB angle in degrees (feature.rs:26-45)
        case 1: { // PDBMotifSinCos: the same with the angle in radians (feature.rs:47-66)
            float ca_dist = calc_distance(c.ca[i], c.ca[j]);
            if (ca_dist > cutoff) return false;
            f[2] = ca_dist;
            f[3] = calc_distance(c.cb[i], c.cb[j]);
            float a = calc_angle(c.ca[i], c.cb[i], c.ca[j], c.cb[j]);
            f[4] = g_hash_type == 0 ? to_degrees(a) : a;
            return true;
        }
        case 2: { // TrRosetta (core.rs:319-343): cutoff on the CB distance
            float cb_dist = calc_distance(c.cb[i], c.cb[j]);
            if (cb_dist > cutoff) return false;
            f[2] = cb_dist;
            f[3] = calc_torsion_radian(c.ca[i], c.cb[i], c.cb[j], c.ca[j]);
            f[4] = calc_torsion_radian(c.n[i], c.ca[i], c.cb[i], c.cb[j]);
            f[5] = calc_torsion_radian(c.cb[i], c.cb[j], c.ca[j], c.n[j]);
            f[6] = calc_angle_radian(c.ca[i], c.cb[i], c.cb[j]);
            f[7] = calc_angle_radian(c.cb[i], c.cb[j], c.ca[j]);
            return true;
        }
        case 4: { // PointPairFeature (core.rs:302-317, coordinate.rs:93-102)
            V3 rel1 = vsub(c.cb[i], c.ca[i]), rel2 = vsub(c.cb[j], c.ca[i]);
            V3 n1 = vnormalize(rel1), n2 = vnormalize(rel2);
            V3 d = vsub(rel2, rel1);
            V3 nd = vnormalize(d);
            float dn = vnorm(d);
            if (dn > cutoff) return false;
            f[2] = dn;
            f[3] = fdo_acosf(vdot(n1, nd));
            f[4] = fdo_acosf(vdot(n2, nd));
            f[5] = fdo_acosf(vdot(n1, n2));
            return true;
        }
        default: { // PDBTrRosetta / FolddiscoAngle / FolddiscoDist: core.rs:378-403 + feature.rs:84-99
            float ca_dist = calc_distance(c.ca[i], c.ca[j]);
            if (ca_dist > cutoff) return false;
            f[2] = ca_dist;
            f[3] = calc_distance(c.cb[i], c.cb[j]);
            f[4] = calc_angle(c.ca[i], c.cb[i], c.ca[j], c.cb[j]);
            f[5] = calc_torsion_radian(c.n[i], c.ca[i], c.cb[i], c.cb[j]);
            f[6] = calc_torsion_radian(c.cb[i], c.cb[j], c.ca[j], c.n[j]);
            return true;
        }
    }
}

// convert.rs:32-36
inline uint32_t discretize(float val, float mn, float mx, float nbin) {
    float cont_f = (mx - mn) / (nbin - 1.0f);
    float disc_f = 1.0f / cont_f;
    return sat_u32((val - mn) * disc_f + 0.5f);
}
inline float continuize(uint32_t v, float mn, float mx, float nbin) {
    float cont_f = (mx - mn) / (nbin - 1.0f);
    return (float)v * cont_f + mn;
}

// geometry/pdb_tr.rs:21-75 (nbin 0 -> default 16/4; >16 / >4 clamp)
uint32_t perfect_hash_pdbtr(const float *f, uint32_t nbd, uint32_t nba) {
    float nbin_dist = nbd > 16 ? 16.0f : (nbd == 0 ? 16.0f : (float)nbd);
    float nbin_angle = nba > 4 ? 4.0f : (nba == 0 ? 4.0f : (float)nba);
    uint32_t res1 = sat_u32(f[0]), res2 = sat_u32(f[1]);
    uint32_t ca = discretize(f[2], 2.0f, 20.0f, nbin_dist);
    uint32_t cb = discretize(f[3], 2.0f, 20.0f, nbin_dist);
    uint32_t s0 = discretize(fdo_sinf(f[4]), -1.0f, 1.0f, nbin_angle);
    uint32_t c0 = discretize(fdo_cosf(f[4]), -1.0f, 1.0f, nbin_angle);
    uint32_t s1 = discretize(fdo_sinf(f[5]), -1.0f, 1.0f, nbin_angle);
    uint32_t c1 = discretize(fdo_cosf(f[5]), -1.0f, 1.0f, nbin_angle);
    uint32_t s2 = discretize(fdo_sinf(f[6]), -1.0f, 1.0f, nbin_angle);
    uint32_t c2 = discretize(fdo_cosf(f[6]), -1.0f, 1.0f, nbin_angle);
    // Rust `<<` on u32 with an in-range shift never panics; bits shifted past 31 are dropped.
    return res1 << 25 | res2 << 20 | ca << 16 | cb << 12 | s0 << 10 | c0 << 8 | s1 << 6 | c1 << 4 | s2 << 2 | c2;
}

const float PI_F32 = 3.14159274101257324f;

// HashValue::perfect_hash(feature, nbin_dist, nbin_angle) of the selected encoding, each with its file's clamping
uint32_t perfect_hash_raw(const float *f, uint32_t nbd, uint32_t nba) {
    uint32_t res1 = sat_u32(f[0]), res2 = sat_u32(f[1]);
    switch (g_hash_type) {
        case 0: { // pdb_motif.rs:26-50 (defaults 18 / 9, clamp 32 / 32, angle in degrees over [0, 180])
            float nd = nbd > 32 ? 32.0f : (nbd == 0 ? 18.0f : (float)nbd);
            float na = nba > 32 ? 32.0f : (nba == 0 ? 9.0f : (float)nba);
            uint32_t ca = discretize(f[2], 2.0f, 20.0f, nd), cb = discretize(f[3], 2.0f, 20.0f, nd);
            uint32_t an = discretize(f[4], 0.0f, 180.0f, na);
            return res1 << 20 | res2 << 15 | ca << 10 | cb << 5 | an;
        }
        case 1: { // pdb_motif_sincos.rs:17-53 (defaults 8 / 3, clamp 16 / 16)
            float nd = nbd > 16 ? 16.0f : (nbd == 0 ? 8.0f : (float)nbd);
            float na = nba > 16 ? 16.0f : (nba == 0 ? 3.0f : (float)nba);
            uint32_t ca = discretize(f[2], 2.0f, 20.0f, nd), cb = discretize(f[3], 2.0f, 20.0f, nd);
            uint32_t sn = discretize(fdo_sinf(f[4]), -1.0f, 1.0f, na), cs = discretize(fdo_cosf(f[4]), -1.0f, 1.0f, na);
            return res1 << 21 | res2 << 16 | ca << 12 | cb << 8 | sn << 4 | cs;
        }
        case 2: { // trrosetta.rs:51-91 (clamp 8 / 4, zeros are NOT replaced here)
            float nd = (float)nbd > 8.0f ? 8.0f : (float)nbd;
            float na = (float)nba > 4.0f ? 4.0f : (float)nba;
            uint32_t pair = res1 * 20 + res2; // map_aa_u32_pair_to_u32 (convert.rs:203-207)
            uint32_t h = pair << 23 | discretize(f[2], 2.0f, 20.0f, nd) << 20;
            int shift = 18;
            for (int k = 3; k <= 7; k++) { // omega, theta1, theta2, phi1, phi2: sin then cos, two bits each
                h |= discretize(fdo_sinf(f[k]), -1.0f, 1.0f, na) << shift;
                h |= discretize(fdo_cosf(f[k]), -1.0f, 1.0f, na) << (shift - 2);
                shift -= 4;
            }
            return h;
        }
        case 4: { // ppf.rs:15-51 (defaults 8 / 3, clamp 16 / 8)
            float nd = nbd > 16 ? 16.0f : (nbd == 0 ? 8.0f : (float)nbd);
            float na = nba > 8 ? 8.0f : (nba == 0 ? 3.0f : (float)nba);
            uint32_t hd = discretize(f[2], 2.0f, 20.0f, nd);
            uint32_t s[3], c[3];
            for (int k = 0; k < 3; k++) {
                s[k] = discretize(fdo_sinf(f[3 + k]), -1.0f, 1.0f, na);
                c[k] = discretize(fdo_cosf(f[3 + k]), -1.0f, 1.0f, na);
            }
            return (res1 << 27) | (res2 << 22) | (hd << 18) | (s[0] << 15) | (c[0] << 12) | (s[1] << 9) | (c[1] << 6) |
                   (s[2] << 3) | c[2];
        }
        case 5: { // tertiary_interaction.rs:21-83 (defaults 8 / 3, clamp 16 / 8)
            float nd = nbd > 16 ? 16.0f : (nbd == 0 ? 8.0f : (float)nbd);
            float na = nba > 8 ? 8.0f : (nba == 0 ? 3.0f : (float)nba);
            uint32_t c[7];
            for (int k = 0; k < 7; k++) c[k] = discretize(fdo_cosf(f[k]), -1.0f, 1.0f, na);
            uint32_t ca = discretize(f[7], 2.0f, 20.0f, nd);
            uint32_t sd = f[8] < -4.0f ? 0u : (f[8] > 4.0f ? 8u : sat_u32(f[8]) + 4u);
            return (c[0] << 26) | (c[1] << 23) | (c[2] << 20) | (c[3] << 17) | (c[4] << 14) | (c[5] << 11) | (c[6] << 8) |
                   (ca << 4) | sd;
        }
        case 6: { // hybrid.rs:21-96 (defaults 16 / 4)
            float nd = nbd > 16 ? 16.0f : (nbd == 0 ? 16.0f : (float)nbd);
            float na = nba > 4 ? 4.0f : (nba == 0 ? 4.0f : (float)nba);
            uint32_t h = res1 << 30 | res2 << 28 | discretize(f[2], 2.0f, 20.0f, nd) << 24 | discretize(f[3], 2.0f, 20.0f, nd) << 20;
            int shift = 18;
            for (int k = 4; k <= 8; k++) { // ca-cb angle, phi1, phi2, bb_phi1, bb_phi2
                h |= discretize(fdo_sinf(f[k]), -1.0f, 1.0f, na) << shift;
                h |= discretize(fdo_cosf(f[k]), -1.0f, 1.0f, na) << (shift - 2);
                shift -= 4;
            }
            return h;
        }
        case 7: { // folddisco_angle.rs:24-71 (defaults 8 / 32, radians binned directly)
            float nd = nbd > 8 ? 8.0f : (nbd == 0 ? 8.0f : (float)nbd);
            float na = nba > 32 ? 32.0f : (nba == 0 ? 32.0f : (float)nba);
            uint32_t pair = res1 * 20 + res2;
            uint32_t ca = discretize(f[2], 2.0f, 20.0f, nd), cb = discretize(f[3], 2.0f, 20.0f, nd);
            uint32_t an = discretize(f[4], 0.0f, PI_F32, std::min(na, 32.0f));
            uint32_t p1 = discretize(f[5], -PI_F32, PI_F32, na), p2 = discretize(f[6], -PI_F32, PI_F32, na);
            return pair << 21 | ca << 18 | cb << 15 | an << 10 | p1 << 5 | p2;
        }
        case 8: { // folddisco_dist.rs:22-68 (defaults 32 / 16; the CA-CB angle gets at most 8 bins)
            float nd = nbd > 32 ? 32.0f : (nbd == 0 ? 32.0f : (float)nbd);
            float na = nba > 16 ? 16.0f : (nba == 0 ? 16.0f : (float)nba);
            uint32_t pair = res1 * 20 + res2;
            uint32_t ca = discretize(f[2], 2.0f, 20.0f, nd), cb = discretize(f[3], 2.0f, 20.0f, nd);
            uint32_t an = discretize(f[4], 0.0f, PI_F32, std::min(na, 8.0f));
            uint32_t p1 = discretize(f[5], -PI_F32, PI_F32, na), p2 = discretize(f[6], -PI_F32, PI_F32, na);
            return pair << 21 | ca << 16 | cb << 11 | an << 8 | p1 << 4 | p2;
        }
        default: return perfect_hash_pdbtr(f, nbd, nba);
    }
}
// HashType::default_dist_bin / default_angle_bin (geometry/core.rs:118-147)
void default_bins(uint32_t *nbd, uint32_t *nba) {
    switch (g_hash_type) {
        case 0: *nbd = 18, *nba = 9; break;
        case 3: *nbd = 16, *nba = 4; break;
        case 6: *nbd = 16, *nba = 4; break;
        case 7: *nbd = 8, *nba = 32; break;
        case 8: *nbd = 32, *nba = 16; break;
        default: *nbd = 8, *nba = 3; break;
    }
}
// `if nbin_dist == 0 || nbin_angle == 0 { perfect_hash_default } else { perfect_hash }` -- the form every single-bin
// call site uses (feature.rs:215-221, query.rs:72-76 and :283-287, retrieve.rs:133-137)
uint32_t perfect_hash(const float *f, uint32_t nbd, uint32_t nba) {
    if (nbd == 0 || nba == 0) default_bins(&nbd, &nba);
    return perfect_hash_raw(f, nbd, nba);
}

// pdb_tr.rs:95-136 (default bins) -> [res1,res2,ca,cb,angle_deg,phi1_deg,phi2_deg]
void reverse_hash_default(uint32_t h, float *o) {
    o[0] = (float)((h >> 25) & 31u);
    o[1] = (float)((h >> 20) & 31u);
    o[2] = continuize((h >> 16) & 15u, 2.0f, 20.0f, 16.0f);
    o[3] = continuize((h >> 12) & 15u, 2.0f, 20.0f, 16.0f);
    float s0 = continuize((h >> 10) & 3u, -1.0f, 1.0f, 4.0f), c0 = continuize((h >> 8) & 3u, -1.0f, 1.0f, 4.0f);
    float s1 = continuize((h >> 6) & 3u, -1.0f, 1.0f, 4.0f), c1 = continuize((h >> 4) & 3u, -1.0f, 1.0f, 4.0f);
    float s2 = continuize((h >> 2) & 3u, -1.0f, 1.0f, 4.0f), c2 = continuize(h & 3u, -1.0f, 1.0f, 4.0f);
    const float deg = 180.0f / 3.14159274101257324f; // f32::to_degrees multiplies by 180/PI (f32)
    o[4] = fdo_atan2f(s0, c0) * deg;
    o[5] = fdo_atan2f(s1, c1) * deg;
    o[6] = fdo_atan2f(s2, c2) * deg;
}
// (res1, res2) of HashValue::reverse_hash_default for the selected encoding
void hash_amino_acids(uint32_t h, uint32_t *a1, uint32_t *a2) {
    switch (g_hash_type) {
        case 0: *a1 = (h >> 20) & 31u, *a2 = (h >> 15) & 31u; break;               // pdb_motif.rs:57-58
        case 1: *a1 = (h >> 21) & 31u, *a2 = (h >> 16) & 31u; break;               // pdb_motif_sincos.rs:60-61
        case 2: *a1 = ((h >> 23) & 511u) / 20, *a2 = ((h >> 23) & 511u) % 20; break; // trrosetta.rs:94-95
        case 4: *a1 = (h >> 27) & 31u, *a2 = (h >> 22) & 31u; break;               // ppf.rs:59-60
        case 7:
        case 8: *a1 = ((h >> 21) & 511u) / 20, *a2 = ((h >> 21) & 511u) % 20; break; // folddisco_*.rs reverse_hash
        default: *a1 = (h >> 25) & 31u, *a2 = (h >> 20) & 31u; break;
    }
}
// HashValue::is_symmetric: pdb_tr.rs:158-162 and its counterparts (always over reverse_hash_default)
bool hash_is_symmetric(uint32_t h) {
    uint32_t a1, a2;
    hash_amino_acids(h, &a1, &a2);
    auto angle = [](uint32_t sbin, uint32_t cbin, float nb) {
        return to_degrees(fdo_atan2f(continuize(sbin, -1.0f, 1.0f, nb), continuize(cbin, -1.0f, 1.0f, nb)));
    };
    switch (g_hash_type) {
        case 5: return false; // tertiary_interaction.rs:145-150
        case 6: // hybrid.rs:185-189: groups equal, phi1 == phi2
            return ((h >> 30) & 3u) == ((h >> 28) & 3u) &&
                   angle((h >> 14) & 3u, (h >> 12) & 3u, 4.0f) == angle((h >> 10) & 3u, (h >> 8) & 3u, 4.0f);
        case 0:
        case 1: return a1 == a2; // pdb_motif.rs:98-102, pdb_motif_sincos.rs:105-109
        case 2: // trrosetta.rs:158-162: theta1 == theta2 and phi1 == phi2 (3 default sin / cos bins)
            return a1 == a2 && angle((h >> 14) & 3u, (h >> 12) & 3u, 3.0f) == angle((h >> 10) & 3u, (h >> 8) & 3u, 3.0f) &&
                   angle((h >> 6) & 3u, (h >> 4) & 3u, 3.0f) == angle((h >> 2) & 3u, h & 3u, 3.0f);
        case 4: // ppf.rs:124-127
            return a1 == a2 && angle((h >> 15) & 7u, (h >> 12) & 7u, 3.0f) == angle((h >> 9) & 7u, (h >> 6) & 7u, 3.0f);
        case 7: // folddisco_angle.rs:133-137
            return a1 == a2 && to_degrees(continuize((h >> 5) & 31u, -PI_F32, PI_F32, 32.0f)) ==
                                   to_degrees(continuize(h & 31u, -PI_F32, PI_F32, 32.0f));
        case 8: // folddisco_dist.rs:126-130
            return a1 == a2 && to_degrees(continuize((h >> 4) & 15u, -PI_F32, PI_F32, 16.0f)) ==
                                   to_degrees(continuize(h & 15u, -PI_F32, PI_F32, 16.0f));
        default: {
            float v[7];
            reverse_hash_default(h, v);
            return v[0] == v[1] && v[5] == v[6];
        }
    }
}

// feature.rs:198-231 + combination.rs:24-44 (row-major ordered pairs, i != j)
void hash_compact(const fdo_compact &c, uint32_t nbd, uint32_t nba, float cutoff, std::vector<uint32_t> &out) {
    const size_t n = c.nres();
    float f[9];
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) {
            if (i == j) continue;
            if (!pair_feature(c, i, j, cutoff, f)) continue;
            if (!g_multi_bins.empty()) { // feature.rs:210-214: no default substitution on this branch
                for (auto &b : g_multi_bins) out.push_back(perfect_hash_raw(f, b.first, b.second));
            } else {
                out.push_back(perfect_hash(f, nbd, nba));
            }
        }
}

// ---------------------------------------------------------------------------------------------
// Inverted index  (src/index/indextable.rs)
// ---------------------------------------------------------------------------------------------

// indextable.rs:397-418
inline size_t varint_len(uint64_t v) {
    size_t n = 1;
    while (v >= 128) {
        v >>= 7;
        n++;
    }
    return n;
}
inline void varint_put(uint64_t v, uint8_t *&p) {
    while (v >= 128) {
        *p++ = (uint8_t)((v & 0x7F) | 0x80);
        v >>= 7;
    }
    *p++ = (uint8_t)v;
}

} // namespace

struct fdo_index {
    std::vector<uint32_t> hashes;  // ascending
    std::vector<uint64_t> offsets; // count + 1
    std::vector<uint8_t> values;
    // indextable.rs:44-81
    bool raw(uint32_t h, const uint8_t **b, const uint8_t **e) const {
        auto it = std::lower_bound(hashes.begin(), hashes.end(), h);
        if (it == hashes.end() || *it != h) return false;
        size_t k = it - hashes.begin();
        *b = values.data() + offsets[k];
        *e = values.data() + offsets[k + 1];
        return *e > *b;
    }
    // indextable.rs:439-463
    void entries(uint32_t h, std::vector<uint64_t> &out) const {
        out.clear();
        const uint8_t *b, *e;
        if (!raw(h, &b, &e)) return;
        uint64_t prev = 0, cur = 0;
        int shift = 0;
        bool first = true;
        for (const uint8_t *p = b; p < e; p++) {
            cur |= (uint64_t)(*p & 0x7F) << shift;
            if (*p & 0x80) {
                shift += 7;
                continue;
            }
            if (first) {
                prev = cur;
                first = false;
            } else {
                prev += cur;
            }
            out.push_back(prev);
            cur = 0;
            shift = 0;
        }
    }
};

namespace {

// The reference counts varint bytes per hash (indextable.rs:88-105), prefix-sums the dense 2^30 table
// (:204-237), fills (:171-202), prunes to sparse (:267-295).  The result only depends on the multiset
// of (hash, id) pairs with ids ascending per hash; build that directly.
// sort (hash, id) pairs; large inputs are bucketed by the top hash bits and the buckets sorted in parallel
void sort_pairs(std::vector<std::pair<uint32_t, uint64_t>> &pairs, int threads) {
    if (pairs.size() < (1u << 20) || threads <= 1) {
        std::sort(pairs.begin(), pairs.end());
        return;
    }
    const int B = 4096; // hash >> 18 (hashes are 30-bit; overflowed bins land in the last buckets)
    auto bucket = [](uint32_t h) { return (int)std::min<uint32_t>(h >> 18, 4095u); };
    std::vector<size_t> cnt(B + 1, 0);
    for (auto &p : pairs) cnt[bucket(p.first) + 1]++;
    for (int b = 0; b < B; b++) cnt[b + 1] += cnt[b];
    std::vector<std::pair<uint32_t, uint64_t>> out(pairs.size());
    std::vector<size_t> cur(cnt.begin(), cnt.end() - 1);
    for (auto &p : pairs) out[cur[bucket(p.first)]++] = p;
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&] {
            for (int b; (b = next++) < B;) std::sort(out.begin() + cnt[b], out.begin() + cnt[b + 1]);
        });
    for (auto &t : th) t.join();
    pairs.swap(out);
}

fdo_index *index_from_pairs(std::vector<std::pair<uint32_t, uint64_t>> &pairs, int threads = 1) {
    sort_pairs(pairs, threads);
    fdo_index *ix = new fdo_index();
    size_t total = 0;
    {
        size_t i = 0;
        while (i < pairs.size()) {
            size_t j = i;
            uint64_t prev = 0;
            bool first = true;
            while (j < pairs.size() && pairs[j].first == pairs[i].first) {
                uint64_t d = first ? pairs[j].second : pairs[j].second - prev;
                total += varint_len(d);
                prev = pairs[j].second;
                first = false;
                j++;
            }
            ix->hashes.push_back(pairs[i].first);
            i = j;
        }
    }
    ix->values.resize(total);
    ix->offsets.reserve(ix->hashes.size() + 1);
    uint8_t *p = ix->values.data();
    size_t i = 0;
    ix->offsets.push_back(0);
    while (i < pairs.size()) {
        size_t j = i;
        uint64_t prev = 0;
        bool first = true;
        while (j < pairs.size() && pairs[j].first == pairs[i].first) {
            uint64_t d = first ? pairs[j].second : pairs[j].second - prev;
            varint_put(d, p);
            prev = pairs[j].second;
            first = false;
            j++;
        }
        ix->offsets.push_back((uint64_t)(p - ix->values.data()));
        i = j;
    }
    return ix;
}

// Rust `{}` for f32: shortest round-trip digits, never scientific notation.
std::string rust_f32(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[512];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

} // namespace

// ---------------------------------------------------------------------------------------------
// Query map (src/controller/query.rs)
// ---------------------------------------------------------------------------------------------

struct QEntry {
    uint32_t hash;
    size_t qi, qj;
    bool primary;
    float idf;
};
struct fdo_qmap {
    std::vector<QEntry> entries;                 // insertion order
    std::unordered_map<uint32_t, size_t> lookup; // hash -> position in entries
    std::vector<size_t> indices;                 // query.rs:236-246
    // residue_count of query_pdb.rs:355-359: the parsed query residues (all residues of the structure for an empty
    // query), whether or not make_query_map resolves them; denominator of the node-ratio filters
    size_t residue_count = 0;
    // observed_distance_map: (aa_i, aa_j) -> [(ca_dist, index_i)]  query.rs:271-280
    std::map<std::pair<uint8_t, uint8_t>, std::vector<std::pair<float, size_t>>> aa_dist;
    const QEntry *find(uint32_t h) const {
        auto it = lookup.find(h);
        return it == lookup.end() ? nullptr : &entries[it->second];
    }
};

namespace {

// query.rs:53-84
void insert_binned_hash(fdo_qmap &m, const float *f, size_t qi, size_t qj, uint32_t nbd, uint32_t nba,
                        bool primary, float idf) {
    auto put = [&](uint32_t h) {
        if (m.lookup.count(h)) return;
        m.lookup[h] = m.entries.size();
        m.entries.push_back({h, qi, qj, primary, idf});
    };
    if (!g_multi_bins.empty()) {
        for (auto &b : g_multi_bins) put(perfect_hash(f, b.first, b.second));
    } else {
        put(perfect_hash(f, nbd, nba));
    }
}

// query.rs:17-32
float idf_for_hash(uint32_t h, const fdo_index *ix, float total) {
    if (ix) {
        std::vector<uint64_t> e;
        ix->entries(h, e);
        if (!e.empty()) return log2f(total / (float)e.size());
    }
    return 0.0f;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// count_query (src/controller/count_query.rs)
// ---------------------------------------------------------------------------------------------

struct Hit {
    uint64_t nid;
    uint32_t match_count, node_count, edge_count;
    float idf;
};
struct fdo_hits {
    std::vector<Hit> v;
};

namespace {

struct CompactEntry { // count_query.rs:60-67
    uint16_t node_count = 0;
    uint32_t edge_count = 0, match_count = 0;
    float idf_sum = 0.0f;
    bool initialized = false;
};

// count_query.rs:222-253; `queries` is the key order of the map (insertion order here; the reference
// iterates an FxHashMap, so its order is unpinned).
std::vector<uint32_t> sample_query(const std::vector<uint32_t> &queries, const fdo_index &ix, float ratio,
                                   int64_t count) {
    bool has_r = ratio >= 0.0f, has_c = count >= 0;
    if (has_r == has_c) return queries; // (None,None) or (Some,Some)
    std::vector<std::pair<uint32_t, size_t>> s;
    std::vector<uint64_t> e;
    for (uint32_t q : queries) {
        ix.entries(q, e);
        s.push_back({q, e.size()});
    }
    std::stable_sort(s.begin(), s.end(), [](auto &a, auto &b) { return a.second < b.second; });
    size_t keep = has_r ? (size_t)std::ceil(ratio * (float)s.size()) : (size_t)count;
    if (keep < s.size()) s.resize(keep);
    std::vector<uint32_t> out;
    for (auto &p : s) out.push_back(p.first);
    return out;
}

void count_query(const fdo_qmap &m, const fdo_index &ix, uint64_t num_ids, const uint64_t *nres,
                 const fdo_count_params &p, int threads, std::vector<Hit> &out, uint64_t *posting_bytes) {
    std::vector<uint32_t> queries;
    for (auto &e : m.entries) queries.push_back(e.hash);
    std::vector<uint32_t> iter = sample_query(queries, ix, p.sampling_ratio, p.sampling_count);
    const float lp = p.length_penalty;
    // build_node_groups (count_query.rs:256-273): group by source node, stable sort by edge
    std::map<size_t, std::vector<std::pair<std::pair<size_t, size_t>, uint32_t>>> groups;
    for (uint32_t q : iter) {
        const QEntry *e = m.find(q);
        if (e) groups[e->qi].push_back({{e->qi, e->qj}, q});
    }
    std::vector<const std::vector<std::pair<std::pair<size_t, size_t>, uint32_t>> *> glist;
    for (auto &g : groups) {
        std::stable_sort(g.second.begin(), g.second.end(), [](auto &a, auto &b) { return a.first < b.first; });
        glist.push_back(&g.second);
    }
    std::vector<std::vector<CompactEntry>> thread_results(glist.size());
    std::atomic<uint64_t> bytes{0};
    auto work = [&](size_t g) {
        auto &chunk = *glist[g];
        std::vector<CompactEntry> local(num_ids);
        std::vector<uint8_t> edge_occ(num_ids, 0);
        bool have_prev = false;
        std::pair<size_t, size_t> prev_edge{};
        std::vector<uint64_t> vals;
        for (auto &it : chunk) {
            if (!have_prev || prev_edge != it.first) {
                if (have_prev)
                    for (uint64_t nid = 0; nid < num_ids; nid++)
                        if (edge_occ[nid]) local[nid].edge_count++;
                std::fill(edge_occ.begin(), edge_occ.end(), 0);
                prev_edge = it.first;
                have_prev = true;
            }
            ix.entries(it.second, vals);
            {
                const uint8_t *b, *e;
                if (ix.raw(it.second, &b, &e)) bytes += (uint64_t)(e - b);
            }
            size_t hash_count = vals.size();
            if (p.freq_filter >= 0.0f)
                if ((float)hash_count / (float)num_ids > p.freq_filter) continue;
            float idf = log2f((float)num_ids / (float)hash_count);
            for (uint64_t v : vals) {
                if (v >= num_ids) continue;
                CompactEntry &en = local[v]; // lookup[value].1 == value for every index we write
                en.initialized = true;
                en.match_count++;
                en.idf_sum += idf;
                edge_occ[v] = 1;
            }
        }
        for (uint64_t nid = 0; nid < num_ids; nid++) {
            if (local[nid].initialized) local[nid].node_count = 1;
            if (edge_occ[nid]) local[nid].edge_count++;
        }
        thread_results[g] = std::move(local);
    };
    if (threads <= 1 || glist.size() <= 1) {
        for (size_t g = 0; g < glist.size(); g++) work(g);
    } else {
        std::atomic<size_t> next{0};
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++)
            th.emplace_back([&] {
                for (size_t g; (g = next++) < glist.size();) work(g);
            });
        for (auto &t : th) t.join();
    }
    if (posting_bytes) *posting_bytes = bytes;
    // merge (count_query.rs:172-217)
    for (uint64_t nid = 0; nid < num_ids; nid++) {
        CompactEntry me;
        bool found = false;
        for (auto &arr : thread_results) {
            const CompactEntry &e = arr[nid];
            if (!e.initialized) continue;
            if (!found) {
                me = e;
                found = true;
            } else {
                me.match_count += e.match_count;
                me.idf_sum += e.idf_sum;
                me.node_count += e.node_count;
                me.edge_count += e.edge_count;
            }
        }
        if (found && me.match_count > 0) {
            me.idf_sum *= powf((float)nres[nid], -lp);
            out.push_back({nid, me.match_count, me.node_count, me.edge_count, me.idf_sum});
        }
    }
}

// filter.rs:76-100
bool filter_before_matching(const Hit &h, uint64_t nres, float plddt, const fdo_count_params &p) {
    bool pass = true;
    if (p.total_match_count > 0) pass = pass && h.match_count >= p.total_match_count;
    if (p.covered_node_count > 0) pass = pass && h.node_count >= p.covered_node_count;
    if (p.covered_node_ratio > 0.0f)
        pass = pass && (float)h.node_count / (float)p.expected_node_count >= p.covered_node_ratio;
    if (p.idf_score_cutoff > 0.0f) pass = pass && h.idf >= p.idf_score_cutoff;
    if (p.num_res_cutoff > 0) pass = pass && nres <= p.num_res_cutoff;
    if (p.plddt_cutoff > 0.0f) pass = pass && plddt >= p.plddt_cutoff;
    return pass;
}

void filter_sort_top(std::vector<Hit> &hits, const uint64_t *nres, const float *plddt,
                     const fdo_count_params &p) {
    std::vector<Hit> kept;
    for (auto &h : hits)
        if (filter_before_matching(h, nres[h.nid], plddt ? plddt[h.nid] : 0.0f, p)) kept.push_back(h);
    // query_pdb.rs:404 stable par_sort_by idf desc; partial_cmp().unwrap() panics on NaN in the reference
    std::stable_sort(kept.begin(), kept.end(), [](const Hit &a, const Hit &b) { return a.idf > b.idf; });
    if (p.top_n != UINT64_MAX && kept.size() > p.top_n) kept.resize(p.top_n);
    hits.swap(kept);
}

} // namespace

// ---------------------------------------------------------------------------------------------
// Kabsch (src/structure/kabsch.rs:157-554), mode 2
// ---------------------------------------------------------------------------------------------
namespace {

bool kabsch(const std::vector<std::array<float, 3>> &xf, const std::vector<std::array<float, 3>> &yf,
            float U[9], float T[3], float *rmsd_out) {
    const double EPSILON = 1.0e-8, TOLERANCE = 0.01, SQRT3 = 1.7320508075688772;
    static const int IP[9] = {0, 1, 3, 1, 2, 4, 3, 4, 5};
    static const int IP2312[4] = {1, 2, 0, 1};
    const size_t n = xf.size();
    auto ident = [&] {
        for (int i = 0; i < 9; i++) U[i] = (i % 4 == 0) ? 1.0f : 0.0f;
        T[0] = T[1] = T[2] = 0.0f;
    };
    if (n == 0 || yf.size() != n) {
        ident();
        *rmsd_out = 3.40282347e+38f;
        return true;
    }
    double rms = 0.0, e0 = 0.0;
    double s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0}, sx[3] = {0, 0, 0}, sy[3] = {0, 0, 0}, sz[3] = {0, 0, 0};
    double xc[3], yc[3], t[3] = {0, 0, 0}, e[3];
    double r[3][3], a[3][3] = {{0}}, b[3][3] = {{0}}, u[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double rr[6], ss[6];
    for (size_t i = 0; i < n; i++) {
        double c1[3] = {xf[i][0], xf[i][1], xf[i][2]}, c2[3] = {yf[i][0], yf[i][1], yf[i][2]};
        for (int j = 0; j < 3; j++) {
            s1[j] += c1[j];
            s2[j] += c2[j];
        }
        sx[0] += c1[0] * c2[0];
        sx[1] += c1[0] * c2[1];
        sx[2] += c1[0] * c2[2];
        sy[0] += c1[1] * c2[0];
        sy[1] += c1[1] * c2[1];
        sy[2] += c1[1] * c2[2];
        sz[0] += c1[2] * c2[0];
        sz[1] += c1[2] * c2[1];
        sz[2] += c1[2] * c2[2];
    }
    for (int j = 0; j < 3; j++) {
        xc[j] = s1[j] / (double)n;
        yc[j] = s2[j] / (double)n;
    }
    for (size_t i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) {
            double dx = (double)xf[i][k] - xc[k], dy = (double)yf[i][k] - yc[k];
            e0 += dx * dx + dy * dy;
        }
    }
    for (int j = 0; j < 3; j++) {
        r[j][0] = sx[j] - s1[0] * s2[j] / (double)n;
        r[j][1] = sy[j] - s1[1] * s2[j] / (double)n;
        r[j][2] = sz[j] - s1[2] * s2[j] / (double)n;
    }
    double det_r = r[0][0] * (r[1][1] * r[2][2] - r[1][2] * r[2][1]) -
                   r[0][1] * (r[1][0] * r[2][2] - r[1][2] * r[2][0]) +
                   r[0][2] * (r[1][0] * r[2][1] - r[1][1] * r[2][0]);
    double sigma = det_r;
    int m = 0;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i <= j; i++) rr[m++] = r[0][i] * r[0][j] + r[1][i] * r[1][j] + r[2][i] * r[2][j];
    double spur = (rr[0] + rr[2] + rr[5]) / 3.0;
    double cof = (((rr[2] * rr[5] - rr[4] * rr[4]) + rr[0] * rr[5] - rr[3] * rr[3]) + rr[0] * rr[2] -
                  rr[1] * rr[1]) /
                 3.0;
    double det = det_r * det_r;
    e[0] = e[1] = e[2] = spur;
    if (spur > 0.0) {
        double d = spur * spur;
        double h = d - cof;
        double g = (spur * cof - det) / 2.0 - spur * h;
        if (h > 0.0) {
            double sqrth = sqrt(h);
            double disc = h * h * h - g * g;
            if (disc < 0.0) disc = 0.0;
            double sqrt_disc = sqrt(disc);
            double d_ang;
            if (fabs(g) > 1e18) d_ang = g > 0.0 ? M_PI / 3.0 : 0.0;
            else d_ang = atan2(sqrt_disc, -g) / 3.0;
            double cth = sqrth * cos(d_ang);
            double sth = sqrth * SQRT3 * sin(d_ang);
            e[0] = spur + 2.0 * cth;
            e[1] = spur - cth + sth;
            e[2] = spur - cth - sth;
            bool a_failed = false, b_failed = false;
            for (int l : {0, 2}) {
                double dl = e[l];
                ss[0] = (dl - rr[2]) * (dl - rr[5]) - rr[4] * rr[4];
                ss[1] = (dl - rr[5]) * rr[1] + rr[3] * rr[4];
                ss[2] = (dl - rr[0]) * (dl - rr[5]) - rr[3] * rr[3];
                ss[3] = (dl - rr[2]) * rr[3] + rr[1] * rr[4];
                ss[4] = (dl - rr[0]) * rr[4] + rr[1] * rr[3];
                ss[5] = (dl - rr[0]) * (dl - rr[2]) - rr[1] * rr[1];
                for (double &s : ss)
                    if (fabs(s) <= EPSILON) s = 0.0;
                double A = fabs(ss[0]), B = fabs(ss[2]), C = fabs(ss[5]);
                int j = (A >= B && A >= C) ? 0 : (B >= C ? 1 : 2);
                double dn = 0.0;
                for (int i = 0; i < 3; i++) {
                    int k = IP[3 * j + i];
                    a[i][l] = ss[k];
                    dn += ss[k] * ss[k];
                }
                dn = dn > EPSILON ? 1.0 / sqrt(dn) : 0.0;
                for (int i = 0; i < 3; i++) a[i][l] *= dn;
            }
            double dot = a[0][0] * a[0][2] + a[1][0] * a[1][2] + a[2][0] * a[2][2];
            int m1, mm;
            if (e[0] - e[1] > e[1] - e[2]) {
                m1 = 2;
                mm = 0;
            } else {
                m1 = 0;
                mm = 2;
            }
            double p = 0.0;
            for (int i = 0; i < 3; i++) {
                a[i][m1] = a[i][m1] - dot * a[i][mm];
                p += a[i][m1] * a[i][m1];
            }
            if (p <= TOLERANCE) {
                int j = 0;
                p = 1.0;
                for (int i = 0; i < 3; i++)
                    if (p < fabs(a[i][mm])) {
                        p = fabs(a[i][mm]);
                        j = i;
                    }
                int k = IP2312[j], l = IP2312[j + 1];
                p = sqrt(a[k][mm] * a[k][mm] + a[l][mm] * a[l][mm]);
                if (p > TOLERANCE) {
                    a[j][m1] = 0.0;
                    a[k][m1] = -a[l][mm] / p;
                    a[l][m1] = a[k][mm] / p;
                } else {
                    a_failed = true;
                }
            } else {
                p = 1.0 / sqrt(p);
                for (int i = 0; i < 3; i++) a[i][m1] *= p;
            }
            if (!a_failed) {
                a[0][1] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
                a[1][1] = a[2][2] * a[0][0] - a[2][0] * a[0][2];
                a[2][1] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
                for (int l = 0; l < 2; l++) {
                    double db = 0.0;
                    for (int i = 0; i < 3; i++) {
                        b[i][l] = r[i][0] * a[0][l] + r[i][1] * a[1][l] + r[i][2] * a[2][l];
                        db += b[i][l] * b[i][l];
                    }
                    db = db > EPSILON ? 1.0 / sqrt(db) : 0.0;
                    for (int i = 0; i < 3; i++) b[i][l] *= db;
                }
                double dot_b = 0.0;
                for (int i = 0; i < 3; i++) dot_b += b[i][0] * b[i][1];
                double pb = 0.0;
                for (int i = 0; i < 3; i++) {
                    b[i][1] -= dot_b * b[i][0];
                    pb += b[i][1] * b[i][1];
                }
                if (pb <= TOLERANCE) {
                    pb = 1.0;
                    int j = 0;
                    for (int i = 0; i < 3; i++)
                        if (pb < fabs(b[i][0])) {
                            pb = fabs(b[i][0]);
                            j = i;
                        }
                    int k = IP2312[j], l = IP2312[j + 1];
                    pb = sqrt(b[k][0] * b[k][0] + b[l][0] * b[l][0]);
                    if (pb > TOLERANCE) {
                        b[j][1] = 0.0;
                        b[k][1] = -b[l][0] / pb;
                        b[l][1] = b[k][0] / pb;
                    } else {
                        b_failed = true;
                    }
                } else {
                    pb = 1.0 / sqrt(pb);
                    for (int i = 0; i < 3; i++) b[i][1] *= pb;
                }
                if (!b_failed) {
                    b[0][2] = b[1][0] * b[2][1] - b[1][1] * b[2][0];
                    b[1][2] = b[2][0] * b[0][1] - b[2][1] * b[0][0];
                    b[2][2] = b[0][0] * b[1][1] - b[0][1] * b[1][0];
                    for (int i = 0; i < 3; i++)
                        for (int j = 0; j < 3; j++)
                            u[i][j] = b[i][0] * a[j][0] + b[i][1] * a[j][1] + b[i][2] * a[j][2];
                    for (int i = 0; i < 3; i++)
                        t[i] = yc[i] - (u[i][0] * xc[0] + u[i][1] * xc[1] + u[i][2] * xc[2]);
                }
            }
        }
    } else {
        for (int i = 0; i < 3; i++) t[i] = yc[i] - (u[i][0] * xc[0] + u[i][1] * xc[1] + u[i][2] * xc[2]);
    }
    (void)sigma;
    (void)e0;
    double sum_sq = 0.0;
    for (size_t i = 0; i < n; i++) {
        double x0 = xf[i][0], x1 = xf[i][1], x2 = xf[i][2];
        double tr[3] = {u[0][0] * x0 + u[0][1] * x1 + u[0][2] * x2 + t[0],
                        u[1][0] * x0 + u[1][1] * x1 + u[1][2] * x2 + t[1],
                        u[2][0] * x0 + u[2][1] * x1 + u[2][2] * x2 + t[2]};
        for (int j = 0; j < 3; j++) {
            double diff = tr[j] - (double)yf[i][j];
            sum_sq += diff * diff;
        }
    }
    rms = sqrt(sum_sq / (double)n);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) U[3 * i + j] = (float)u[i][j];
    for (int i = 0; i < 3; i++) T[i] = (float)t[i];
    float rf = (float)rms;
    if (std::isnan(rf)) rf = 3.40282347e+38f;
    *rmsd_out = rf;
    return true;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// Retrieval (src/controller/retrieve.rs, graph.rs)
// ---------------------------------------------------------------------------------------------

struct ResMatch {
    bool some = false;
    uint8_t chain = 0;
    uint64_t serial = 0;
    bool operator==(const ResMatch &o) const {
        return some == o.some && (!some || (chain == o.chain && serial == o.serial));
    }
};
struct MatchRow {
    std::vector<ResMatch> res;
    float rmsd;
    float U[9], T[3];
    float idf;
    float metrics[5] = {0, 0, 0, 0, 0}; // tm_score, gdt_ts, gdt_ha, chamfer_distance, hausdorff_distance
    std::vector<std::array<float, 3>> target_ca; // matching_coordinates (retrieve.rs:769-771)
};
struct fdo_matches {
    size_t n_query = 0;
    std::vector<MatchRow> from_hash, result;
    size_t max_node = 0;
    float min_rmsd = 0.0f;
    std::vector<std::array<int64_t, 2>> edges;
    std::vector<uint32_t> edge_hash;
};

namespace {

struct Edge {
    size_t i, j;
    uint32_t h;
};

// graph.rs:29-50: tarjan_scc U connected components of the undirected graph, size >= node_count,
// each sorted by node index, list sorted, deduped.  Node index = order of first appearance.
std::vector<std::vector<size_t>> components(size_t n_nodes, const std::vector<std::pair<size_t, size_t>> &edges,
                                            size_t min_size) {
    std::vector<std::vector<size_t>> adj(n_nodes), und(n_nodes);
    for (auto &e : edges) {
        adj[e.first].push_back(e.second);
        und[e.first].push_back(e.second);
        und[e.second].push_back(e.first);
    }
    std::vector<std::vector<size_t>> out;
    { // Tarjan, iterative
        std::vector<int64_t> index(n_nodes, -1), low(n_nodes, 0);
        std::vector<uint8_t> on(n_nodes, 0);
        std::vector<size_t> stack;
        int64_t counter = 0;
        struct Fr {
            size_t v, k;
        };
        for (size_t s = 0; s < n_nodes; s++) {
            if (index[s] >= 0) continue;
            std::vector<Fr> cs{{s, 0}};
            index[s] = low[s] = counter++;
            stack.push_back(s);
            on[s] = 1;
            while (!cs.empty()) {
                Fr &f = cs.back();
                if (f.k < adj[f.v].size()) {
                    size_t w = adj[f.v][f.k++];
                    if (index[w] < 0) {
                        index[w] = low[w] = counter++;
                        stack.push_back(w);
                        on[w] = 1;
                        cs.push_back({w, 0});
                    } else if (on[w]) {
                        low[f.v] = std::min(low[f.v], index[w]);
                    }
                } else {
                    size_t v = f.v;
                    cs.pop_back();
                    if (!cs.empty()) low[cs.back().v] = std::min(low[cs.back().v], low[v]);
                    if (low[v] == index[v]) {
                        std::vector<size_t> comp;
                        size_t w;
                        do {
                            w = stack.back();
                            stack.pop_back();
                            on[w] = 0;
                            comp.push_back(w);
                        } while (w != v);
                        out.push_back(comp);
                    }
                }
            }
        }
    }
    { // undirected components
        std::vector<uint8_t> seen(n_nodes, 0);
        for (size_t s = 0; s < n_nodes; s++) {
            if (seen[s]) continue;
            std::vector<size_t> comp, st{s};
            seen[s] = 1;
            while (!st.empty()) {
                size_t v = st.back();
                st.pop_back();
                comp.push_back(v);
                for (size_t w : und[v])
                    if (!seen[w]) {
                        seen[w] = 1;
                        st.push_back(w);
                    }
            }
            out.push_back(comp);
        }
    }
    std::vector<std::vector<size_t>> kept;
    for (auto &c : out)
        if (c.size() >= min_size) {
            std::sort(c.begin(), c.end());
            kept.push_back(c);
        }
    std::sort(kept.begin(), kept.end());
    kept.erase(std::unique(kept.begin(), kept.end()), kept.end());
    return kept;
}

// retrieve.rs:604-702.  sub_edges: edges of the component subgraph in insertion order, endpoints are
// target residue indices.
void map_query_and_retrieved(const std::vector<Edge> &sub_edges, const fdo_qmap &m, size_t node_count,
                             std::vector<size_t> &q_out, std::vector<size_t> &r_out) {
    size_t max_q = 0, max_r = 0;
    for (auto &e : m.entries) max_q = std::max(max_q, std::max(e.qi, e.qj));
    for (auto &e : sub_edges) max_r = std::max(max_r, std::max(e.i, e.j));
    size_t q_size = max_q + 1, r_size = max_r + 1;
    std::vector<uint8_t> counts(q_size * r_size, 0);
    std::vector<std::pair<uint8_t, size_t>> best(q_size, {0, 0});
    for (auto &e : sub_edges) {
        const QEntry *q = m.find(e.h);
        if (!q) continue;
        std::pair<size_t, size_t> pr[2];
        if (hash_is_symmetric(e.h)) {
            size_t q1, q2, r1, r2;
            if (q->qi < q->qj) {
                q1 = q->qi;
                q2 = q->qj;
            } else {
                q1 = q->qj;
                q2 = q->qi;
            }
            if (e.i < e.j) {
                r1 = e.i;
                r2 = e.j;
            } else {
                r1 = e.j;
                r2 = e.i;
            }
            pr[0] = {q1, r1};
            pr[1] = {q2, r2};
        } else {
            pr[0] = {q->qi, e.i};
            pr[1] = {q->qj, e.j};
        }
        for (auto &qr : pr) {
            uint8_t &c = counts[qr.first * r_size + qr.second];
            if (c != 255) c++;
            if (c > best[qr.first].first || (c == best[qr.first].first && qr.second < best[qr.first].second))
                best[qr.first] = {c, qr.second};
        }
    }
    std::vector<std::vector<std::pair<size_t, size_t>>> buckets(256);
    for (size_t q = 0; q < q_size; q++)
        if (best[q].first > 0) buckets[best[q].first].push_back({q, best[q].second});
    std::vector<uint8_t> q_used(q_size, 0), r_used(r_size, 0);
    for (size_t bi = 256; bi-- > 0;) {
        for (auto &qr : buckets[bi]) {
            if (!q_used[qr.first] && !r_used[qr.second]) {
                q_out.push_back(qr.first);
                r_out.push_back(qr.second);
                q_used[qr.first] = r_used[qr.second] = 1;
                if (q_out.size() == node_count) return;
            }
        }
    }
}

// src/structure/metrics.rs:44-273 over PrecomputedDistances::new(reference, transformed): the n x n matrix of f32
// distances (f64 inside), then the five metrics with the reference's use of a distance where a squared distance is meant
void similarity_metrics(const std::vector<std::array<float, 3>> &ref, const std::vector<std::array<float, 3>> &mov,
                        const float *U, const float *T, float *out) {
    const size_t n = ref.size();
    if (n == 0 || n != mov.size()) {
        out[0] = out[1] = out[2] = 0.0f;
        out[3] = out[4] = INFINITY;
        return;
    }
    // kabsch.rs:84-93, 141-151: transformed = rot * coord + tran in f32
    std::vector<std::array<float, 3>> tr(n);
    for (size_t i = 0; i < n; i++)
        for (int r = 0; r < 3; r++)
            tr[i][r] = (U[3 * r] * mov[i][0] + U[3 * r + 1] * mov[i][1] + U[3 * r + 2] * mov[i][2]) + T[r];
    std::vector<float> pd(n * n); // pairwise_dist[i * n + j] = dist(reference[j], coords[i])  (metrics.rs:74-78)
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) {
            double dx = (double)ref[j][0] - (double)tr[i][0], dy = (double)ref[j][1] - (double)tr[i][1],
                   dz = (double)ref[j][2] - (double)tr[i][2];
            pd[i * n + j] = (float)std::sqrt(dx * dx + dy * dy + dz * dz);
        }
    float d0 = n > 21 ? 1.24f * powf((float)n - 15.0f, 1.0f / 3.0f) - 1.8f : 0.5f; // metrics.rs:116-122
    double d0_sq = (double)(d0 * d0), sum = 0.0;
    for (size_t i = 0; i < n; i++) sum += 1.0 / (1.0 + (double)pd[i * n + i] / d0_sq); // metrics.rs:141-147
    out[0] = (float)(sum / (double)n);
    auto gdt = [&](const double *cut) { // metrics.rs:152-173
        double s2 = 0.0;
        for (int k = 0; k < 4; k++) {
            size_t cnt = 0;
            for (size_t i = 0; i < n; i++) cnt += (double)pd[i * n + i] <= cut[k] * cut[k] ? 1 : 0;
            s2 += (double)cnt / (double)n;
        }
        return (float)(s2 / 4.0);
    };
    const double ts[4] = {1.0, 2.0, 4.0, 8.0}, ha[4] = {0.5, 1.0, 2.0, 4.0};
    out[1] = gdt(ts);
    out[2] = gdt(ha);
    double cs = 0.0; // metrics.rs:205-221, 235-252
    float hd = 0.0f;
    for (size_t i = 0; i < n; i++) {
        float mn = pd[i * n];
        for (size_t j = 1; j < n; j++) mn = std::min(mn, pd[i * n + j]);
        cs += (double)mn;
        hd = i == 0 ? mn : std::max(hd, mn);
    }
    out[3] = (float)(cs / (double)n);
    out[4] = hd;
}

// ---------------------------------------------------------------------------------------------
// LMS-QCP partial superposition (src/structure/lms_qcp.rs), default parameters (:28-40)
// ---------------------------------------------------------------------------------------------
int g_partial_fit = 0; // fdo_set_partial_fit: retrieve() superposes like `--partial-fit`

struct LmsStats { // RunningStats :245-287
    size_t n = 0;
    double sum_x[3] = {0, 0, 0}, sum_y[3] = {0, 0, 0}, sxx = 0, syy = 0, syx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    void add(const std::array<float, 3> &m, const std::array<float, 3> &r) {
        double x[3] = {m[0], m[1], m[2]}, y[3] = {r[0], r[1], r[2]};
        n += 1;
        sum_x[0] += x[0]; sum_x[1] += x[1]; sum_x[2] += x[2];
        sum_y[0] += y[0]; sum_y[1] += y[1]; sum_y[2] += y[2];
        sxx += x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
        syy += y[0] * y[0] + y[1] * y[1] + y[2] * y[2];
        syx[0][0] += y[0] * x[0]; syx[0][1] += y[0] * x[1]; syx[0][2] += y[0] * x[2];
        syx[1][0] += y[1] * x[0]; syx[1][1] += y[1] * x[1]; syx[1][2] += y[1] * x[2];
        syx[2][0] += y[2] * x[0]; syx[2][1] += y[2] * x[1]; syx[2][2] += y[2] * x[2];
    }
};
struct RotTran {
    float r[3][3], t[3];
};
// qcp_from_a_e0 :330-447
void lms_qcp_rotation(const double a[3][3], double e0, double rot[3][3]) {
    double sxx = a[0][0], sxy = a[0][1], sxz = a[0][2], syx = a[1][0], syy = a[1][1], syz = a[1][2], szx = a[2][0], szy = a[2][1], szz = a[2][2];
    double sxx2 = sxx * sxx, syy2 = syy * syy, szz2 = szz * szz, sxy2 = sxy * sxy, syz2 = syz * syz, sxz2 = sxz * sxz;
    double syx2 = syx * syx, szy2 = szy * szy, szx2 = szx * szx;
    double syz_szy_m_syy_szz2 = 2.0 * (syz * szy - syy * szz);
    double sxx2_syy2_szz2_syz2_szy2 = syy2 + szz2 - sxx2 + syz2 + szy2;
    double c2 = -2.0 * (sxx2 + syy2 + szz2 + sxy2 + syx2 + sxz2 + szx2 + syz2 + szy2);
    double c1 = 8.0 * (sxx * syz * szy + syy * szx * sxz + szz * sxy * syx - sxx * syy * szz - syz * szx * sxy - szy * syx * sxz);
    double sxz_p_szx = sxz + szx, syz_p_szy = syz + szy, sxy_p_syx = sxy + syx;
    double syz_m_szy = syz - szy, sxz_m_szx = sxz - szx, sxy_m_syx = sxy - syx;
    double sxx_p_syy = sxx + syy, sxx_m_syy = sxx - syy;
    double sxy2_sxz2_syx2_szx2 = sxy2 + sxz2 - syx2 - szx2;
    double neg_sxz_p_szx = -sxz_p_szx, neg_sxz_m_szx = -sxz_m_szx, neg_sxy_m_syx = -sxy_m_syx;
    double sxx_p_syy_p_szz = sxx_p_syy + szz;
    double c0 = sxy2_sxz2_syx2_szx2 * sxy2_sxz2_syx2_szx2
        + (sxx2_syy2_szz2_syz2_szy2 + syz_szy_m_syy_szz2) * (sxx2_syy2_szz2_syz2_szy2 - syz_szy_m_syy_szz2)
        + (neg_sxz_p_szx * (syz_m_szy) + (sxy_m_syx) * (sxx_m_syy - szz)) * (neg_sxz_m_szx * (syz_p_szy) + (sxy_m_syx) * (sxx_m_syy + szz))
        + (neg_sxz_p_szx * (syz_p_szy) - (sxy_p_syx) * (sxx_p_syy - szz)) * (neg_sxz_m_szx * (syz_m_szy) - (sxy_p_syx) * sxx_p_syy_p_szz)
        + ((sxy_p_syx) * (syz_p_szy) + (sxz_p_szx) * (sxx_m_syy + szz)) * (neg_sxy_m_syx * (syz_m_szy) + (sxz_p_szx) * sxx_p_syy_p_szz)
        + ((sxy_p_syx) * (syz_m_szy) + (sxz_m_szx) * (sxx_m_syy - szz)) * (neg_sxy_m_syx * (syz_p_szy) + (sxz_m_szx) * (sxx_p_syy - szz));
    double lam = std::max(e0, 0.0);
    const double eps = 1e-15;
    for (int it = 0; it < 10; it++) {
        double x2 = lam * lam;
        double b = (x2 + c2) * lam;
        double aa = b + c1;
        double f = aa * lam + c0;
        double fp = 2.0 * x2 * lam + b + aa;
        double delta = f / (fp + eps);
        double nlam = std::fabs(lam - delta);
        if (std::fabs(nlam - lam) < eps * nlam) {
            lam = nlam;
            break;
        }
        lam = nlam;
    }
    double a11 = sxx_p_syy + szz - lam, a12 = syz_m_szy, a13 = neg_sxz_m_szx, a14 = sxy_m_syx;
    double a21 = a12, a22 = sxx_m_syy - szz - lam, a23 = sxy_p_syx, a24 = sxz_p_szx;
    double a31 = a13, a32 = a23, a33 = syy - sxx - szz - lam, a34 = syz_p_szy;
    double a41 = a14, a42 = a24, a43 = a34, a44 = szz - sxx_p_syy - lam;
    double a3344_4334 = a33 * a44 - a43 * a34, a3244_4234 = a32 * a44 - a42 * a34, a3243_4233 = a32 * a43 - a42 * a33;
    double a3143_4133 = a31 * a43 - a41 * a33, a3144_4134 = a31 * a44 - a41 * a34, a3142_4132 = a31 * a42 - a41 * a32;
    double q1 = a22 * a3344_4334 - a23 * a3244_4234 + a24 * a3243_4233;
    double q2 = -a21 * a3344_4334 + a23 * a3144_4134 - a24 * a3143_4133;
    double q3 = a21 * a3244_4234 - a22 * a3144_4134 + a24 * a3142_4132;
    double q4 = -a21 * a3243_4233 + a22 * a3143_4133 - a23 * a3142_4132;
    double qsqr = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
    if (qsqr < 1e-12) {
        q1 = a12 * a3344_4334 - a13 * a3244_4234 + a14 * a3243_4233;
        q2 = -a11 * a3344_4334 + a13 * a3144_4134 - a14 * a3143_4133;
        q3 = a11 * a3244_4234 - a12 * a3144_4134 + a14 * a3142_4132;
        q4 = -a11 * a3243_4233 + a12 * a3143_4133 - a13 * a3142_4132;
        qsqr = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
        if (qsqr < 1e-12) {
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) rot[r][c] = r == c ? 1.0 : 0.0;
            return;
        }
    }
    double inv = 1.0 / std::sqrt(qsqr);
    q1 *= inv; q2 *= inv; q3 *= inv; q4 *= inv;
    double a2 = q1 * q1, x2 = q2 * q2, y2 = q3 * q3, z2 = q4 * q4;
    double xy = q2 * q3, az = q1 * q4, zx = q4 * q2, ay = q1 * q3, yz = q3 * q4, ax = q1 * q2;
    rot[0][0] = a2 + x2 - y2 - z2; rot[0][1] = 2.0 * (xy + az);     rot[0][2] = 2.0 * (zx - ay);
    rot[1][0] = 2.0 * (xy - az);     rot[1][1] = a2 - x2 + y2 - z2; rot[1][2] = 2.0 * (yz + ax);
    rot[2][0] = 2.0 * (zx + ay);     rot[2][1] = 2.0 * (yz - ax);     rot[2][2] = a2 - x2 - y2 + z2;
}
// qcp_from_stats :294-326
RotTran lms_qcp_from_stats(const LmsStats &st) {
    double inv = 1.0 / (double)st.n;
    double mux[3] = {st.sum_x[0] * inv, st.sum_x[1] * inv, st.sum_x[2] * inv};
    double muy[3] = {st.sum_y[0] * inv, st.sum_y[1] * inv, st.sum_y[2] * inv};
    double a[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) a[r][c] = st.syx[r][c] - (double)st.n * (muy[r] * mux[c]);
    double mu2x = mux[0] * mux[0] + mux[1] * mux[1] + mux[2] * mux[2];
    double mu2y = muy[0] * muy[0] + muy[1] * muy[1] + muy[2] * muy[2];
    double e0 = 0.5 * std::max((st.syy - (double)st.n * mu2y) + (st.sxx - (double)st.n * mu2x), 0.0);
    double rot[3][3];
    lms_qcp_rotation(a, e0, rot);
    RotTran o;
    for (int r = 0; r < 3; r++) {
        double rx = rot[r][0] * mux[0] + rot[r][1] * mux[1] + rot[r][2] * mux[2];
        o.t[r] = (float)(muy[r] - rx);
        for (int c = 0; c < 3; c++) o.r[r][c] = (float)rot[r][c];
    }
    return o;
}
inline std::array<float, 3> lms_apply(const RotTran &q, const std::array<float, 3> &v) { // :479-499
    std::array<float, 3> w;
    for (int r = 0; r < 3; r++) w[r] = (q.r[r][0] * v[0] + q.r[r][1] * v[1] + q.r[r][2] * v[2]) + q.t[r];
    return w;
}
inline float lms_dist2(const std::array<float, 3> &a, const std::array<float, 3> &b) {
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return dx * dx + dy * dy + dz * dz;
}
float lms_quantile(std::vector<float> &v, float q) { // select_quantile_squared :451-463
    if (v.empty()) return 0.0f;
    if (v.size() == 1) return v[0];
    size_t pos = (size_t)std::round(q * (float)(v.size() - 1));
    std::nth_element(v.begin(), v.begin() + pos, v.end());
    return v[pos];
}
struct LmsRng { // SmallRng :543-564
    uint64_t state;
    explicit LmsRng(uint64_t seed) {
        uint64_t x = seed + 0x9E3779B97F4A7C15ull;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        state = x ^ (x >> 31);
    }
    uint64_t next() {
        uint64_t x = state;
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        state = x;
        return x;
    }
    size_t below(size_t end) { return (size_t)(next() % end); }
};
// LmsQcpSuperimposer::run + finish (:91-236).  mov is rotated onto ref.  Returns the inlier indices.
std::vector<size_t> lms_qcp(const std::vector<std::array<float, 3>> &mov, const std::vector<std::array<float, 3>> &ref,
                            float *U, float *T, float *rms_inliers) {
    const size_t n = mov.size();
    LmsRng rng(0xC0FFEE005EEDull);
    size_t best_seed[3] = {0, 1, 2};
    float best_q = INFINITY;
    std::vector<float> res;
    for (int trial = 0; trial < 500; trial++) {
        size_t seed[3];
        bool found = false;
        for (int tries = 0; tries < 64 && !found; tries++) { // sample_three_non_collinear :509-526
            size_t i = rng.below(n);
            size_t j = rng.below(n);
            if (j == i) j = (j + 1) % n;
            size_t k = rng.below(n);
            while (k == i || k == j) k = (k + 1) % n;
            float v1[3] = {mov[j][0] - mov[i][0], mov[j][1] - mov[i][1], mov[j][2] - mov[i][2]};
            float v2[3] = {mov[k][0] - mov[i][0], mov[k][1] - mov[i][1], mov[k][2] - mov[i][2]};
            float cx = v1[1] * v2[2] - v1[2] * v2[1], cy = v1[2] * v2[0] - v1[0] * v2[2], cz = v1[0] * v2[1] - v1[1] * v2[0];
            float area2 = cx * cx + cy * cy + cz * cz;
            if (area2 > 1e-6f) {
                seed[0] = i; seed[1] = j; seed[2] = k;
                found = true;
            }
        }
        if (!found) continue;
        LmsStats st;
        for (size_t s : seed) st.add(mov[s], ref[s]);
        RotTran q = lms_qcp_from_stats(st);
        res.clear();
        for (size_t i = 0; i < n; i++) {
            if (i == seed[0] || i == seed[1] || i == seed[2]) continue;
            res.push_back(lms_dist2(lms_apply(q, mov[i]), ref[i]));
        }
        float qv = lms_quantile(res, 0.5f);
        if (qv < best_q) {
            best_q = qv;
            best_seed[0] = seed[0]; best_seed[1] = seed[1]; best_seed[2] = seed[2];
        }
    }
    const size_t min_core = std::max<size_t>(n / 2, 3);
    const float r2_max = 2.0f * 2.0f;
    LmsStats st;
    std::vector<uint8_t> in_core(n, 0);
    std::vector<size_t> core;
    for (size_t s : best_seed) {
        st.add(mov[s], ref[s]);
        in_core[s] = 1;
        core.push_back(s);
    }
    RotTran q;
    for (;;) {
        q = lms_qcp_from_stats(st);
        bool any = false;
        size_t best_i = 0;
        float best_r2 = INFINITY;
        for (size_t i = 0; i < n; i++) {
            if (in_core[i]) continue;
            float d2 = lms_dist2(lms_apply(q, mov[i]), ref[i]);
            if (d2 < best_r2) {
                best_r2 = d2;
                best_i = i;
                any = true;
            }
        }
        if (!any) break;
        if (core.size() >= min_core && best_r2 > r2_max) break;
        st.add(mov[best_i], ref[best_i]);
        in_core[best_i] = 1;
        core.push_back(best_i);
        if (core.size() == n) break;
    }
    float sum = 0.0f;
    for (size_t i : core) sum += lms_dist2(lms_apply(q, mov[i]), ref[i]);
    *rms_inliers = std::sqrt(sum / (float)core.size());
    for (int r = 0; r < 3; r++) {
        T[r] = q.t[r];
        for (int c = 0; c < 3; c++) U[3 * r + c] = q.r[r][c];
    }
    return core;
}

// retrieve.rs:756-834 (no partial fit): CA,CB interleaved; target is rotated onto query.
void rmsd_with_calpha(const fdo_compact &query, const fdo_compact &target, const std::vector<size_t> &qi,
                      const std::vector<size_t> &ti, MatchRow &row) {
    std::vector<std::array<float, 3>> ref, mov;
    for (size_t i : qi) {
        ref.push_back({query.ca[i].x, query.ca[i].y, query.ca[i].z});
        ref.push_back({query.cb[i].x, query.cb[i].y, query.cb[i].z});
    }
    for (size_t i : ti) {
        mov.push_back({target.ca[i].x, target.ca[i].y, target.ca[i].z});
        mov.push_back({target.cb[i].x, target.cb[i].y, target.cb[i].z});
    }
    if (g_partial_fit && qi.size() > 3) // retrieve.rs:773-814: LMS-QCP above three residues, RMSD of the inlier core
        lms_qcp(mov, ref, row.U, row.T, &row.rmsd);
    else
        kabsch(mov, ref, row.U, row.T, &row.rmsd);
    similarity_metrics(ref, mov, row.U, row.T, row.metrics); // retrieve.rs:819-830
    row.target_ca.clear();
    for (size_t i : ti) row.target_ca.push_back({target.ca[i].x, target.ca[i].y, target.ca[i].z});
}

fdo_matches *retrieve(const fdo_qmap &m, const fdo_compact &query, const fdo_compact &t, uint32_t nbd,
                      uint32_t nba, float dist_cutoff, float ca_cutoff) {
    fdo_matches *res = new fdo_matches();
    res->n_query = m.indices.size();
    if (m.entries.empty()) return res;
    // prefilter_amino_acid (retrieve.rs:563-602): exact canonical residue-name match
    std::vector<size_t> set1, set2;
    if (m.entries.size() <= 200 && g_hash_type != 5 && g_hash_type != 6) { // amino_acid_index() is None for those two
        std::set<uint8_t> aa1s, aa2s;
        for (auto &e : m.entries) {
            uint32_t a1, a2;
            hash_amino_acids(e.hash, &a1, &a2);
            aa1s.insert((uint8_t)a1);
            aa2s.insert((uint8_t)a2);
        }
        for (size_t i = 0; i < t.nres(); i++) {
            for (uint8_t a : aa1s)
                if (memcmp(t.res_name[i].data(), map_u8_to_aa(a), 3) == 0) {
                    set1.push_back(i);
                    break;
                }
            for (uint8_t a : aa2s)
                if (memcmp(t.res_name[i].data(), map_u8_to_aa(a), 3) == 0) {
                    set2.push_back(i);
                    break;
                }
        }
    }
    // retrieve_with_prefilter (retrieve.rs:52-156)
    std::vector<Edge> found;
    std::vector<std::pair<size_t, std::pair<size_t, size_t>>> cand;
    std::vector<std::pair<size_t, std::pair<size_t, size_t>>> tmp;
    float f[9];
    auto process_pair = [&](size_t i, size_t j) {
        float d = calc_distance(t.ca[i], t.ca[j]);
        if (!(d <= dist_cutoff)) return;
        auto it = m.aa_dist.find({t.aa[i], t.aa[j]});
        if (it == m.aa_dist.end()) return;
        tmp.clear();
        for (auto &dq : it->second)
            if (fabsf(d - dq.first) < ca_cutoff) tmp.push_back({dq.second, {i, j}});
        if (tmp.empty()) return;
        if (pair_feature(t, i, j, dist_cutoff, f)) {
            cand.insert(cand.end(), tmp.begin(), tmp.end());
            if (!g_multi_bins.empty()) { // retrieve.rs:124-131
                for (auto &b : g_multi_bins) {
                    uint32_t h = perfect_hash_raw(f, b.first, b.second);
                    if (m.lookup.count(h)) found.push_back({i, j, h});
                }
            } else {
                uint32_t h = perfect_hash(f, nbd, nba);
                if (m.lookup.count(h)) found.push_back({i, j, h});
            }
        }
    };
    if (set1.empty() || set2.empty()) { // CombinationVecIterator::is_empty -> all pairs (combination.rs:24-44)
        for (size_t i = 0; i < t.nres(); i++) // the iterator yields all n*n pairs, diagonal included
            for (size_t j = 0; j < t.nres(); j++) process_pair(i, j);
    } else {
        for (size_t i : set1)
            for (size_t j : set2) process_pair(i, j);
    }
    for (auto &e : found) {
        res->edges.push_back({(int64_t)e.i, (int64_t)e.j});
        res->edge_hash.push_back(e.h);
    }
    std::unordered_map<size_t, std::vector<std::pair<size_t, size_t>>> cand_map;
    for (auto &c : cand) cand_map[c.first].push_back(c.second);
    // graph (graph.rs:16-27): node ids by first appearance
    std::unordered_map<size_t, size_t> node_of;
    std::vector<size_t> node_res;
    std::vector<std::pair<size_t, size_t>> gedges;
    for (auto &e : found) {
        for (size_t r : {e.i, e.j})
            if (!node_of.count(r)) {
                node_of[r] = node_res.size();
                node_res.push_back(r);
            }
        gedges.push_back({node_of[e.i], node_of[e.j]});
    }
    auto comps = components(node_res.size(), gedges, 2);
    for (auto &comp_nodes : comps) {
        std::unordered_set<size_t> in_comp;
        for (size_t nd : comp_nodes) in_comp.insert(node_res[nd]);
        std::vector<Edge> sub;
        for (auto &e : found)
            if (in_comp.count(e.i) && in_comp.count(e.j)) sub.push_back(e);
        size_t node_count = comp_nodes.size();
        float sub_idf = 0.0f; // retrieve.rs:705-719
        for (auto &e : sub) {
            const QEntry *q = m.find(e.h);
            if (q) sub_idf += q->idf;
        }
        std::vector<size_t> qidx, ridx;
        map_query_and_retrieved(sub, m, node_count, qidx, ridx);
        std::unordered_set<size_t> r_set(ridx.begin(), ridx.end());
        std::unordered_map<size_t, size_t> q2r;
        for (size_t k = 0; k < qidx.size(); k++) q2r[qidx[k]] = ridx[k]; // later pairs overwrite (collect into map)
        std::vector<size_t> q_scan, r_scan;
        std::unordered_set<size_t> r_scan_set;
        MatchRow rh, rr;
        for (size_t qi_ : m.indices) { // retrieve.rs:453-516
            std::map<size_t, size_t> count_map;
            auto f2 = q2r.find(qi_);
            if (f2 != q2r.end()) {
                size_t ri = f2->second;
                ResMatch rm{true, t.chain[ri], t.serial[ri]};
                rh.res.push_back(rm);
                if (!r_scan_set.count(ri)) {
                    rr.res.push_back(rm);
                    q_scan.push_back(qi_);
                    r_scan.push_back(ri);
                    r_scan_set.insert(ri);
                } else {
                    auto pos = std::find(r_scan.begin(), r_scan.end(), ri);
                    if (pos != r_scan.end()) {
                        size_t pp = pos - r_scan.begin();
                        rr.res[pp] = ResMatch{}; // sic: position in the scanned list indexes res_vec
                        rr.res.push_back(rm);
                        q_scan.erase(q_scan.begin() + pp);
                        r_scan.erase(r_scan.begin() + pp);
                        r_scan_set.erase(ri);
                        q_scan.push_back(qi_);
                        r_scan.push_back(ri);
                        r_scan_set.insert(ri);
                    }
                }
            } else {
                rh.res.push_back(ResMatch{});
                auto cp = cand_map.find(qi_);
                if (cp != cand_map.end())
                    for (auto &jk : cp->second)
                        if (r_set.count(jk.second)) count_map[jk.first]++;
                if (!count_map.empty()) {
                    size_t mx = 0;
                    for (auto &kv : count_map) mx = std::max(mx, kv.second);
                    std::vector<std::pair<size_t, size_t>> maxes;
                    for (auto &kv : count_map)
                        if (kv.second == mx) maxes.push_back(kv);
                    if (maxes.size() == 1 && maxes[0].second >= 2 && !r_scan_set.count(maxes[0].first)) {
                        size_t ri = maxes[0].first;
                        rr.res.push_back(ResMatch{true, t.chain[ri], t.serial[ri]});
                        q_scan.push_back(qi_);
                        r_scan.push_back(ri);
                        r_scan_set.insert(ri);
                    } else {
                        rr.res.push_back(ResMatch{});
                    }
                } else {
                    rr.res.push_back(ResMatch{});
                }
            }
        }
        rmsd_with_calpha(query, t, qidx, ridx, rh);
        rh.idf = sub_idf;
        if (rr.res == rh.res) {
            rr.rmsd = rh.rmsd;
            memcpy(rr.U, rh.U, sizeof(rh.U));
            memcpy(rr.T, rh.T, sizeof(rh.T));
            memcpy(rr.metrics, rh.metrics, sizeof(rh.metrics)); // retrieve.rs:522-523: ca_coords and metrics too
            rr.target_ca = rh.target_ca;
        } else {
            rmsd_with_calpha(query, t, q_scan, r_scan, rr);
        }
        rr.idf = sub_idf;
        res->from_hash.push_back(rh);
        res->result.push_back(rr);
    }
    for (auto &row : res->result) { // retrieve.rs:539-551
        size_t cnt = 0;
        for (auto &x : row.res) cnt += x.some;
        if (cnt > res->max_node) {
            res->max_node = cnt;
            res->min_rmsd = row.rmsd;
        } else if (cnt == res->max_node && row.rmsd < res->min_rmsd) {
            res->min_rmsd = row.rmsd;
        }
    }
    return res;
}

} // namespace

// =============================================================================================
// C API
// =============================================================================================
extern "C" {

void fdo_set_math_mode(int use_libm) { fdo_math_use_libm = use_libm; }
float fdo_math_sinf(float x) { return fdo_sinf(x); }
float fdo_math_cosf(float x) { return fdo_cosf(x); }
float fdo_math_acosf(float x) { return fdo_acosf(x); }
float fdo_math_atan2f(float y, float x) { return fdo_atan2f(y, x); }

fdo_structure *fdo_structure_read_pdb(const char *path) {
    fdo_structure *s = new fdo_structure();
    if (!read_pdb(path, s)) {
        delete s;
        return nullptr;
    }
    return s;
}
fdo_structure *fdo_structure_from_atoms(int64_t n, const float *x, const float *y, const float *z,
                                        const uint8_t *an, const uint8_t *ch, const uint8_t *rn,
                                        const uint64_t *rs, const float *bf) {
    fdo_structure *s = new fdo_structure();
    for (int64_t i = 0; i < n; i++) s->update(x[i], y[i], z[i], an + 4 * i, ch[i], rn + 3 * i, rs[i], bf[i]);
    return s;
}
int64_t fdo_structure_num_atoms(const fdo_structure *s) { return (int64_t)s->x.size(); }
int64_t fdo_structure_num_residues(const fdo_structure *s) { return (int64_t)s->num_residues; }
int fdo_structure_first_chain(const fdo_structure *s) { return s->chains.empty() ? -1 : s->chains[0]; }
void fdo_structure_get_atoms(const fdo_structure *s, float *x, float *y, float *z, uint8_t *an, uint8_t *ch,
                             uint8_t *rn, uint64_t *rs, float *bf) {
    for (size_t i = 0; i < s->x.size(); i++) {
        x[i] = s->x[i];
        y[i] = s->y[i];
        z[i] = s->z[i];
        memcpy(an + 4 * i, s->atom_name[i].data(), 4);
        ch[i] = s->chain[i];
        memcpy(rn + 3 * i, s->res_name[i].data(), 3);
        rs[i] = s->res_serial[i];
        bf[i] = s->b_factor[i];
    }
}
void fdo_structure_free(fdo_structure *s) { delete s; }

fdo_compact *fdo_compact_build(const fdo_structure *s) { return build_compact(*s); }
fdo_compact *fdo_compact_from_soa(int64_t n, const float *nx, const float *cax, const float *cbx,
                                  const uint8_t *cbv, const uint8_t *aa, const uint8_t *chain,
                                  const uint64_t *serial, const float *bf) {
    fdo_compact *c = new fdo_compact();
    for (int64_t i = 0; i < n; i++) {
        c->n.push_back({nx[3 * i], nx[3 * i + 1], nx[3 * i + 2]});
        c->ca.push_back({cax[3 * i], cax[3 * i + 1], cax[3 * i + 2]});
        c->cb.push_back({cbx[3 * i], cbx[3 * i + 1], cbx[3 * i + 2]});
        c->cb_valid.push_back(cbv ? cbv[i] : 1);
        // 128 + code: a modified residue (non-canonical name) that maps to amino acid `code`
        const bool modified = aa[i] != 255 && (aa[i] & 0x80);
        const uint8_t code = modified ? (uint8_t)(aa[i] & 0x7F) : aa[i];
        c->aa.push_back(code);
        const char *nm = modified ? "mod" : map_u8_to_aa(code);
        c->res_name.push_back({(uint8_t)nm[0], (uint8_t)nm[1], (uint8_t)nm[2]});
        c->chain.push_back(chain ? chain[i] : (uint8_t)'A');
        c->serial.push_back(serial ? serial[i] : (uint64_t)(i + 1));
        c->b_factor.push_back(bf ? bf[i] : 0.0f);
    }
    if (n > 0) c->chains.push_back(c->chain[0]);
    return c;
}
int64_t fdo_compact_nres(const fdo_compact *c) { return (int64_t)c->nres(); }
void fdo_compact_get(const fdo_compact *c, float *nx, float *cax, float *cbx, uint8_t *cbv, uint8_t *aa,
                     uint8_t *rn, uint8_t *chain, uint64_t *serial, float *bf) {
    for (size_t i = 0; i < c->nres(); i++) {
        if (nx) { nx[3 * i] = c->n[i].x; nx[3 * i + 1] = c->n[i].y; nx[3 * i + 2] = c->n[i].z; }
        if (cax) { cax[3 * i] = c->ca[i].x; cax[3 * i + 1] = c->ca[i].y; cax[3 * i + 2] = c->ca[i].z; }
        if (cbx) { cbx[3 * i] = c->cb[i].x; cbx[3 * i + 1] = c->cb[i].y; cbx[3 * i + 2] = c->cb[i].z; }
        if (cbv) cbv[i] = c->cb_valid[i];
        if (aa) aa[i] = c->aa[i];
        if (rn) memcpy(rn + 3 * i, c->res_name[i].data(), 3);
        if (chain) chain[i] = c->chain[i];
        if (serial) serial[i] = c->serial[i];
        if (bf) bf[i] = c->b_factor[i];
    }
}
// core.rs:446-456
float fdo_compact_avg_plddt(const fdo_compact *c) {
    float sum = 0.0f;
    for (float b : c->b_factor) sum += b;
    return sum / (float)c->nres();
}
// core.rs:215-223
int64_t fdo_compact_get_index(const fdo_compact *c, uint8_t chain, uint64_t serial) {
    for (size_t i = 0; i < c->nres(); i++)
        if (c->chain[i] == chain && c->serial[i] == serial) return (int64_t)i;
    return -1;
}
void fdo_compact_free(fdo_compact *c) { delete c; }
uint8_t fdo_map_aa_to_u8(const uint8_t *rn) { return map_aa_to_u8(rn); }

int fdo_pair_feature(const fdo_compact *c, int64_t i, int64_t j, float cutoff, float *out7) {
    float f[9];
    if (!pair_feature(*c, (size_t)i, (size_t)j, cutoff, f)) return 0;
    memcpy(out7, f, 7 * sizeof(float));
    return 1;
}
int fdo_pair_feature9(const fdo_compact *c, int64_t i, int64_t j, float cutoff, float *out9) {
    return pair_feature(*c, (size_t)i, (size_t)j, cutoff, out9) ? 1 : 0;
}
uint32_t fdo_perfect_hash(const float *f, uint32_t nbd, uint32_t nba) { return perfect_hash(f, nbd, nba); }
uint32_t fdo_perfect_hash_raw(const float *f9, uint32_t nbd, uint32_t nba) { return perfect_hash_raw(f9, nbd, nba); }
int fdo_set_hash_type(int t) {
    if (t < 0 || t > 8) return -1;
    g_hash_type = t;
    return 0;
}
int fdo_get_hash_type(void) { return g_hash_type; }
void fdo_set_multiple_bins(int n, const uint32_t *dist_angle_pairs) {
    g_multi_bins.clear();
    for (int k = 0; k < n; k++) g_multi_bins.push_back({dist_angle_pairs[2 * k], dist_angle_pairs[2 * k + 1]});
}
int fdo_hash_is_symmetric(uint32_t h) { return hash_is_symmetric(h) ? 1 : 0; }
int64_t fdo_hash_compact(const fdo_compact *c, uint32_t nbd, uint32_t nba, float cutoff, int su, uint32_t *out,
                         int64_t cap) {
    std::vector<uint32_t> v;
    hash_compact(*c, nbd, nba, cutoff, v);
    if (su) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    }
    for (int64_t i = 0; i < (int64_t)v.size() && i < cap; i++) out[i] = v[i];
    return (int64_t)v.size();
}

fdo_index *fdo_index_from_csr(const uint32_t *hashes, const uint64_t *ro, uint64_t S) {
    std::vector<std::pair<uint32_t, uint64_t>> pairs;
    pairs.reserve(ro[S]);
    for (uint64_t s = 0; s < S; s++)
        for (uint64_t k = ro[s]; k < ro[s + 1]; k++) pairs.push_back({hashes[k], s});
    return index_from_pairs(pairs);
}
fdo_index *fdo_index_build(const fdo_compact *const *structs, uint64_t S, uint32_t nbd, uint32_t nba,
                           float cutoff, int threads) {
    std::vector<std::vector<uint32_t>> per(S);
    std::atomic<uint64_t> next{0};
    auto work = [&] {
        for (uint64_t s; (s = next++) < S;) {
            hash_compact(*structs[s], nbd, nba, cutoff, per[s]);
            std::sort(per[s].begin(), per[s].end());
            per[s].erase(std::unique(per[s].begin(), per[s].end()), per[s].end());
        }
    };
    if (threads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    std::vector<std::pair<uint32_t, uint64_t>> pairs;
    size_t total = 0;
    for (uint64_t s = 0; s < S; s++) total += per[s].size();
    pairs.reserve(total);
    for (uint64_t s = 0; s < S; s++) {
        for (uint32_t h : per[s]) pairs.push_back({h, s});
        std::vector<uint32_t>().swap(per[s]);
    }
    return index_from_pairs(pairs, threads);
}
fdo_index *fdo_index_from_buffers(const uint32_t *hashes, const uint64_t *offsets, uint64_t count,
                                  const uint8_t *values, uint64_t vb) {
    fdo_index *ix = new fdo_index();
    ix->hashes.assign(hashes, hashes + count);
    ix->offsets.assign(offsets, offsets + count + 1);
    ix->values.assign(values, values + vb);
    return ix;
}
// indextable.rs:331-394
fdo_index *fdo_index_load(const char *prefix) {
    std::string op = std::string(prefix) + ".offset";
    std::string vp = std::string(prefix) + ".value";
    FILE *f = fopen(vp.c_str(), "rb");
    if (!f) f = fopen(prefix, "rb");
    if (!f) return nullptr;
    fdo_index *ix = new fdo_index();
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    ix->values.resize((size_t)sz);
    if (sz > 0 && fread(ix->values.data(), 1, (size_t)sz, f) != (size_t)sz) {
        fclose(f);
        delete ix;
        return nullptr;
    }
    fclose(f);
    FILE *o = fopen(op.c_str(), "rb");
    if (!o) {
        delete ix;
        return nullptr;
    }
    uint64_t count = 0;
    bool ok = fread(&count, 8, 1, o) == 1;
    if (ok) {
        ix->hashes.resize(count);
        ix->offsets.resize(count + 1);
        ok = (count == 0 || fread(ix->hashes.data(), 4, count, o) == count) &&
             fread(ix->offsets.data(), 8, count + 1, o) == count + 1;
    }
    fclose(o);
    if (!ok) {
        delete ix;
        return nullptr;
    }
    return ix;
}
// indextable.rs:239-264, 297-326
int fdo_index_save(const fdo_index *ix, const char *prefix) {
    FILE *f = fopen(prefix, "wb");
    if (!f) return -1;
    if (!ix->values.empty()) fwrite(ix->values.data(), 1, ix->values.size(), f);
    fclose(f);
    std::string op = std::string(prefix) + ".offset";
    FILE *o = fopen(op.c_str(), "wb");
    if (!o) return -1;
    uint64_t count = ix->hashes.size();
    fwrite(&count, 8, 1, o);
    if (count) fwrite(ix->hashes.data(), 4, count, o);
    fwrite(ix->offsets.data(), 8, count + 1, o);
    fclose(o);
    return 0;
}
uint64_t fdo_index_count(const fdo_index *ix) { return ix->hashes.size(); }
uint64_t fdo_index_value_bytes(const fdo_index *ix) { return ix->values.size(); }
const uint32_t *fdo_index_hashes(const fdo_index *ix) { return ix->hashes.data(); }
const uint64_t *fdo_index_offsets(const fdo_index *ix) { return ix->offsets.data(); }
const uint8_t *fdo_index_values(const fdo_index *ix) { return ix->values.data(); }
int64_t fdo_index_get_entries(const fdo_index *ix, uint32_t h, uint64_t *out, int64_t cap) {
    std::vector<uint64_t> v;
    ix->entries(h, v);
    for (int64_t i = 0; i < (int64_t)v.size() && i < cap; i++) out[i] = v[i];
    return (int64_t)v.size();
}
void fdo_index_free(fdo_index *ix) { delete ix; }

// index/lookup.rs:17-58: id \t path \t nres \t plddt \t db_key
int fdo_lookup_save(const char *path, uint64_t n, const char *const *names, const uint64_t *nres,
                    const float *plddt) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    for (uint64_t i = 0; i < n; i++)
        fprintf(f, "%llu\t%s\t%llu\t%s\t%llu\n", (unsigned long long)i, names[i], (unsigned long long)nres[i],
                rust_f32(plddt[i]).c_str(), (unsigned long long)i);
    fclose(f);
    return 0;
}
// cli/config.rs:64-97 via the `toml` crate: keys of a table serialise in sorted (BTreeMap) order,
// floats always carry a fractional part.
int fdo_type_save(const char *path, uint32_t nbd, uint32_t nba, float grid, uint64_t chunk, uint64_t maxres,
                  const char *foldcomp_db) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    std::string g;
    {
        char buf[64];
        auto r = std::to_chars(buf, buf + sizeof(buf), (double)grid, std::chars_format::fixed);
        g = std::string(buf, r.ptr);
        if (g.find('.') == std::string::npos) g += ".0";
    }
    fprintf(f, "chunk_size = %llu\n", (unsigned long long)chunk);
    if (foldcomp_db) fprintf(f, "foldcomp_db = \"%s\"\n", foldcomp_db);
    fprintf(f,
            "grid_width = %s\nhash_type = \"PDBTrRosetta\"\ninput_format = \"PDB\"\n"
            "max_residue = %llu\nnum_bin_angle = %u\nnum_bin_dist = %u\n",
            g.c_str(), (unsigned long long)maxres, nba, nbd);
    fclose(f);
    return 0;
}

// query.rs:331-384
int64_t fdo_parse_query_string(const char *q, uint8_t default_chain, uint8_t *chains, uint64_t *serials,
                               int64_t *subs_off, uint8_t *subs, int64_t cap_res, int64_t cap_subs) {
    std::string s;
    for (const char *p = q; *p; p++)
        if (*p != ' ') s.push_back(*p);
    if (s.empty()) return 0;
    if (!isalpha(default_chain)) default_chain = 'A';
    int64_t n = 0, ns = 0;
    auto one_letter = [](char c, std::vector<uint8_t> &o) { // convert.rs:223-262
        static const char *std20 = "ARNDCQEGHILKMFPSTWYV";
        const char *p = strchr(std20, c);
        if (p && c) { o.push_back((uint8_t)(p - std20)); return; }
        auto all = [&] { for (int i = 0; i < 20; i++) o.push_back((uint8_t)i); };
        switch (c) {
            case 'B': o.insert(o.end(), {2, 3}); break;
            case 'Z': o.insert(o.end(), {5, 6}); break;
            case 'X': case 'x': all(); break;
            case 'J': o.insert(o.end(), {9, 10}); break;
            case 'U': o.push_back(4); break;
            case 'O': o.push_back(11); break;
            case 'p': o.insert(o.end(), {1, 8, 11}); break;
            case 'n': o.insert(o.end(), {3, 6}); break;
            case 'h': o.insert(o.end(), {2, 5, 15, 16, 18}); break;
            case 'b': o.insert(o.end(), {0, 4, 7, 9, 10, 12, 13, 14, 19}); break;
            case 'a': o.insert(o.end(), {8, 13, 17, 18}); break;
            default: o.push_back(255);
        }
    };
    size_t pos = 0;
    while (pos <= s.size()) {
        size_t comma = s.find(',', pos);
        std::string seg = s.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
        pos = comma == std::string::npos ? s.size() + 1 : comma + 1;
        uint8_t chain = default_chain;
        std::string rest = seg;
        if (!seg.empty() && isalpha((unsigned char)seg[0])) {
            chain = (uint8_t)seg[0];
            rest = seg.substr(1);
        }
        bool has_sub = false;
        std::vector<uint8_t> sub;
        size_t colon = rest.find(':');
        std::string range = rest;
        if (colon != std::string::npos) {
            has_sub = true;
            range = rest.substr(0, colon);
            for (char c : rest.substr(colon + 1))
                if (isalpha((unsigned char)c)) one_letter(c, sub);
        }
        uint64_t a, b;
        size_t dash = range.find('-');
        if (dash != std::string::npos) {
            if (!parse_u64(range.substr(0, dash), &a) || !parse_u64(range.substr(dash + 1), &b)) return -1;
        } else {
            if (!parse_u64(range, &a)) return -1;
            b = a;
        }
        for (uint64_t r = a; r <= b; r++) {
            if (n >= cap_res) return -2;
            chains[n] = chain;
            serials[n] = r;
            if (has_sub) {
                subs_off[n] = ns;
                for (uint8_t v : sub) {
                    if (ns >= cap_subs) return -2;
                    subs[ns++] = v;
                }
            } else {
                subs_off[n] = -1;
            }
            n++;
        }
    }
    subs_off[n] = ns;
    return n;
}

// query.rs:208-329
fdo_qmap *fdo_qmap_make(const fdo_compact *c, const uint8_t *chains, const uint64_t *serials, int64_t n_res,
                        const int64_t *sub_off, const int64_t *sub_end, const uint8_t *subs, uint32_t nbd,
                        uint32_t nba, const float *dist_thr, int n_dt, const float *angle_thr, int n_at,
                        float cutoff, int serial_query, const fdo_index *ix, float total) {
    fdo_qmap *m = new fdo_qmap();
    std::vector<std::pair<uint8_t, uint64_t>> qres;
    std::vector<std::vector<uint8_t>> qsub;
    std::vector<uint8_t> qhas;
    for (int64_t i = 0; i < n_res; i++) {
        qres.push_back({chains[i], serials[i]});
        bool has = sub_off && sub_off[i] >= 0;
        qhas.push_back(has);
        qsub.emplace_back();
        if (has) qsub.back().assign(subs + sub_off[i], subs + sub_end[i]);
    }
    if (qres.empty()) {
        for (size_t i = 0; i < c->nres(); i++) {
            qres.push_back({c->chain[i], c->serial[i]});
            qhas.push_back(0);
            qsub.emplace_back();
        }
    }
    m->residue_count = qres.size();
    std::unordered_map<size_t, std::vector<uint8_t>> submap;
    for (size_t i = 0; i < qres.size(); i++) {
        int64_t idx = serial_query ? (int64_t)qres[i].second : fdo_compact_get_index(c, qres[i].first, qres[i].second);
        if (idx < 0) continue;
        m->indices.push_back((size_t)idx);
        if (qhas[i]) submap[(size_t)idx] = qsub[i];
    }
    float f[9] = {0}, fn[9], ff[9];
    const size_t K = m->indices.size();
    const float rad = 3.14159274101257324f / 180.0f; // f32::to_radians
    for (size_t a = 0; a < K; a++)
        for (size_t b = 0; b < K; b++) {
            if (a == b) continue;
            size_t I = m->indices[a], J = m->indices[b];
            if (I >= c->nres() || J >= c->nres()) continue; // the reference would panic (serial_query out of range)
            if (!pair_feature(*c, I, J, cutoff, f)) continue;
            memcpy(fn, f, sizeof(f));
            memcpy(ff, f, sizeof(f));
            { // core.rs:458-473
                float d = calc_distance(c->ca[I], c->ca[J]);
                if (d <= 20.0f) m->aa_dist[{c->aa[I], c->aa[J]}].push_back({d, I});
            }
            uint32_t observed = perfect_hash(f, nbd, nba);
            float idf = idf_for_hash(observed, ix, total);
            insert_binned_hash(*m, f, I, J, nbd, nba, true, idf);
            if (g_hash_type != 5 && g_hash_type != 6) { // apply_substitutions, query.rs:86-156 (only with amino_acid_index)
                float o1 = fn[0], o2 = fn[1];
                auto si = submap.find(I), sj = submap.find(J);
                if (si != submap.end()) {
                    for (uint8_t s : si->second) {
                        float tf[9];
                        memcpy(tf, fn, sizeof(tf));
                        tf[0] = (float)s;
                        insert_binned_hash(*m, tf, I, J, nbd, nba, false, idf);
                    }
                    if (sj != submap.end())
                        for (uint8_t s : si->second)
                            for (uint8_t s2 : sj->second) {
                                fn[0] = (float)s;
                                fn[1] = (float)s2;
                                insert_binned_hash(*m, fn, I, J, nbd, nba, false, idf);
                                fn[0] = o1;
                                fn[1] = o2;
                            }
                } else if (sj != submap.end()) {
                    for (uint8_t s : sj->second) {
                        float tf[9];
                        memcpy(tf, fn, sizeof(tf));
                        tf[1] = (float)s;
                        insert_binned_hash(*m, tf, I, J, nbd, nba, false, idf);
                    }
                }
            }
            auto expand = [&](const int *idxs, int nidx, const float *thr, int nthr, float scale, bool use_scale) {
                for (int t = 0; t < nthr; t++) { // query.rs:179-206
                    float delta = use_scale ? thr[t] * scale : thr[t];
                    for (int k = 0; k < nidx; k++) {
                        int ix2 = idxs[k];
                        fn[ix2] -= delta;
                        ff[ix2] += delta;
                        insert_binned_hash(*m, fn, I, J, nbd, nba, false, idf);
                        insert_binned_hash(*m, ff, I, J, nbd, nba, false, idf);
                        fn[ix2] += delta;
                        ff[ix2] -= delta;
                    }
                }
            };
            // HashType::dist_index / angle_index (feature.rs:269-289)
            int di[2] = {2, 3}, ai[7] = {4, 5, 6, 0, 0, 0, 0}, ndi = 2, nai = 3;
            if (g_hash_type == 5) {
                di[0] = 7, ndi = 1;
                for (int k = 0; k < 7; k++) ai[k] = k;
                nai = 7;
            }
            if (g_hash_type == 6) {
                for (int k = 0; k < 5; k++) ai[k] = 4 + k;
                nai = 5;
            }
            if (g_hash_type == 2 || g_hash_type == 4) ndi = 1;
            if (g_hash_type == 0 || g_hash_type == 1) nai = 1;
            if (g_hash_type == 2) {
                for (int k = 0; k < 5; k++) ai[k] = 3 + k;
                nai = 5;
            }
            if (g_hash_type == 4) ai[0] = 3, ai[1] = 4, ai[2] = 5;
            expand(di, ndi, dist_thr, n_dt, 1.0f, false);
            expand(ai, nai, angle_thr, n_at, rad, true);
        }
    return m;
}
int64_t fdo_qmap_size(const fdo_qmap *m) { return (int64_t)m->entries.size(); }
void fdo_qmap_get(const fdo_qmap *m, uint32_t *hash, int64_t *qi, int64_t *qj, uint8_t *primary, float *idf) {
    for (size_t i = 0; i < m->entries.size(); i++) {
        if (hash) hash[i] = m->entries[i].hash;
        if (qi) qi[i] = (int64_t)m->entries[i].qi;
        if (qj) qj[i] = (int64_t)m->entries[i].qj;
        if (primary) primary[i] = m->entries[i].primary;
        if (idf) idf[i] = m->entries[i].idf;
    }
}
int64_t fdo_qmap_num_indices(const fdo_qmap *m) { return (int64_t)m->indices.size(); }
int64_t fdo_qmap_residue_count(const fdo_qmap *m) { return (int64_t)m->residue_count; }
void fdo_qmap_get_indices(const fdo_qmap *m, int64_t *indices) {
    for (size_t i = 0; i < m->indices.size(); i++) indices[i] = (int64_t)m->indices[i];
}
void fdo_qmap_free(fdo_qmap *m) { delete m; }

fdo_hits *fdo_count_query(const fdo_qmap *m, const fdo_index *ix, uint64_t S, const uint64_t *nres,
                          const float *plddt, const fdo_count_params *p) {
    fdo_hits *h = new fdo_hits();
    count_query(*m, *ix, S, nres, *p, 1, h->v, nullptr);
    if (p->apply_filter_and_sort) filter_sort_top(h->v, nres, plddt, *p);
    return h;
}
int64_t fdo_hits_size(const fdo_hits *h) { return (int64_t)h->v.size(); }
void fdo_hits_get(const fdo_hits *h, uint64_t *nid, uint32_t *mc, uint32_t *nc, uint32_t *ec, float *idf) {
    for (size_t i = 0; i < h->v.size(); i++) {
        if (nid) nid[i] = h->v[i].nid;
        if (mc) mc[i] = h->v[i].match_count;
        if (nc) nc[i] = h->v[i].node_count;
        if (ec) ec[i] = h->v[i].edge_count;
        if (idf) idf[i] = h->v[i].idf;
    }
}
void fdo_hits_free(fdo_hits *h) { delete h; }

fdo_matches *fdo_retrieve(const fdo_qmap *m, const fdo_compact *q, const fdo_compact *t, uint32_t nbd,
                          uint32_t nba, float cutoff, float ca_cutoff) {
    return retrieve(*m, *q, *t, nbd, nba, cutoff, ca_cutoff);
}
int64_t fdo_matches_size(const fdo_matches *r) { return (int64_t)r->result.size(); }
int64_t fdo_matches_num_query(const fdo_matches *r) { return (int64_t)r->n_query; }
void fdo_matches_get(const fdo_matches *r, int which, uint8_t *some, uint8_t *chain, uint64_t *serial,
                     float *rmsd, float *idf, float *U, float *t) {
    const auto &rows = which ? r->from_hash : r->result;
    for (size_t k = 0; k < rows.size(); k++) {
        for (size_t q = 0; q < r->n_query; q++) {
            const ResMatch &x = rows[k].res[q];
            if (some) some[k * r->n_query + q] = x.some;
            if (chain) chain[k * r->n_query + q] = x.chain;
            if (serial) serial[k * r->n_query + q] = x.serial;
        }
        if (rmsd) rmsd[k] = rows[k].rmsd;
        if (idf) idf[k] = rows[k].idf;
        if (U) memcpy(U + 9 * k, rows[k].U, 9 * sizeof(float));
        if (t) memcpy(t + 3 * k, rows[k].T, 3 * sizeof(float));
    }
}
void fdo_set_partial_fit(int on) { g_partial_fit = on ? 1 : 0; }
int64_t fdo_lms_qcp(int64_t n, const float *ref3, const float *mov3, float *U9, float *t3, float *rms_inliers,
                    int64_t *inliers, int64_t cap) {
    std::vector<std::array<float, 3>> ref((size_t)n), mov((size_t)n);
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            ref[i][k] = ref3[3 * i + k];
            mov[i][k] = mov3[3 * i + k];
        }
    std::vector<size_t> core = lms_qcp(mov, ref, U9, t3, rms_inliers);
    for (size_t k = 0; k < core.size() && (int64_t)k < cap; k++) inliers[k] = (int64_t)core[k];
    return (int64_t)core.size();
}
void fdo_matches_get_metrics(const fdo_matches *r, int which, float *out5) {
    const auto &rows = which ? r->from_hash : r->result;
    for (size_t k = 0; k < rows.size(); k++) memcpy(out5 + 5 * k, rows[k].metrics, 5 * sizeof(float));
}
void fdo_similarity_metrics(int64_t n, const float *ref3, const float *mov3, const float *U9, const float *t3, float *out5) {
    std::vector<std::array<float, 3>> ref((size_t)n), mov((size_t)n);
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            ref[i][k] = ref3[3 * i + k];
            mov[i][k] = mov3[3 * i + k];
        }
    similarity_metrics(ref, mov, U9, t3, out5);
}
int64_t fdo_matches_max_node_count(const fdo_matches *r) { return (int64_t)r->max_node; }
float fdo_matches_min_rmsd(const fdo_matches *r) { return r->min_rmsd; }
int64_t fdo_matches_num_edges(const fdo_matches *r) { return (int64_t)r->edges.size(); }
void fdo_matches_get_edges(const fdo_matches *r, int64_t *ei, int64_t *ej, uint32_t *eh) {
    for (size_t k = 0; k < r->edges.size(); k++) {
        ei[k] = r->edges[k][0];
        ej[k] = r->edges[k][1];
        eh[k] = r->edge_hash[k];
    }
}
void fdo_matches_free(fdo_matches *r) { delete r; }

int fdo_kabsch(int64_t n, const float *x3, const float *y3, float *U9, float *t3, float *rmsd) {
    std::vector<std::array<float, 3>> x((size_t)n), y((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        x[i] = {x3[3 * i], x3[3 * i + 1], x3[3 * i + 2]};
        y[i] = {y3[3 * i], y3[3 * i + 1], y3[3 * i + 2]};
    }
    return kabsch(x, y, U9, t3, rmsd) ? 0 : -1;
}

int64_t fdo_query_batch(const fdo_qmap *const *maps, const fdo_compact *const *queries, int64_t n_q,
                        const fdo_index *ix, const fdo_compact *const *store, uint64_t S, const uint64_t *nres,
                        const float *plddt, const fdo_count_params *p, uint32_t nbd, uint32_t nba, float cutoff,
                        float ca_cutoff, int skip_match, int threads, uint64_t *hits_pq, uint64_t *matches_pq,
                        uint64_t *bytes_pq) {
    // query_pdb.rs:348: queries.into_par_iter(); inside each, count_query is node-group parallel and
    // retrieval is candidate parallel, all on one rayon pool.  Here: queries are distributed over the
    // threads (the batch is much wider than the core count), each worker runs its query serially.
    std::atomic<int64_t> next{0};
    std::atomic<int64_t> total{0};
    auto work = [&] {
        for (int64_t q; (q = next++) < n_q;) {
            std::vector<Hit> hits;
            uint64_t bytes = 0;
            fdo_count_params pp = *p;
            pp.expected_node_count = maps[q]->residue_count;
            count_query(*maps[q], *ix, S, nres, pp, 1, hits, &bytes);
            filter_sort_top(hits, nres, plddt, pp);
            uint64_t nm = 0;
            if (!skip_match)
                for (auto &h : hits) {
                    fdo_matches *r = retrieve(*maps[q], *queries[q], *store[h.nid], nbd, nba, cutoff, ca_cutoff);
                    nm += r->result.size();
                    delete r;
                }
            if (hits_pq) hits_pq[q] = hits.size();
            if (matches_pq) matches_pq[q] = nm;
            if (bytes_pq) bytes_pq[q] = bytes;
            total += (int64_t)nm;
        }
    };
    if (threads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    return total;
}

} // extern "C"
