// fd_host.cpp -- host side above the C ABI (include/folddisco_b200_host.h).
//
// Mirrors the reference's host orchestration for the hot path.  All heavy lifting is delegated to the CUDA
// kernels through the fd_* entry points; what runs here is parsing, query-map construction for the k(k-1)
// query pairs, and the small irregular graph / residue-assignment step between K4 and K5.
// Compiled by nvcc (-x cu) so that fd_geom.cuh is the same source the kernels use.
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <functional>
#include <array>
#include <atomic>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../../include/folddisco_b200_host.h"
#include "../fd_geom.cuh"
#include "../fd_hashtypes.cuh"

// persistent host worker pool (fd_ctx.cu): fn(0) .. fn(nt - 1) concurrently, fn(0) on the caller
void fd_parallel(int nt, const std::function<void(int)> &fn);

namespace {
thread_local std::string g_err;
void set_err(const std::string &s) { g_err = s; }

const char *CANON[20] = {"ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE",
                         "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL"};

// map_aa_to_u8 (src/utils/convert.rs:53-81) as a sorted table of three-letter codes.
struct AaName {
    char n[4];
    uint8_t v;
};
const AaName AA_TABLE[] = {
    {"0AF", 17}, {"0TD", 3},  {"ABA", 0},  {"AGM", 1},  {"AIB", 0},  {"ALA", 0},  {"ALC", 0},  {"ALY", 11},
    {"ARG", 1},  {"ASN", 2},  {"ASP", 3},  {"ASX", 3},  {"B3E", 6},  {"BFD", 3},  {"BMT", 16}, {"CAF", 4},
    {"CAS", 4},  {"CGU", 6},  {"CIR", 1},  {"CME", 4},  {"CR2", 7},  {"CR8", 8},  {"CRF", 16}, {"CRO", 16},
    {"CRQ", 5},  {"CSD", 4},  {"CSH", 15}, {"CSO", 4},  {"CSS", 4},  {"CSX", 4},  {"CXM", 12}, {"CYS", 4},
    {"DAB", 0},  {"DAL", 0},  {"DAR", 1},  {"DAS", 3},  {"DCY", 4},  {"DGL", 6},  {"DGN", 5},  {"DHA", 15},
    {"DHI", 8},  {"DIL", 9},  {"DLE", 10}, {"DLY", 11}, {"DPN", 13}, {"DPR", 14}, {"DSG", 2},  {"DSN", 15},
    {"DTH", 16}, {"DTR", 17}, {"DTY", 18}, {"DVA", 19}, {"FGA", 6},  {"FME", 12}, {"FVA", 19}, {"GHP", 7},
    {"GL3", 7},  {"GLN", 5},  {"GLU", 6},  {"GLX", 6},  {"GLY", 7},  {"GYS", 15}, {"HIC", 8},  {"HIS", 8},
    {"HYP", 14}, {"IAS", 3},  {"ILE", 9},  {"KCX", 11}, {"KPI", 11}, {"LEU", 10}, {"LLP", 11}, {"LYS", 11},
    {"M3L", 11}, {"MAA", 0},  {"MDO", 0},  {"MEA", 13}, {"MED", 12}, {"MEN", 2},  {"MEQ", 5},  {"MET", 12},
    {"MHO", 12}, {"MHS", 8},  {"MK8", 10}, {"MLE", 10}, {"MLY", 11}, {"MLZ", 11}, {"MSE", 12}, {"MVA", 19},
    {"NEP", 8},  {"NLE", 10}, {"NRQ", 12}, {"OAS", 15}, {"OCS", 4},  {"OMY", 18}, {"ORN", 0},  {"PCA", 6},
    {"PHD", 3},  {"PHE", 13}, {"PHI", 13}, {"PHL", 13}, {"PRO", 14}, {"PTR", 18}, {"PYL", 11}, {"SAC", 15},
    {"SAR", 7},  {"SCH", 4},  {"SCY", 4},  {"SEC", 4},  {"SEP", 15}, {"SER", 15}, {"SMC", 4},  {"SME", 12},
    {"SNC", 4},  {"SNN", 2},  {"THR", 16}, {"TOX", 17}, {"TPO", 16}, {"TPQ", 18}, {"TRP", 17}, {"TRQ", 17},
    {"TYR", 18}, {"TYS", 18}, {"VAL", 19}, {"YCM", 4},
};
// -> fd_struct_batch convention: code, 128 + code for a non-canonical name, 255 unknown
uint8_t aa_code(const uint8_t *name) {
    int lo = 0, hi = (int)(sizeof(AA_TABLE) / sizeof(AA_TABLE[0])) - 1;
    while (lo <= hi) {
        int mid = (lo + hi) / 2;
        int c = memcmp(AA_TABLE[mid].n, name, 3);
        if (c == 0) {
            uint8_t v = AA_TABLE[mid].v;
            return memcmp(CANON[v], name, 3) == 0 ? v : (uint8_t)(128 + v);
        }
        if (c < 0) lo = mid + 1;
        else hi = mid - 1;
    }
    return 255;
}

struct Atoms {
    std::vector<float> x, y, z, b;
    std::vector<uint8_t> name; // 4 per atom
    std::vector<uint8_t> rname; // 3 per atom
    std::vector<uint8_t> chain;
    std::vector<uint64_t> serial;
    size_t size() const { return x.size(); }
};

} // namespace

struct fdh_compact {
    std::vector<float> n, ca, cb; // 3 per residue
    std::vector<uint8_t> cb_valid, aa, chain;
    std::vector<uint64_t> serial;
    std::vector<float> bfac;
    std::vector<uint8_t> chains;
    uint64_t raw_residues = 0;
    // A handle returned by the library owns itself through `self` (adopt_compact); fdh_compact_free drops that
    // reference, and query batches that were given the handle hold their own: a batch shares the structure instead of
    // copying it (a structure is immutable once built) and stays valid after the caller frees the handle.  Copies of
    // the object do not inherit the reference.
    struct SelfRef {
        std::shared_ptr<const fdh_compact> p;
        SelfRef() = default;
        SelfRef(const SelfRef &) {}
        SelfRef &operator=(const SelfRef &) { return *this; }
    } self;
    size_t nres() const { return aa.size(); }
    fdg::V3 N(size_t i) const { return {n[3 * i], n[3 * i + 1], n[3 * i + 2]}; }
    fdg::V3 CA(size_t i) const { return {ca[3 * i], ca[3 * i + 1], ca[3 * i + 2]}; }
    fdg::V3 CB(size_t i) const { return {cb[3 * i], cb[3 * i + 1], cb[3 * i + 2]}; }
};

namespace {

fdh_compact *adopt_compact(fdh_compact *c) {
    c->self.p.reset(c);
    return c;
}
// the structure behind a handle: shared (handles made by the library) or copied (any other object)
std::shared_ptr<const fdh_compact> share_compact(const fdh_compact *c) {
    return c->self.p ? c->self.p : std::make_shared<const fdh_compact>(*c);
}

// virtual C-beta (src/structure/coordinate.rs:167-186), f32, same operation order
fdg::V3 approx_cb(fdg::V3 ca, fdg::V3 n, fdg::V3 c) {
    using namespace fdg;
    auto add = [](V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; };
    auto scale = [](V3 a, float f) { return V3{a.x * f, a.y * f, a.z * f}; };
    V3 v1 = normalize(sub(c, ca));
    V3 v2 = normalize(sub(n, ca));
    V3 b1 = add(v2, scale(v1, 1.0f / 3.0f));
    V3 b2 = cross(v1, b1);
    V3 u1 = normalize(b1), u2 = normalize(b2);
    V3 v4 = sub(scale(u1, -1.0f / 2.0f), scale(u2, sqrtf(3.0f) / 2.0f));
    v4 = scale(v4, sqrtf(8.0f) / 3.0f);
    v4 = add(v4, scale(v1, -1.0f / 3.0f));
    return add(ca, scale(v4, 1.5336f));
}

// CompactStructure::build (src/structure/core.rs:70-214).  A residue is flushed when res_serial changes or at
// the last atom index; chain and B-factor come from the atom that triggered the flush; the backbone C used for
// a virtual CB is the most recent one seen (never reset); residues without N or CA are dropped.
fdh_compact *compact_from_atoms(const Atoms &a) {
    fdh_compact *c = new fdh_compact();
    { // Structure::update bookkeeping (core.rs:29-43)
        uint8_t rc = ' ';
        uint64_t rs = 0;
        for (size_t i = 0; i < a.size(); i++) {
            if (rc != a.chain[i]) {
                c->chains.push_back(a.chain[i]);
                rc = a.chain[i];
            }
            if (rs != a.serial[i]) {
                c->raw_residues++;
                rs = a.serial[i];
            }
        }
    }
    enum { HAS_N = 1, HAS_CA = 2, HAS_CB = 4 };
    int have = 0;
    bool have_c = false, started = false;
    fdg::V3 n{}, ca{}, cb{}, cc{};
    uint64_t cur_serial = 0;
    size_t cur_first = 0;
    const size_t na = a.size();
    for (size_t idx = 0; idx < na; idx++) {
        if (!started || cur_serial != a.serial[idx] || idx + 1 == na) {
            if ((have & HAS_N) && (have & HAS_CA)) {
                fdg::V3 cbv{0.f, 0.f, 0.f};
                uint8_t ok = 1;
                if (have & HAS_CB) cbv = cb;
                else if (have_c) cbv = approx_cb(ca, n, cc);
                else ok = 0;
                c->n.insert(c->n.end(), {n.x, n.y, n.z});
                c->ca.insert(c->ca.end(), {ca.x, ca.y, ca.z});
                c->cb.insert(c->cb.end(), {cbv.x, cbv.y, cbv.z});
                c->cb_valid.push_back(ok);
                c->aa.push_back(aa_code(&a.rname[3 * cur_first]));
                c->serial.push_back(cur_serial);
                c->chain.push_back(a.chain[idx]);
                c->bfac.push_back(a.b[idx]);
            }
            have = 0;
            started = true;
            cur_serial = a.serial[idx];
            cur_first = idx;
        }
        const uint8_t *nm = &a.name[4 * idx];
        const fdg::V3 p{a.x[idx], a.y[idx], a.z[idx]};
        if (memcmp(nm, " CA ", 4) == 0) {
            ca = p;
            have |= HAS_CA;
        } else if (memcmp(nm, " CB ", 4) == 0) {
            cb = p;
            have |= HAS_CB;
        } else if (memcmp(nm, " C  ", 4) == 0) {
            cc = p;
            have_c = true;
        } else if (memcmp(nm, " N  ", 4) == 0) {
            n = p; // GLY and non-GLY both end up here (core.rs:176-183)
            have |= HAS_N;
        }
    }
    return c;
}

bool field_f32(const char *s, size_t len, float *out) {
    while (len && isspace((unsigned char)*s)) s++, len--;
    while (len && isspace((unsigned char)s[len - 1])) len--;
    if (!len || len > 63) return false;
    { // plain decimals of at most 7 digits (every coordinate / B-factor column of a PDB file): m / 10^e in binary64 and
      // one rounding to f32 is the correctly rounded value -- m < 2^24 and 10^e <= 10^7 are exact, and the quotient is
      // either exactly a tie of two floats or more than 2^-49 (relative) away from one, out of reach of the 2^-53 of the
      // division -- i.e. what strtof / Rust's f32::from_str return (checked against strtof for all 10^7 mantissas
      // with three decimals, the %8.3f columns, and a 1-in-7 sample of the other seven decimal positions: no
      // difference).  Anything else takes strtof below.
        static const double P10[8] = {1.0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7};
        size_t i = 0;
        const bool neg = s[0] == '-';
        if (s[0] == '-' || s[0] == '+') i = 1;
        uint32_t m = 0;
        int digits = 0, frac = -1;
        for (; i < len; i++) {
            const unsigned d = (unsigned)(s[i] - '0');
            if (d <= 9) {
                m = m * 10 + d;
                digits++;
                if (frac >= 0) frac++;
            } else if (s[i] == '.' && frac < 0) {
                frac = 0;
            } else {
                digits = 99;
                break;
            }
        }
        if (digits >= 1 && digits <= 7) {
            const float v = (float)((double)m / P10[frac < 0 ? 0 : frac]);
            *out = neg ? -v : v;
            return true;
        }
    }
    char buf[64];
    memcpy(buf, s, len);
    buf[len] = 0;
    for (size_t i = 0; i < len; i++)
        if (buf[i] == 'x' || buf[i] == 'X' || buf[i] == '(' || isspace((unsigned char)buf[i])) return false;
    char *end = nullptr;
    float v = strtof(buf, &end);
    if (end != buf + len) return false;
    *out = v;
    return true;
}
bool field_u64(const char *s, size_t len, uint64_t *out) {
    while (len && isspace((unsigned char)*s)) s++, len--;
    while (len && isspace((unsigned char)s[len - 1])) len--;
    if (len && *s == '+') s++, len--;
    if (!len) return false;
    uint64_t v = 0;
    for (size_t i = 0; i < len; i++) {
        if (s[i] < '0' || s[i] > '9') return false;
        uint64_t nv = v * 10 + (uint64_t)(s[i] - '0');
        if (nv / 10 != v) return false;
        v = nv;
    }
    *out = v;
    return true;
}

// Whole file as text; .gz (and plain) through zlib like the reference's flate2 GzDecoder (pdb.rs:96-118, cif.rs:62-94)
bool read_file_text(const char *path, bool gz, std::string &data) {
    char buf[1 << 16];
    if (gz) {
        gzFile g = gzopen(path, "rb");
        if (!g) return false;
        if (gzdirect(g)) { // not a gzip stream: zlib would hand the bytes through; the reference's GzDecoder fails
            gzclose(g);
            return false;
        }
        gzbuffer(g, 1u << 18);
        {
            struct stat st;
            if (stat(path, &st) == 0 && st.st_size > 0) data.reserve((size_t)st.st_size * 6); // text deflates ~4-5x
        }
        int r;
        while ((r = gzread(g, buf, sizeof(buf))) > 0) data.append(buf, (size_t)r);
        const bool ok = r == 0;
        gzclose(g);
        return ok;
    }
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    struct stat st;
    if (fstat(fileno(f), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) { // one read into the final buffer
        data.resize((size_t)st.st_size);
        const size_t got = fread(&data[0], 1, data.size(), f);
        data.resize(got);
    }
    size_t r;
    while ((r = fread(buf, 1, sizeof(buf), f)) > 0) data.append(buf, r); // (a file that grew, or not a regular file)
    fclose(f);
    return true;
}

// PDB reader: src/structure/io/pdb.rs:37-76 + parser.rs:3-55 (first model, ATOM records only; records whose
// fixed columns do not parse are skipped, negative residue numbers included)
bool parse_pdb_atoms(const std::string &data, Atoms &a) {
    int model = 0;
    size_t pos = 0;
    { // an ATOM record is 81 bytes
        const size_t guess = data.size() / 81 + 1;
        a.x.reserve(guess);
        a.y.reserve(guess);
        a.z.reserve(guess);
        a.b.reserve(guess);
        a.name.reserve(4 * guess);
        a.rname.reserve(3 * guess);
        a.chain.reserve(guess);
        a.serial.reserve(guess);
    }
    while (pos < data.size()) {
        const char *nlp = (const char *)memchr(data.data() + pos, '\n', data.size() - pos);
        size_t nl = nlp ? (size_t)(nlp - data.data()) : data.size();
        size_t len = nl - pos;
        const char *l = data.data() + pos;
        pos = nl + 1;
        if (len && l[len - 1] == '\r') len--;
        if (model > 1) break;
        if (len < 6) continue;
        if (memcmp(l, "MODEL ", 6) == 0) {
            model++;
            continue;
        }
        if (memcmp(l, "ATOM  ", 6) != 0 || len < 54) continue;
        float x, y, z, b = 1.0f;
        uint64_t as, rs;
        if (!field_f32(l + 30, 8, &x) || !field_f32(l + 38, 8, &y) || !field_f32(l + 46, 8, &z)) continue;
        if (!field_u64(l + 6, 5, &as) || !field_u64(l + 22, 4, &rs)) continue;
        if (len >= 66 && !field_f32(l + 60, 6, &b)) continue;
        a.x.push_back(x);
        a.y.push_back(y);
        a.z.push_back(z);
        a.b.push_back(b);
        a.name.insert(a.name.end(), l + 12, l + 16);
        a.rname.insert(a.rname.end(), l + 17, l + 20);
        a.chain.push_back((uint8_t)l[21]);
        a.serial.push_back(rs);
    }
    return true;
}

// mmCIF reader: src/structure/io/cif.rs:97-287 over the `_atom_site` loop (lexing of pdbtbx_cif 0.x: whitespace-separated
// values, '...' / "..." quoting, # comments, ; text fields).  Every row of the loop becomes an atom -- the reference
// does not separate HETATM here (cif.rs:262-270) -- until the model number changes (:239-245).  Residue number =
// auth_seq_id, else label_seq_id; chain = a one-character auth_asym_id, else label_asym_id (:253-259); B = 1.0 when absent.
bool parse_cif_atoms(const std::string &data, Atoms &a, std::string &err) {
    struct Tok { // a view into `data` (no copy per value: an atom_site loop has ~20 values per atom)
        std::string_view text;
        bool quoted;
    };
    const std::string_view dv(data);
    auto is_ws = [](char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }; // isspace of the C locale, inlined
    size_t pos = 0;
    const size_t N = data.size();
    bool at_line_start = true;
    auto next = [&](Tok &t) -> bool {
        for (;;) {
            while (pos < N && is_ws(data[pos])) { // (\v and \f too: a value ends at them, so they must be skipped here)
                at_line_start = data[pos] == '\n';
                pos++;
            }
            if (pos >= N) return false;
            if (data[pos] == '#') { // comment to the end of the line
                while (pos < N && data[pos] != '\n') pos++;
                continue;
            }
            break;
        }
        t.quoted = false;
        t.text = std::string_view();
        if (data[pos] == ';' && (at_line_start || pos == 0)) { // text field: up to a line that starts with ';'
            size_t b = pos + 1, e = data.find("\n;", b);
            if (e == std::string::npos) e = N;
            t.text = dv.substr(b, e - b);
            t.quoted = true;
            pos = std::min(N, e + 2);
            at_line_start = false;
            return true;
        }
        at_line_start = false;
        if (data[pos] == '\'' || data[pos] == '"') { // closes at the same quote followed by white space
            const char q = data[pos++];
            const size_t b = pos;
            while (pos < N && !(data[pos] == q && (pos + 1 >= N || is_ws(data[pos + 1]))) && data[pos] != '\n') pos++;
            t.text = dv.substr(b, pos - b);
            t.quoted = true;
            if (pos < N && data[pos] == q) pos++;
            return true;
        }
        const size_t b = pos;
        while (pos < N && !is_ws(data[pos])) pos++;
        t.text = dv.substr(b, pos - b);
        return true;
    };
    auto lower = [](std::string_view v) {
        std::string s(v);
        for (char &c : s) c = (char)tolower((unsigned char)c);
        return s;
    };
    auto is_keyword = [&](const Tok &t) {
        if (t.quoted || t.text.empty()) return false;
        const char c0 = (char)tolower((unsigned char)t.text[0]); // loop_ / data_ / save_ / stop_ / global_
        if (c0 != 'l' && c0 != 'd' && c0 != 's' && c0 != 'g') return false;
        const std::string l = lower(t.text);
        return l == "loop_" || l.rfind("data_", 0) == 0 || l.rfind("save_", 0) == 0 || l == "stop_" || l == "global_";
    };
    Tok t;
    bool have = next(t);
    while (have) {
        if (t.quoted || t.text.size() != 5 || lower(t.text) != "loop_") {
            have = next(t);
            continue;
        }
        std::vector<std::string> header;
        while ((have = next(t)) && !t.quoted && !t.text.empty() && t.text[0] == '_') header.push_back(std::string(t.text.substr(1)));
        if (std::find(header.begin(), header.end(), "atom_site.group_PDB") == header.end()) continue; // another loop
        auto col = [&](const char *name) -> int {
            auto it = std::find(header.begin(), header.end(), name);
            return it == header.end() ? -1 : (int)(it - header.begin());
        };
        const int c_asym = col("atom_site.label_asym_id"), c_auth_asym = col("atom_site.auth_asym_id"),
                  c_b = col("atom_site.B_iso_or_equiv"), c_comp = col("atom_site.label_comp_id"), c_id = col("atom_site.id"),
                  c_model = col("atom_site.pdbx_PDB_model_num"), c_name = col("atom_site.label_atom_id"),
                  c_seq = col("atom_site.label_seq_id"), c_auth_seq = col("atom_site.auth_seq_id"),
                  c_type = col("atom_site.type_symbol"), c_x = col("atom_site.Cartn_x"), c_y = col("atom_site.Cartn_y"),
                  c_z = col("atom_site.Cartn_z");
        if (c_asym < 0 || c_comp < 0 || c_id < 0 || c_name < 0 || c_seq < 0 || c_type < 0 || c_x < 0 || c_y < 0 || c_z < 0) {
            err = "Missing column in coordinate atoms data loop"; // cif.rs:190-201: the loop is not parsed
            return true;
        }
        const size_t W = header.size();
        std::vector<Tok> row(W);
        // a value that is `.` / `?` (unquoted) is inapplicable / unknown: None
        auto missing = [](const Tok &v) { return !v.quoted && (v.text == "." || v.text == "?"); };
        auto num = [&](const Tok &v, float *out) { // Value::Numeric is an f32 (integers travel through it too)
            if (v.quoted || missing(v)) return false;
            return field_f32(v.text.data(), v.text.size(), out);
        };
        bool first = true;
        long long first_model = 0;
        while (have && !is_keyword(t) && !(!t.quoted && !t.text.empty() && t.text[0] == '_')) {
            row[0] = t;
            size_t k = 1;
            for (; k < W && (have = next(t)); k++) row[k] = t;
            if (k < W) break; // truncated last row
            have = next(t);
            float fv;
            long long model = 1;
            if (c_model >= 0 && num(row[c_model], &fv)) model = (long long)fv;
            if (first) first_model = model;
            else if (model != first_model) break;
            first = false;
            const std::string_view an = row[c_name].text;
            if (missing(row[c_name]) || an.empty() || an.size() > 4) {
                err = "Invalid atom name in the atom_site loop";
                return false;
            }
            uint8_t name4[4] = {' ', ' ', ' ', ' '};
            if (an.size() == 4) memcpy(name4, an.data(), 4);
            else memcpy(name4 + 1, an.data(), an.size());
            uint8_t res3[3] = {' ', ' ', ' '};
            const std::string_view rn = row[c_comp].text;
            if (missing(row[c_comp])) {
                err = "Residue name should be provided";
                return false;
            }
            if (rn.size() <= 3) memcpy(res3, rn.data(), rn.size()); // longer names become blank (cif.rs:319-321)
            float seq;
            if (!(c_auth_seq >= 0 && num(row[c_auth_seq], &seq) && truncf(seq) == seq) &&
                !(num(row[c_seq], &seq) && truncf(seq) == seq)) {
                err = "Residue number should be provided";
                return false;
            }
            uint8_t chain = 0;
            if (c_auth_asym >= 0 && !missing(row[c_auth_asym]) && row[c_auth_asym].text.size() == 1) chain = (uint8_t)row[c_auth_asym].text[0];
            else if (!missing(row[c_asym]) && row[c_asym].text.size() == 1) chain = (uint8_t)row[c_asym].text[0];
            else {
                err = "Chain name should be provided (one character)";
                return false;
            }
            float x, y, z, b = 1.0f;
            if (!num(row[c_x], &x) || !num(row[c_y], &y) || !num(row[c_z], &z)) {
                err = "Atom position should be provided";
                return false;
            }
            if (c_b >= 0) {
                float bv;
                if (num(row[c_b], &bv)) b = bv;
            }
            a.x.push_back(x);
            a.y.push_back(y);
            a.z.push_back(z);
            a.b.push_back(b);
            a.name.insert(a.name.end(), name4, name4 + 4);
            a.rname.insert(a.rname.end(), res3, res3 + 3);
            a.chain.push_back(chain);
            a.serial.push_back((uint64_t)(int64_t)seq); // `isize as u64`
        }
        // (a second atom_site loop would be parsed as well, like the reference's walk over every loop)
    }
    return true;
}

bool has_suffix(const std::string &s, const char *suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}
// ---- Foldcomp input (src/structure/io/fcz.rs) ----------------------------------------------------------------------
// The Foldcomp codec is a third-party C++ library that the reference vendors (lib/foldcomp) and binds through a four-
// function C ABI (lib/foldcomp/foldcompffi.h:18-21, called at fcz.rs:80-94).  It is bound here the same way, at run
// time: dlopen of $FD_FOLDCOMP_LIB, else "libfoldcomp_ffi.so" on the loader path.  Everything around the codec -- the
// database files, the entry lookup, atoms -> Structure -> CompactStructure -- is this library's own code.
struct FczAtom { // atom_t (foldcompffi.h:8-16) = Atom (src/structure/atom.rs:3-15, repr(C))
    float x, y, z;
    char atom[4];
    uint64_t atom_idx;
    char chain;
    char aa[3];
    uint64_t res_idx;
    float bfactor;
};
static_assert(sizeof(FczAtom) == 48, "atom_t layout");
struct FczCodec {
    void *(*create)() = nullptr;
    FczAtom *(*process)(void *, const unsigned char *, size_t, size_t *) = nullptr;
    void (*release)(FczAtom *) = nullptr;
    void (*destroy)(void *) = nullptr;
    std::string err;
};
const FczCodec &fcz_codec() {
    static const FczCodec codec = [] {
        FczCodec c;
        const char *env = getenv("FD_FOLDCOMP_LIB");
        const char *name = env && *env ? env : "libfoldcomp_ffi.so";
        void *h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (!h) {
            c.err = std::string("Foldcomp input needs the Foldcomp codec library (foldcomp_create / foldcomp_process / "
                                "foldcomp_free / foldcomp_destroy); set FD_FOLDCOMP_LIB to libfoldcomp_ffi.so: ") + dlerror();
            return c;
        }
        c.create = (void *(*)())dlsym(h, "foldcomp_create");
        c.process = (FczAtom * (*)(void *, const unsigned char *, size_t, size_t *)) dlsym(h, "foldcomp_process");
        c.release = (void (*)(FczAtom *))dlsym(h, "foldcomp_free");
        c.destroy = (void (*)(void *))dlsym(h, "foldcomp_destroy");
        if (!c.create || !c.process || !c.release || !c.destroy) {
            c.create = nullptr;
            c.err = std::string(name) + " does not export the Foldcomp C ABI (foldcompffi.h)";
        }
        return c;
    }();
    return codec;
}
// one compressed entry -> atoms (fcz.rs:80-90: create, process, Structure::update per atom, destroy, free)
bool fcz_decode(const uint8_t *p, size_t n, Atoms &a, std::string &err) {
    if (n < 4 || memcmp(p, "FCMP", 4) != 0) { // the codec's own tag check (foldcomp.cpp:908-915) -- its C wrapper drops the verdict
        err = "not a Foldcomp entry (no FCMP tag)";
        return false;
    }
    const FczCodec &c = fcz_codec();
    if (!c.create) {
        err = c.err;
        return false;
    }
    void *inst = nullptr;
    size_t count = 0;
    FczAtom *out = nullptr;
    try { // the codec is C++ behind a C wrapper that catches nothing: a corrupt entry makes it throw (length_error, bad_alloc)
        inst = c.create();
        out = inst ? c.process(inst, p, n, &count) : nullptr;
    } catch (const std::exception &e) {
        err = std::string("the Foldcomp codec failed on this entry: ") + e.what();
        return false; // (the instance is abandoned: its state is unknown)
    } catch (...) {
        err = "the Foldcomp codec failed on this entry";
        return false;
    }
    if (!out) {
        if (inst) c.destroy(inst);
        err = "the Foldcomp codec returned no atoms";
        return false;
    }
    a.x.reserve(count);
    for (size_t i = 0; i < count; i++) {
        const FczAtom &t = out[i];
        a.x.push_back(t.x);
        a.y.push_back(t.y);
        a.z.push_back(t.z);
        a.b.push_back(t.bfactor);
        a.name.insert(a.name.end(), t.atom, t.atom + 4);
        a.rname.insert(a.rname.end(), t.aa, t.aa + 3);
        a.chain.push_back((uint8_t)t.chain);
        a.serial.push_back(t.res_idx);
    }
    c.destroy(inst);
    c.release(out);
    return true;
}
bool read_whole_file(const std::string &path, std::string &out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}

// read_structure_from_path (src/controller/io.rs:337-379): .pdb / .ent / .cif, each optionally .gz
bool read_structure_atoms(const char *path, Atoms &a, std::string &err) {
    const std::string p = path;
    const bool gz = has_suffix(p, ".gz");
    const std::string base = gz ? p.substr(0, p.size() - 3) : p;
    const bool cif = has_suffix(base, ".cif");
    if (!cif && !has_suffix(base, ".pdb") && !has_suffix(base, ".ent") && !has_suffix(base, ".PDB")) {
        err = std::string("unsupported structure file extension (expected .pdb, .ent or .cif, optionally .gz): ") + path;
        return false;
    }
    std::string data;
    if (!read_file_text(path, gz, data)) {
        err = std::string(cif ? "Failed to read CIF file: " : "Failed to read PDB file: ") + path;
        return false;
    }
    if (cif) {
        if (!parse_cif_atoms(data, a, err)) {
            err = std::string("Failed to read structure from CIF file: ") + path + ": " + err;
            return false;
        }
        return true;
    }
    return parse_pdb_atoms(data, a);
}

std::string rust_f32(float v) { // Rust `{}`: shortest round-trip, never scientific
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[512];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

} // namespace

// ---------------------------------------------------------------------------------------------
struct fdh_store {
    std::vector<uint64_t> row_offsets{0};
    std::vector<float> n, ca, cb;
    std::vector<uint8_t> aa, cb_valid;
    std::vector<std::string> names;
    std::vector<float> plddt;
    // per residue labels for result formatting
    std::vector<uint8_t> chain;
    std::vector<uint64_t> serial;
};

struct fdh_index {
    // owned buffers (build) or mmaps (load)
    std::vector<uint32_t> own_hashes;
    std::vector<uint64_t> own_offsets;
    std::vector<uint8_t> own_values;
    void *map_off = nullptr, *map_val = nullptr;
    size_t map_off_len = 0, map_val_len = 0;
    const uint32_t *hashes = nullptr;
    const uint64_t *offsets = nullptr; // may be unaligned when mapped and count is odd -> copied instead
    const uint8_t *values = nullptr;
    uint64_t count = 0, value_bytes = 0;
    std::vector<std::string> names;
    std::vector<uint32_t> nres;
    std::vector<float> plddt;
    std::vector<uint64_t> db_key; // 5th lookup column (lookup.rs:17-58); empty = the structure id
    bool fcz_input = false;       // built from a Foldcomp database: PREFIX.type says input_format = "FCZDB"
    std::string foldcomp_db;      // PREFIX.type foldcomp_db (the `index -p` argument, build_index.rs:228-232); "" = none
    fd_hash_params params{0, 0, 20.0f, 0, 0, {0}};
    ~fdh_index() {
        if (map_off) munmap(map_off, map_off_len);
        if (map_val) munmap(map_val, map_val_len);
    }
};

namespace {

struct QEntry {
    uint32_t hash;
    uint32_t qi, qj; // structure indices of the query residues
    uint8_t primary;
    uint32_t pair; // which ordered query pair produced it (its idf is the idf of that pair's observed hash)
    float idf;
};
struct AAD {
    uint8_t aa1, aa2;
    float dist;
    uint32_t qi;
};
struct Query {
    std::shared_ptr<const fdh_compact> st; // shared between the queries of a batch that use the same structure
    std::string qstring;
    std::vector<uint32_t> indices;
    // residue_count of query_pdb.rs:355-359: the PARSED query residues (the structure's residues for an empty query),
    // including those make_query_map cannot resolve in the structure; the denominator of the node-ratio filters
    uint32_t residue_count = 0;
    std::vector<QEntry> entries;
    std::vector<uint32_t> pair_hash; // observed hash per pair
    std::vector<AAD> aad;
    // derived
    std::vector<uint16_t> edge_of_hash, edge_node;
    uint32_t n_nodes = 0;
    std::vector<uint32_t> hashes_sorted;
    uint32_t max_q = 0;
    std::vector<uint32_t> hashes_flat;
    std::vector<uint8_t> aad_aa1, aad_aa2;
    std::vector<float> aad_dist;
    std::vector<uint32_t> aad_qi;
    // hash-range shards (fdh_queries_set_shards): one vote bit per (query edge, rank that owns some of its hashes)
    std::vector<uint16_t> s_edge_of_hash, s_edge_node, s_edge_group;
    // fd_verify_query arrays aligned with hashes_sorted (vs_idf is filled when the batch is finalised)
    std::vector<uint32_t> vs_qi, vs_qj;
    std::vector<uint8_t> vs_sym;
    std::vector<float> vs_idf;
    std::vector<uint32_t> vs_entry; // position in entries
    // position of hash h in hashes_sorted (and the vs_* arrays); h must be a hash of this query
    size_t sorted_pos(uint32_t h) const {
        return (size_t)(std::lower_bound(hashes_sorted.begin(), hashes_sorted.end(), h) - hashes_sorted.begin());
    }
};

// pdb_tr.rs:95-162 with default bins: res1 == res2 and the two torsions decode to the same angle.  The decoded angles
// depend only on the four 2-bit sin / cos bins of the torsions (the low 8 bits of the hash), so the verdict is tabulated
// once (256 entries; the same arithmetic as before, evaluated per bin pattern instead of per hash).
bool torsion_bins_symmetric(uint32_t low8) {
    auto cont = [](uint32_t v, float mn, float mx, float nb) { return (float)v * ((mx - mn) / (nb - 1.0f)) + mn; };
    const float deg = 180.0f / 3.14159274101257324f;
    float s1 = cont((low8 >> 6) & 3u, -1.f, 1.f, 4.f), c1 = cont((low8 >> 4) & 3u, -1.f, 1.f, 4.f);
    float s2 = cont((low8 >> 2) & 3u, -1.f, 1.f, 4.f), c2 = cont(low8 & 3u, -1.f, 1.f, 4.f);
    return fdm::atan2f_exact(s1, c1) * deg == fdm::atan2f_exact(s2, c2) * deg;
}
bool hash_is_symmetric(uint32_t h) {
    static const std::array<uint8_t, 256> table = [] {
        std::array<uint8_t, 256> t{};
        for (uint32_t k = 0; k < 256; k++) t[k] = torsion_bins_symmetric(k) ? 1 : 0;
        return t;
    }();
    if (((h >> 25) & 31u) != ((h >> 20) & 31u)) return false;
    return table[h & 0xffu] != 0;
}

bool host_pair_feature(const fdh_compact &c, size_t i, size_t j, float cutoff, float *f) {
    if (i == j) return false;
    const uint8_t a1 = c.aa[i], a2 = c.aa[j];
    if (a1 == 255 || a2 == 255 || !c.cb_valid[i] || !c.cb_valid[j]) return false;
    const float d = fdg::dist(c.CA(i), c.CA(j));
    if (d > cutoff) return false;
    fdg::pair_feature(c.N(i), c.CA(i), c.CB(i), c.N(j), c.CA(j), c.CB(j), (float)(a1 & 0x7F), (float)(a2 & 0x7F), d, f);
    return true;
}

void one_letter(char ch, std::vector<uint8_t> &o) { // convert.rs:223-262
    static const char *std20 = "ARNDCQEGHILKMFPSTWYV";
    const char *p = ch ? strchr(std20, ch) : nullptr;
    if (p) {
        o.push_back((uint8_t)(p - std20));
        return;
    }
    auto put = [&](std::initializer_list<int> l) {
        for (int v : l) o.push_back((uint8_t)v);
    };
    switch (ch) {
        case 'B': put({2, 3}); break;
        case 'Z': put({5, 6}); break;
        case 'X':
        case 'x':
            for (int i = 0; i < 20; i++) o.push_back((uint8_t)i);
            break;
        case 'J': put({9, 10}); break;
        case 'U': put({4}); break;
        case 'O': put({11}); break;
        case 'p': put({1, 8, 11}); break;
        case 'n': put({3, 6}); break;
        case 'h': put({2, 5, 15, 16, 18}); break;
        case 'b': put({0, 4, 7, 9, 10, 12, 13, 14, 19}); break;
        case 'a': put({8, 13, 17, 18}); break;
        default: o.push_back(255);
    }
}

struct ParsedQuery {
    std::vector<uint8_t> chains;
    std::vector<uint64_t> serials;
    std::vector<int> has_sub;
    std::vector<std::vector<uint8_t>> subs;
};
// parse_query_string (src/controller/query.rs:331-384)
bool parse_query(const char *q, uint8_t default_chain, ParsedQuery &out) {
    std::string s;
    for (const char *p = q; *p; p++)
        if (*p != ' ') s.push_back(*p);
    if (s.empty()) return true;
    if (!isalpha(default_chain)) default_chain = 'A';
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
        bool hs = false;
        std::vector<uint8_t> sub;
        std::string range = rest;
        size_t colon = rest.find(':');
        if (colon != std::string::npos) {
            hs = true;
            range = rest.substr(0, colon);
            for (char ch : rest.substr(colon + 1))
                if (isalpha((unsigned char)ch)) one_letter(ch, sub);
        }
        uint64_t a, b;
        size_t dash = range.find('-');
        if (dash != std::string::npos) {
            if (!field_u64(range.data(), dash, &a) || !field_u64(range.data() + dash + 1, range.size() - dash - 1, &b))
                return false;
        } else {
            if (!field_u64(range.data(), range.size(), &a)) return false;
            b = a;
        }
        for (uint64_t r = a; r <= b; r++) {
            out.chains.push_back(chain);
            out.serials.push_back(r);
            out.has_sub.push_back(hs);
            out.subs.push_back(sub);
        }
    }
    return true;
}

} // namespace

struct fdh_queries {
    fdh_query_params p;
    std::vector<float> dist_thr, angle_thr;
    std::vector<Query> q;
    bool finalized = false;
    // identity of the batch for fd_count_query_*_id: a process-wide serial number and a generation that every
    // modification (add, finalize, set_shards) bumps -- a search of an unchanged batch reuses its uploaded form
    uint64_t uid = next_uid();
    mutable uint64_t generation = 1;
    static uint64_t next_uid() {
        static std::atomic<uint64_t> n{1};
        return n.fetch_add(1);
    }
    uint64_t batch_id() const { return (uid << 24) | (generation & 0xffffffull); }
    std::vector<uint64_t> shard_bounds; // world + 1 ascending hash boundaries; empty = unsharded
    // query side of the verification, resident on the device of the context that searched first (fd_verify_prepare);
    // built by fdh_queries_finalize (or lazily by the first search), dropped whenever the idf values change
    mutable fd_verify_prepared *vprep = nullptr;
    mutable int vprep_device = -1;
    // id-range shards (fdh_queries_finalize_sharded): the WHOLE batch -- the queries of every rank, rank by rank -- as
    // the flat scan inputs of fd_count_query_sharded, with the global list length of every hash
    struct ShardedBatch {
        int world = 0, rank = 0;
        uint64_t first_id = 0, total_structs = 0;
        std::vector<uint32_t> slice_begin;                    // [world + 1]
        std::vector<uint32_t> q_nh, q_ne, q_nn, q_expected;   // per query of the whole batch
        std::vector<uint32_t> hashes, gcounts;                // per hash
        std::vector<uint16_t> edge_of_hash, edge_node;        // per hash / per edge
    } sh;
    ~fdh_queries() { fd_verify_prepared_free(vprep); }
};

namespace {

// open-addressing set of the hashes already in Q.entries (local to build_query_map: nothing to free per query later)
struct SeenHashes {
    std::vector<uint32_t> slot; // entry index + 1, 0 = empty
    uint32_t mask = 0, used = 0;
    // cap = a power of two, at least twice the number of hashes expected (no rehash while the query map is built)
    explicit SeenHashes(uint32_t cap = 256) { resize(cap); }
    void resize(uint32_t cap) {
        slot.assign(cap, 0);
        mask = cap - 1;
    }
    static uint32_t mix(uint32_t h) {
        h ^= h >> 16;
        h *= 0x7feb352du;
        h ^= h >> 15;
        return h;
    }
    // true if h was absent (and is now recorded as entry `idx`)
    bool insert(uint32_t h, uint32_t idx, const std::vector<QEntry> &entries) {
        if (2 * (used + 1) > slot.size()) {
            std::vector<uint32_t> old;
            old.swap(slot);
            resize((uint32_t)old.size() * 2);
            for (uint32_t v : old)
                if (v) {
                    uint32_t p = mix(entries[v - 1].hash) & mask;
                    while (slot[p]) p = (p + 1) & mask;
                    slot[p] = v;
                }
        }
        uint32_t p = mix(h) & mask;
        while (slot[p]) {
            if (entries[slot[p] - 1].hash == h) return false;
            p = (p + 1) & mask;
        }
        slot[p] = idx + 1;
        used++;
        return true;
    }
};

// fdg::perfect_hash with the (sin, cos) bins of the angle features memoised on the exact f32 value: the ~11
// variants of one residue pair (query.rs:179-206) perturb one feature at a time, so most of their binary64
// sincos evaluations repeat.  Bit-identical to fdg::perfect_hash (same functions on the same inputs).
struct AngleBinCache {
    struct E {
        uint32_t bits, bins; // f32 bits of the angle -> sin bin << 16 | cos bin
    };
    E e[3][12];
    int n[3] = {0, 0, 0};
    // fdg::discretize(val, mn, mx, nbin) = sat_u32((val - mn) * disc + 0.5) with disc = 1 / ((mx - mn) / (nbin - 1)):
    // disc depends on the hash parameters only, so its two divisions are done once per query (same f32 operations,
    // same results as calling fdg::discretize)
    float disc_dist = 0.f, disc_angle = 0.f;
    explicit AngleBinCache(const fdg::HashParams &hp) {
        disc_dist = FD_DIV(1.0f, FD_DIV(FD_SUB(20.0f, 2.0f), FD_SUB(hp.nbin_dist, 1.0f)));
        disc_angle = FD_DIV(1.0f, FD_DIV(FD_SUB(1.0f, -1.0f), FD_SUB(hp.nbin_angle, 1.0f)));
    }
    static uint32_t disc(float val, float mn, float d) { return fdg::sat_u32(FD_ADD(FD_MUL(FD_SUB(val, mn), d), 0.5f)); }
    void reset() { n[0] = n[1] = n[2] = 0; }
    uint32_t bins(int slot, float a) {
        uint32_t bits;
        memcpy(&bits, &a, 4);
        for (int k = 0; k < n[slot]; k++)
            if (e[slot][k].bits == bits) return e[slot][k].bins;
        // Only the bins are needed: the platform's sinf / cosf (faithful to well under 1e-6) decide them whenever the
        // value that is truncated to the bin stays 1e-4 away from an integer -- the same margin argument as
        // fdg::pair_hash_fast -- and the binary64 route is taken for the rare angle next to a bin boundary.
        uint32_t b;
        const float ts = FD_ADD(FD_MUL(FD_SUB(sinf(a), -1.0f), disc_angle), 0.5f);
        const float tc = FD_ADD(FD_MUL(FD_SUB(cosf(a), -1.0f), disc_angle), 0.5f);
        const float fs = ts - floorf(ts), fc = tc - floorf(tc);
        if (fs > 1.0e-4f && fs < 1.0f - 1.0e-4f && fc > 1.0e-4f && fc < 1.0f - 1.0e-4f) {
            b = fdg::sat_u32(ts) << 16 | fdg::sat_u32(tc);
        } else {
            float sn, cs;
            fdm::sincosf_exact(a, &sn, &cs);
            b = disc(sn, -1.0f, disc_angle) << 16 | disc(cs, -1.0f, disc_angle);
        }
        if (n[slot] < 12) e[slot][n[slot]++] = E{bits, b};
        return b;
    }
    uint32_t hash(const float *f, const fdg::HashParams &) {
        const uint32_t res1 = fdg::sat_u32(f[0]), res2 = fdg::sat_u32(f[1]);
        const uint32_t ca = disc(f[2], 2.0f, disc_dist);
        const uint32_t cb = disc(f[3], 2.0f, disc_dist);
        const uint32_t b0 = bins(0, f[4]), b1 = bins(1, f[5]), b2 = bins(2, f[6]);
        // the sin / cos bins are OR-ed in unmasked, exactly like perfect_hash (pdb_tr.rs:21-75: no field masking)
        return res1 << 25 | res2 << 20 | ca << 16 | cb << 12 | (b0 >> 16) << 10 | (b0 & 0xffffu) << 8 |
               (b1 >> 16) << 6 | (b1 & 0xffffu) << 4 | (b2 >> 16) << 2 | (b2 & 0xffffu);
    }
};

// Hashing of one query: the tuned PDBTrRosetta single-bin route (memoised angle bins), or any encoding /
// `--multiple-bins` list through fd_hashtypes.cuh
struct QueryHasher {
    bool typed;
    fdg::HashParams hp;
    fdg::TypedParams tp;
    uint32_t obs_nbd = 0, obs_nba = 0; // bins of the observed hash that carries a pair's idf (query.rs:283-287)
    AngleBinCache bins;
    explicit QueryHasher(const fd_hash_params &p)
        : typed(!fdg::ht_default_route(&p)), hp(fdg::make_params(p.nbin_dist, p.nbin_angle, p.dist_cutoff)), bins(hp) {
        fdg::typed_params_from(&p, &tp); // validated by fdh_queries_new
        fdg::ht_resolve_single(tp.type, p.nbin_dist, p.nbin_angle, &obs_nbd, &obs_nba);
        if (p.n_multiple_bins) // insert_binned_hash: a pair with a zero takes the defaults (query.rs:60-64)
            for (uint32_t k = 0; k < tp.n_bins; k++) fdg::ht_resolve_single(tp.type, tp.nbd[k], tp.nba[k], &tp.nbd[k], &tp.nba[k]);
    }
    uint32_t observed(const float *f) { return typed ? fdg::typed_hash(tp.type, f, obs_nbd, obs_nba) : bins.hash(f, hp); }
};

void qinsert(Query &Q, SeenHashes &seen, QueryHasher &H, const float *f, uint32_t qi, uint32_t qj, bool primary,
             uint32_t pair) { // insert_binned_hash (query.rs:53-84): first writer wins
    const uint32_t nb = H.typed ? H.tp.n_bins : 1u;
    for (uint32_t b = 0; b < nb; b++) {
        const uint32_t h = H.typed ? fdg::typed_hash(H.tp.type, f, H.tp.nbd[b], H.tp.nba[b]) : H.bins.hash(f, H.hp);
        if (!seen.insert(h, (uint32_t)Q.entries.size(), Q.entries)) continue;
        Q.entries.push_back(QEntry{h, qi, qj, (uint8_t)primary, pair, 0.f});
    }
}

// get_single_feature (feature.rs:11-190) of the selected encoding on the host; *ca_dist = the CA-CA distance
bool host_typed_feature(const fdh_compact &c, size_t i, size_t j, const fdg::TypedParams &tp, float *f, float *ca_dist) {
    if (i == j) return false;
    const uint8_t a1 = c.aa[i], a2 = c.aa[j];
    if (a1 == 255 || a2 == 255) return false;
    if (fdg::ht_needs_cb(tp.type) && (!c.cb_valid[i] || !c.cb_valid[j])) return false;
    fdg::Nbr nb;
    if (fdg::ht_needs_neighbours(tp.type)) { // feature.rs:113, 163: not for the first / last residue
        if (i == 0 || j == 0 || i + 1 >= c.nres() || j + 1 >= c.nres()) return false;
        nb = fdg::Nbr{c.CA(i - 1), c.CA(i + 1), c.CA(j - 1), c.CA(j + 1), (float)j - (float)i};
    }
    const float d = fdg::typed_screen_dist(tp.type, c.CA(i), c.CB(i), c.CA(j), c.CB(j));
    if (d > tp.dist_cutoff) return false;
    fdg::typed_feature(tp.type, c.N(i), c.CA(i), c.CB(i), c.N(j), c.CA(j), c.CB(j), (float)(a1 & 0x7F), (float)(a2 & 0x7F),
                       d, f, &nb);
    *ca_dist = fdg::dist(c.CA(i), c.CA(j));
    return true;
}

// make_query_map (src/controller/query.rs:208-329) minus the idf values, which need the index
bool build_query_map(Query &Q, const ParsedQuery &pq, const fdh_queries &qs) {
    const fdh_compact &c = *Q.st;
    const fdg::HashParams hp = fdg::make_params(qs.p.hash.nbin_dist, qs.p.hash.nbin_angle, qs.p.hash.dist_cutoff);
    std::unordered_map<uint32_t, std::vector<uint8_t>> submap;
    for (size_t i = 0; i < pq.chains.size(); i++) {
        int64_t idx = -1;
        if (qs.p.serial_query) {
            idx = (int64_t)pq.serials[i];
        } else {
            for (size_t r = 0; r < c.nres(); r++) // CompactStructure::get_index (core.rs:215-223)
                if (c.chain[r] == pq.chains[i] && c.serial[r] == pq.serials[i]) {
                    idx = (int64_t)r;
                    break;
                }
        }
        if (idx < 0) continue;
        Q.indices.push_back((uint32_t)idx);
        if (pq.has_sub[i]) submap[(uint32_t)idx] = pq.subs[i];
    }
    const float rad = 3.14159274101257324f / 180.0f; // f32::to_radians
    float f[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, fn[9], ff[9];
    QueryHasher QH(qs.p.hash);
    AngleBinCache &bins = QH.bins;
    int dist_idx[2], angle_idx[7];
    const int n_dist_idx = fdg::typed_dist_index(QH.tp.type, dist_idx), n_angle_idx = fdg::typed_angle_index(QH.tp.type, angle_idx);
    const size_t K = Q.indices.size();
    uint32_t seen_cap = 256;
    {
        const size_t n_pairs = K > 1 ? K * (K - 1) : 0;
        const size_t per_pair = (1 + 4 * qs.dist_thr.size() + 10 * qs.angle_thr.size()) * (QH.typed ? QH.tp.n_bins : 1);
        Q.entries.reserve(std::min<size_t>(n_pairs * per_pair, 1u << 16));
        Q.pair_hash.reserve(n_pairs);
        Q.aad.reserve(n_pairs);
        while (seen_cap < 2 * n_pairs * per_pair && seen_cap < (1u << 16)) seen_cap *= 2;
    }
    SeenHashes seen(seen_cap);
    for (size_t a = 0; a < K; a++)
        for (size_t b = 0; b < K; b++) {
            if (a == b) continue;
            const uint32_t I = Q.indices[a], J = Q.indices[b];
            if (I >= c.nres() || J >= c.nres()) continue;
            float ca_dist;
            if (QH.typed) {
                if (!host_typed_feature(c, I, J, QH.tp, f, &ca_dist)) continue;
            } else {
                if (!host_pair_feature(c, I, J, hp.dist_cutoff, f)) continue;
                ca_dist = f[2];
            }
            memcpy(fn, f, sizeof(f));
            memcpy(ff, f, sizeof(f));
            // get_list_amino_acids_and_distances (core.rs:462-477): always the CA distance
            if (ca_dist <= 20.0f) Q.aad.push_back(AAD{(uint8_t)(c.aa[I] & 0x7F), (uint8_t)(c.aa[J] & 0x7F), ca_dist, I});
            const uint32_t pair = (uint32_t)Q.pair_hash.size();
            bins.reset();
            Q.pair_hash.push_back(QH.observed(f));
            qinsert(Q, seen, QH, f, I, J, true, pair);
            if (fdg::ht_has_aa_index(QH.tp.type)) { // apply_substitutions (query.rs:86-156, :301-306)
                const float o1 = fn[0], o2 = fn[1];
                auto si = submap.find(I), sj = submap.find(J);
                if (si != submap.end()) {
                    for (uint8_t s : si->second) {
                        float t[9];
                        memcpy(t, fn, sizeof(t));
                        t[0] = (float)s;
                        qinsert(Q, seen, QH, t, I, J, false, pair);
                    }
                    if (sj != submap.end())
                        for (uint8_t s : si->second)
                            for (uint8_t s2 : sj->second) {
                                fn[0] = (float)s;
                                fn[1] = (float)s2;
                                qinsert(Q, seen, QH, fn, I, J, false, pair);
                                fn[0] = o1;
                                fn[1] = o2;
                            }
                } else if (sj != submap.end()) {
                    for (uint8_t s : sj->second) {
                        float t[9];
                        memcpy(t, fn, sizeof(t));
                        t[1] = (float)s;
                        qinsert(Q, seen, QH, t, I, J, false, pair);
                    }
                }
            }
            auto expand = [&](const int *idxs, int n_idx, const std::vector<float> &thr, bool to_rad) {
                for (float t : thr) { // expand_and_insert (query.rs:179-206): mutate, hash, restore
                    const float delta = to_rad ? t * rad : t;
                    for (int u = 0; u < n_idx; u++) {
                        const int k = idxs[u];
                        fn[k] -= delta;
                        ff[k] += delta;
                        qinsert(Q, seen, QH, fn, I, J, false, pair);
                        qinsert(Q, seen, QH, ff, I, J, false, pair);
                        fn[k] += delta;
                        ff[k] -= delta;
                    }
                }
            };
            expand(dist_idx, n_dist_idx, qs.dist_thr, false);    // HashType::dist_index (feature.rs:269-277)
            expand(angle_idx, n_angle_idx, qs.angle_thr, true);  // HashType::angle_index (feature.rs:279-289)
        }
    // derived kernel inputs: edges and nodes numbered by first appearance in the entries.  The entries of one residue
    // pair are consecutive and, as long as no residue is listed twice in the query, distinct pairs are distinct
    // (qi, qj) edges: a new edge starts where the pair number changes, no map needed.
    bool repeated_index = false;
    for (size_t a = 0; a < K && !repeated_index; a++)
        for (size_t b = a + 1; b < K; b++)
            if (Q.indices[a] == Q.indices[b]) repeated_index = true;
    const size_t H0 = Q.entries.size();
    Q.edge_of_hash.reserve(H0);
    Q.hashes_flat.reserve(H0);
    std::vector<uint32_t> node_of; // residue index of every node, in first-appearance order
    auto node_id = [&](uint32_t qi) -> uint16_t {
        for (size_t k = 0; k < node_of.size(); k++)
            if (node_of[k] == qi) return (uint16_t)k;
        node_of.push_back(qi);
        return (uint16_t)(node_of.size() - 1);
    };
    if (!repeated_index) {
        uint32_t last_pair = 0xffffffffu;
        uint32_t n_edges = 0;
        for (auto &e : Q.entries) {
            if (e.pair != last_pair) {
                if (n_edges >= 65535) return false;
                last_pair = e.pair;
                n_edges++;
                Q.edge_node.push_back(node_id(e.qi));
            }
            Q.edge_of_hash.push_back((uint16_t)(n_edges - 1));
            Q.hashes_flat.push_back(e.hash);
            Q.max_q = std::max(Q.max_q, std::max(e.qi, e.qj));
        }
    } else {
        std::unordered_map<uint64_t, uint16_t> edges;
        for (auto &e : Q.entries) {
            const uint64_t key = ((uint64_t)e.qi << 32) | e.qj;
            auto it = edges.find(key);
            if (it == edges.end()) {
                if (edges.size() >= 65535) return false;
                it = edges.emplace(key, (uint16_t)edges.size()).first;
                Q.edge_node.push_back(node_id(e.qi));
            }
            Q.edge_of_hash.push_back(it->second);
            Q.hashes_flat.push_back(e.hash);
            Q.max_q = std::max(Q.max_q, std::max(e.qi, e.qj));
        }
    }
    Q.n_nodes = (uint32_t)node_of.size();
    const size_t H = Q.entries.size();
    Q.vs_entry.resize(H);
    { // entries ordered by hash (the hashes of a query are distinct): sort (hash, position) keys, not positions
        std::vector<uint64_t> keys(H);
        for (size_t k = 0; k < H; k++) keys[k] = (uint64_t)Q.entries[k].hash << 32 | (uint64_t)k;
        std::sort(keys.begin(), keys.end());
        for (size_t k = 0; k < H; k++) Q.vs_entry[k] = (uint32_t)keys[k];
    }
    Q.hashes_sorted.resize(H);
    Q.vs_qi.resize(H);
    Q.vs_qj.resize(H);
    Q.vs_sym.resize(H);
    for (size_t k = 0; k < H; k++) {
        const QEntry &e = Q.entries[Q.vs_entry[k]];
        Q.hashes_sorted[k] = e.hash;
        Q.vs_qi[k] = e.qi;
        Q.vs_qj[k] = e.qj;
        Q.vs_sym[k] = QH.typed ? fdg::typed_is_symmetric(QH.tp.type, e.hash) : hash_is_symmetric(e.hash);
    }
    Q.vs_idf.assign(H, 0.f);
    for (auto &d : Q.aad) {
        Q.aad_aa1.push_back(d.aa1);
        Q.aad_aa2.push_back(d.aa2);
        Q.aad_dist.push_back(d.dist);
        Q.aad_qi.push_back(d.qi);
    }
    return true;
}

// ---- graph components (src/controller/graph.rs:29-50): Tarjan SCCs U undirected components, size >= 2,
// each sorted by node id, list sorted and deduped.  Graphs here have a handful of nodes.
struct Graph {
    std::vector<uint32_t> node_res;               // node id -> target residue
    std::vector<std::pair<uint32_t, uint32_t>> e; // directed edges (node ids), insertion order
};

void graph_components(const Graph &g, std::vector<std::vector<uint32_t>> &out) {
    const uint32_t n = (uint32_t)g.node_res.size();
    out.clear();
    std::vector<std::vector<uint32_t>> adj(n), und(n);
    for (auto &e : g.e) {
        adj[e.first].push_back(e.second);
        und[e.first].push_back(e.second);
        und[e.second].push_back(e.first);
    }
    std::vector<int> index(n, -1), low(n, 0);
    std::vector<uint8_t> on(n, 0);
    std::vector<uint32_t> stack;
    int counter = 0;
    struct Fr {
        uint32_t v, k;
    };
    std::vector<Fr> cs;
    for (uint32_t s = 0; s < n; s++) {
        if (index[s] >= 0) continue;
        cs.push_back({s, 0});
        index[s] = low[s] = counter++;
        stack.push_back(s);
        on[s] = 1;
        while (!cs.empty()) {
            Fr &f = cs.back();
            if (f.k < adj[f.v].size()) {
                const uint32_t w = adj[f.v][f.k++];
                if (index[w] < 0) {
                    index[w] = low[w] = counter++;
                    stack.push_back(w);
                    on[w] = 1;
                    cs.push_back({w, 0});
                } else if (on[w]) {
                    low[f.v] = std::min(low[f.v], index[w]);
                }
            } else {
                const uint32_t v = f.v;
                cs.pop_back();
                if (!cs.empty()) low[cs.back().v] = std::min(low[cs.back().v], low[v]);
                if (low[v] == index[v]) {
                    std::vector<uint32_t> comp;
                    uint32_t w;
                    do {
                        w = stack.back();
                        stack.pop_back();
                        on[w] = 0;
                        comp.push_back(w);
                    } while (w != v);
                    if (comp.size() >= 2) out.push_back(std::move(comp));
                }
            }
        }
    }
    std::vector<uint8_t> seen(n, 0);
    std::vector<uint32_t> st;
    for (uint32_t s = 0; s < n; s++) {
        if (seen[s]) continue;
        std::vector<uint32_t> comp;
        st.assign(1, s);
        seen[s] = 1;
        while (!st.empty()) {
            const uint32_t v = st.back();
            st.pop_back();
            comp.push_back(v);
            for (uint32_t w : und[v])
                if (!seen[w]) {
                    seen[w] = 1;
                    st.push_back(w);
                }
        }
        if (comp.size() >= 2) out.push_back(std::move(comp));
    }
    for (auto &c : out) std::sort(c.begin(), c.end());
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
}

struct MatchTmp {             // one connected component of one candidate
    uint32_t cand;
    float idf;
    std::vector<uint32_t> res_hash, res_final; // per query residue: target residue index + 1, 0 = none
    std::vector<uint32_t> aq, at;              // alignment pairs (query structure idx, target residue idx)
    uint32_t node_count_final = 0;
};

// map_query_and_retrieved_residues + the rescue loop (retrieve.rs:604-702, 453-516)
void match_component(const Query &Q, const std::vector<fd_cand_edge> &edges, const std::vector<uint32_t> &sub,
                     const fd_cand_pair *pairs, size_t n_pairs, uint32_t node_count, bool skip_ca_match,
                     MatchTmp &m) {
    uint32_t max_r = 0;
    for (uint32_t k : sub) max_r = std::max(max_r, std::max(edges[k].i, edges[k].j));
    const size_t q_size = (size_t)Q.max_q + 1, r_size = (size_t)max_r + 1;
    std::vector<uint8_t> counts(q_size * r_size, 0);
    std::vector<std::pair<uint8_t, uint32_t>> best(q_size, {0, 0});
    float idf = 0.f;
    for (uint32_t k : sub) {
        const fd_cand_edge &e = edges[k];
        const size_t sp = Q.sorted_pos(e.hash);
        const QEntry &qe = Q.entries[Q.vs_entry[sp]];
        idf += qe.idf; // calculate_subgraph_idf (retrieve.rs:705-719), edge order
        std::pair<uint32_t, uint32_t> pr[2];
        if (Q.vs_sym[sp]) {
            const uint32_t q1 = std::min(qe.qi, qe.qj), q2 = std::max(qe.qi, qe.qj);
            const uint32_t r1 = std::min(e.i, e.j), r2 = std::max(e.i, e.j);
            pr[0] = {q1, r1};
            pr[1] = {q2, r2};
        } else {
            pr[0] = {qe.qi, e.i};
            pr[1] = {qe.qj, e.j};
        }
        for (auto &qr : pr) {
            uint8_t &c = counts[qr.first * r_size + qr.second];
            if (c != 255) c++;
            auto &b = best[qr.first];
            if (c > b.first || (c == b.first && qr.second < b.second)) b = {c, qr.second};
        }
    }
    m.idf = idf;
    // bucket by count (descending), queries ascending inside a bucket; greedy one-to-one assignment
    std::vector<uint32_t> order;
    for (uint32_t q = 0; q < q_size; q++)
        if (best[q].first > 0) order.push_back(q);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return best[a].first > best[b].first; });
    std::vector<uint8_t> q_used(q_size, 0), r_used(r_size, 0);
    std::vector<uint32_t> qidx, ridx;
    for (uint32_t q : order) {
        const uint32_t r = best[q].second;
        if (!q_used[q] && !r_used[r]) {
            qidx.push_back(q);
            ridx.push_back(r);
            q_used[q] = r_used[r] = 1;
            if (qidx.size() == node_count) break;
        }
    }
    // rescue loop over all query residues
    const size_t NQ = Q.indices.size();
    m.res_hash.assign(NQ, 0);
    m.res_final.clear();
    std::vector<uint32_t> q_scan, r_scan;
    auto in_ridx = [&](uint32_t r) { return std::find(ridx.begin(), ridx.end(), r) != ridx.end(); };
    for (size_t qi_pos = 0; qi_pos < NQ; qi_pos++) {
        const uint32_t qi = Q.indices[qi_pos];
        int64_t mapped = -1;
        for (size_t k = 0; k < qidx.size(); k++)
            if (qidx[k] == qi) mapped = ridx[k]; // last wins, like collecting into a HashMap
        if (mapped >= 0) {
            const uint32_t ri = (uint32_t)mapped;
            m.res_hash[qi_pos] = ri + 1;
            auto pos = std::find(r_scan.begin(), r_scan.end(), ri);
            if (pos == r_scan.end()) {
                m.res_final.push_back(ri + 1);
                q_scan.push_back(qi);
                r_scan.push_back(ri);
            } else {
                const size_t pp = pos - r_scan.begin();
                m.res_final[pp] = 0; // sic (retrieve.rs:475): the scanned position indexes res_vec
                m.res_final.push_back(ri + 1);
                q_scan.erase(q_scan.begin() + pp);
                r_scan.erase(r_scan.begin() + pp);
                q_scan.push_back(qi);
                r_scan.push_back(ri);
            }
        } else {
            // candidate pairs (q_index == qi) whose second residue is already matched vote for their first
            std::vector<std::pair<uint32_t, uint32_t>> cm; // (target residue, count)
            for (size_t k = 0; k < n_pairs; k++) {
                if (pairs[k].q_index != qi || !in_ridx(pairs[k].j)) continue;
                bool found = false;
                for (auto &kv : cm)
                    if (kv.first == pairs[k].i) {
                        kv.second++;
                        found = true;
                        break;
                    }
                if (!found) cm.push_back({pairs[k].i, 1});
            }
            uint32_t mx = 0, nmax = 0, arg = 0;
            for (auto &kv : cm) mx = std::max(mx, kv.second);
            for (auto &kv : cm)
                if (kv.second == mx) {
                    nmax++;
                    arg = kv.first;
                }
            if (!cm.empty() && nmax == 1 && mx >= 2 &&
                std::find(r_scan.begin(), r_scan.end(), arg) == r_scan.end()) {
                m.res_final.push_back(arg + 1);
                q_scan.push_back(qi);
                r_scan.push_back(arg);
            } else {
                m.res_final.push_back(0);
            }
        }
    }
    const bool same = m.res_final == m.res_hash;
    if (skip_ca_match || same) {
        m.aq = qidx;
        m.at = ridx;
    } else {
        m.aq = q_scan;
        m.at = r_scan;
    }
    if (skip_ca_match) m.res_final = m.res_hash;
    m.node_count_final = 0;
    for (uint32_t v : m.res_final) m.node_count_final += v != 0;
}

} // namespace

// Large result blocks are recycled: a batch's result arrays are tens of megabytes, and a fresh malloc of that size is
// an mmap whose pages fault in (and are zeroed by the kernel) while the assembly threads write the rows, then an
// munmap at release -- every batch.  Freed blocks of >= 256 KB wait in a small cache (at most 16 blocks / 512 MB) and
// serve the next request they fit without wasting more than 3/4 of the block.
namespace {
struct BlockCache {
    struct Block {
        void *p;
        size_t cap;
    };
    std::mutex m;
    std::vector<Block> blocks;
    size_t bytes = 0;
    static constexpr size_t kMinBlock = 256u << 10, kMaxBlocks = 16, kMaxBytes = 512u << 20;

    void *get(size_t need, size_t *cap) {
        if (need >= kMinBlock) {
            std::lock_guard<std::mutex> lk(m);
            size_t best = blocks.size();
            for (size_t k = 0; k < blocks.size(); k++)
                if (blocks[k].cap >= need && blocks[k].cap / 4 <= need && (best == blocks.size() || blocks[k].cap < blocks[best].cap))
                    best = k;
            if (best != blocks.size()) {
                Block b = blocks[best];
                blocks.erase(blocks.begin() + (long)best);
                bytes -= b.cap;
                *cap = b.cap;
                return b.p;
            }
        }
        *cap = need;
        if (need >= kMinBlock) { // large blocks are page-locked: the device writes result rows straight into them
            void *p = nullptr;
            if (cudaHostAlloc(&p, need, cudaHostAllocDefault) == cudaSuccess) return p;
            cudaGetLastError();
            return nullptr;
        }
        return malloc(need);
    }
    void put(void *p, size_t cap) {
        if (!p) return;
        if (cap >= kMinBlock) {
            {
                std::lock_guard<std::mutex> lk(m);
                if (blocks.size() < kMaxBlocks && bytes + cap <= kMaxBytes) {
                    blocks.push_back(Block{p, cap});
                    bytes += cap;
                    return;
                }
            }
            cudaFreeHost(p);
            cudaGetLastError(); // (a result released after the CUDA context is gone)
            return;
        }
        free(p);
    }
};
BlockCache &block_cache() {
    static BlockCache *c = new BlockCache(); // leaked on purpose (results may be released during interpreter shutdown)
    return *c;
}
} // namespace

// result array without value-initialisation (the rows are written in place by the assembly threads)
template <typename T>
struct RawArray {
    T *p = nullptr;
    size_t n = 0, cap_bytes = 0;
    RawArray() {}
    RawArray(const RawArray &) = delete;
    RawArray &operator=(const RawArray &) = delete;
    ~RawArray() { block_cache().put(p, cap_bytes); }
    bool alloc(size_t count) {
        block_cache().put(p, cap_bytes);
        n = count;
        p = (T *)block_cache().get(std::max<size_t>(count, 1) * sizeof(T), &cap_bytes);
        return p != nullptr;
    }
    void shrink(size_t count) { n = std::min(n, count); } // rows in use (the block keeps its capacity)
    T *data() { return p; }
    const T *data() const { return p; }
    size_t size() const { return n; }
    T &operator[](size_t i) { return p[i]; }
};

struct fdh_results {
    std::vector<uint64_t> struct_off, match_off;
    RawArray<uint64_t> match_order;
    RawArray<fdh_struct_row> structs;
    RawArray<fdh_match_row> matches;
    RawArray<fdh_residue_match> residues;
    // fdh_search_params.want_metrics: per residue entry the target residue index + 1 (0 = none), per match row the
    // five similarity metrics (fd_metrics_store_batch)
    std::vector<uint32_t> res_index;
    std::vector<float> metrics;
    double host_ms = 0.0;
    uint64_t h2d_bytes = 0, d2h_bytes = 0; // bytes this search moved between host and device
    double wall_ms[4] = {0, 0, 0, 0};      // count_query call, verification call(s), row assembly, total
};

// =============================================================================================
extern "C" {

const char *fdh_last_error(void) { return g_err.c_str(); }

fdh_compact *fdh_compact_read_structure(const char *path) {
    Atoms a;
    std::string err;
    const std::string ps = path;
    if (ps.find(':') != std::string::npos) { // "DB:name" = an entry of a Foldcomp database (read_compact_structure, io.rs:303-334)
        const size_t c = ps.find(':');
        const size_t c2 = ps.find(':', c + 1);
        const std::string name = ps.substr(c + 1, c2 == std::string::npos ? std::string::npos : c2 - c - 1);
        fdh_fcz_db *db = fdh_fcz_db_open(ps.substr(0, c).c_str());
        if (!db) return nullptr;
        const int64_t k = fdh_fcz_db_find(db, name.c_str());
        fdh_compact *out = nullptr;
        if (k < 0) set_err("Entry with name " + name + " not found.");
        else out = fdh_fcz_db_read(db, k);
        fdh_fcz_db_close(db);
        return out;
    }
    if (has_suffix(ps, ".fcz")) { // one Foldcomp entry in a file of its own (the reference's test file data/foldcomp/7m0y.fcz)
        std::string bytes;
        if (!read_whole_file(ps, bytes)) {
            set_err("Failed to read Foldcomp file: " + ps);
            return nullptr;
        }
        return fdh_compact_from_fcz((const uint8_t *)bytes.data(), bytes.size());
    }
    if (!read_structure_atoms(path, a, err)) {
        set_err(err);
        return nullptr;
    }
    return adopt_compact(compact_from_atoms(a));
}
fdh_compact *fdh_compact_read_pdb(const char *path) { return fdh_compact_read_structure(path); }

// FoldcompDbReader (fcz.rs:21-136): PATH (concatenated entries), PATH.index (key \t start \t length), PATH.lookup
// (key \t name [\t ...]); both tables sorted by key (fcz.rs:47-52)
struct fdh_fcz_db {
    std::string path;
    void *map = nullptr;
    size_t map_len = 0;
    struct Entry {
        uint64_t key, start, len;
        std::string name;
    };
    std::vector<Entry> entries;       // the entries that have a name, ascending key: get_paths (fcz.rs:208-219)
    std::vector<uint32_t> by_name;    // entries[] positions sorted by name (sort_lookup_by_name + binary search, io.rs:325-326)
    ~fdh_fcz_db() {
        if (map && map_len) munmap(map, map_len);
    }
};
static bool fcz_parse_table(const std::string &path, int want_cols, std::vector<std::array<std::string, 3>> &rows) {
    std::string text;
    if (!read_whole_file(path, text)) return false;
    size_t pos = 0;
    while (pos < text.size()) {
        size_t nl = text.find('\n', pos);
        if (nl == std::string::npos) nl = text.size();
        std::string line = text.substr(pos, nl - pos);
        pos = nl + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        std::array<std::string, 3> r;
        size_t a = 0;
        int c = 0;
        for (; c < want_cols; c++) {
            size_t t = line.find('\t', a);
            r[c] = line.substr(a, t == std::string::npos ? std::string::npos : t - a);
            if (t == std::string::npos) {
                c++;
                break;
            }
            a = t + 1;
        }
        if (c < want_cols) return false;
        rows.push_back(std::move(r));
    }
    return true;
}
fdh_fcz_db *fdh_fcz_db_open(const char *path) {
    auto db = std::make_unique<fdh_fcz_db>();
    db->path = path;
    std::vector<std::array<std::string, 3>> idx, lk;
    if (!fcz_parse_table(db->path + ".lookup", 2, lk)) {
        set_err(std::string("Error reading foldcomp db lookup file: ") + path + ".lookup");
        return nullptr;
    }
    if (!fcz_parse_table(db->path + ".index", 3, idx)) {
        set_err(std::string("Error reading foldcomp db index file: ") + path + ".index");
        return nullptr;
    }
    int fd = open(path, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) {
        if (fd >= 0) close(fd);
        set_err(std::string("Error reading foldcomp db file: ") + path);
        return nullptr;
    }
    db->map_len = (size_t)st.st_size;
    if (db->map_len) {
        db->map = mmap(nullptr, db->map_len, PROT_READ, MAP_PRIVATE, fd, 0);
        if (db->map == MAP_FAILED) {
            db->map = nullptr;
            close(fd);
            set_err(std::string("cannot map ") + path);
            return nullptr;
        }
    }
    close(fd);
    std::vector<std::pair<uint64_t, std::string>> names;
    for (auto &r : lk) {
        uint64_t k;
        if (!field_u64(r[0].data(), r[0].size(), &k)) {
            set_err(std::string("malformed key in ") + path + ".lookup");
            return nullptr;
        }
        names.emplace_back(k, r[1]);
    }
    std::stable_sort(names.begin(), names.end(), [](auto &a, auto &b) { return a.first < b.first; });
    for (auto &r : idx) {
        fdh_fcz_db::Entry e;
        if (!field_u64(r[0].data(), r[0].size(), &e.key) || !field_u64(r[1].data(), r[1].size(), &e.start) ||
            !field_u64(r[2].data(), r[2].size(), &e.len)) {
            set_err(std::string("malformed row in ") + path + ".index");
            return nullptr;
        }
        if (e.start > db->map_len || e.len > db->map_len - e.start) {
            set_err(std::string("entry outside the database file in ") + path + ".index");
            return nullptr;
        }
        auto it = std::lower_bound(names.begin(), names.end(), e.key, [](auto &a, uint64_t k) { return a.first < k; });
        if (it == names.end() || it->first != e.key || it->second.empty()) continue; // no name: not a path (fcz.rs:211-217)
        e.name = it->second;
        db->entries.push_back(std::move(e));
    }
    std::stable_sort(db->entries.begin(), db->entries.end(), [](auto &a, auto &b) { return a.key < b.key; });
    db->by_name.resize(db->entries.size());
    for (size_t k = 0; k < db->entries.size(); k++) db->by_name[k] = (uint32_t)k;
    std::stable_sort(db->by_name.begin(), db->by_name.end(),
                     [&](uint32_t a, uint32_t b) { return db->entries[a].name < db->entries[b].name; });
    return db.release();
}
void fdh_fcz_db_close(fdh_fcz_db *db) { delete db; }
int64_t fdh_fcz_db_size(const fdh_fcz_db *db) { return (int64_t)db->entries.size(); }
const char *fdh_fcz_db_name(const fdh_fcz_db *db, int64_t k) {
    return k >= 0 && (size_t)k < db->entries.size() ? db->entries[(size_t)k].name.c_str() : nullptr;
}
uint64_t fdh_fcz_db_key(const fdh_fcz_db *db, int64_t k) {
    return k >= 0 && (size_t)k < db->entries.size() ? db->entries[(size_t)k].key : UINT64_MAX;
}
int64_t fdh_fcz_db_find(const fdh_fcz_db *db, const char *name) {
    const std::string n = name;
    auto it = std::lower_bound(db->by_name.begin(), db->by_name.end(), n,
                               [&](uint32_t a, const std::string &v) { return db->entries[a].name < v; });
    return it != db->by_name.end() && db->entries[*it].name == n ? (int64_t)*it : -1;
}
fdh_compact *fdh_fcz_db_read(const fdh_fcz_db *db, int64_t k) { // read_single_structure_by_id (fcz.rs:100-121)
    if (k < 0 || (size_t)k >= db->entries.size()) {
        set_err("Foldcomp entry position out of range");
        return nullptr;
    }
    const fdh_fcz_db::Entry &e = db->entries[(size_t)k];
    Atoms a;
    std::string err;
    if (!fcz_decode((const uint8_t *)db->map + e.start, (size_t)e.len, a, err)) {
        set_err("Foldcomp entry " + e.name + ": " + err);
        return nullptr;
    }
    return adopt_compact(compact_from_atoms(a));
}
fdh_compact *fdh_compact_from_fcz(const uint8_t *bytes, uint64_t n) {
    Atoms a;
    std::string err;
    if (!fcz_decode(bytes, (size_t)n, a, err)) {
        set_err(err);
        return nullptr;
    }
    return adopt_compact(compact_from_atoms(a));
}
fdh_compact *fdh_compact_from_atoms(int64_t n, const float *x, const float *y, const float *z, const uint8_t *an,
                                    const uint8_t *ch, const uint8_t *rn, const uint64_t *rs, const float *bf) {
    Atoms a;
    a.x.assign(x, x + n);
    a.y.assign(y, y + n);
    a.z.assign(z, z + n);
    a.b.assign(bf, bf + n);
    a.name.assign(an, an + 4 * n);
    a.rname.assign(rn, rn + 3 * n);
    a.chain.assign(ch, ch + n);
    a.serial.assign(rs, rs + n);
    return adopt_compact(compact_from_atoms(a));
}
fdh_compact *fdh_compact_from_soa(int64_t n, const float *nx, const float *cax, const float *cbx, const uint8_t *cbv,
                                  const uint8_t *aa, const uint8_t *chain, const uint64_t *serial, const float *bf) {
    fdh_compact *c = new fdh_compact();
    c->n.assign(nx, nx + 3 * n);
    c->ca.assign(cax, cax + 3 * n);
    c->cb.assign(cbx, cbx + 3 * n);
    if (cbv) c->cb_valid.assign(cbv, cbv + n);
    else c->cb_valid.assign((size_t)n, 1);
    c->aa.assign(aa, aa + n);
    if (chain) c->chain.assign(chain, chain + n);
    else c->chain.assign((size_t)n, (uint8_t)'A');
    c->serial.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) c->serial[i] = serial ? serial[i] : (uint64_t)(i + 1);
    if (bf) c->bfac.assign(bf, bf + n);
    else c->bfac.assign((size_t)n, 0.f);
    if (n) c->chains.push_back(c->chain[0]);
    c->raw_residues = (uint64_t)n;
    return adopt_compact(c);
}
int64_t fdh_compact_nres(const fdh_compact *c) { return (int64_t)c->nres(); }
int64_t fdh_compact_num_residues_raw(const fdh_compact *c) { return (int64_t)c->raw_residues; }
int fdh_compact_first_chain(const fdh_compact *c) { return c->chains.empty() ? -1 : c->chains[0]; }
float fdh_compact_avg_plddt(const fdh_compact *c) { // core.rs:446-456
    if (c->nres() == 0) return 0.f; // a skipped structure's slot (mod.rs:313-319): nres 0, plddt 0, not 0 / 0
    float s = 0.f;
    for (float b : c->bfac) s += b;
    return s / (float)c->nres();
}
void fdh_compact_get(const fdh_compact *c, float *nx, float *cax, float *cbx, uint8_t *cbv, uint8_t *aa,
                     uint8_t *chain, uint64_t *serial, float *bf) {
    const size_t n = c->nres();
    if (nx) memcpy(nx, c->n.data(), 12 * n);
    if (cax) memcpy(cax, c->ca.data(), 12 * n);
    if (cbx) memcpy(cbx, c->cb.data(), 12 * n);
    if (cbv) memcpy(cbv, c->cb_valid.data(), n);
    if (aa) memcpy(aa, c->aa.data(), n);
    if (chain) memcpy(chain, c->chain.data(), n);
    if (serial) memcpy(serial, c->serial.data(), 8 * n);
    if (bf) memcpy(bf, c->bfac.data(), 4 * n);
}
void fdh_compact_free(fdh_compact *c) {
    if (!c) return;
    std::shared_ptr<const fdh_compact> last = std::move(c->self.p); // destroyed here unless a query batch still shares it
    if (!last) delete c;
}

// parse_path_by_id_type (src/controller/mode.rs:19-31, 70-125): the structure id written to PREFIX.lookup for an input
// path under `--id TYPE`.  Returns the length of the id (written to out with a terminating 0 when it fits in cap).
int64_t fdh_parse_path_by_id_type(const char *path_c, const char *id_type_c, char *out, uint64_t cap) {
    const std::string path = path_c, id_type = id_type_c;
    auto is = [&](std::initializer_list<const char *> a) {
        for (const char *x : a)
            if (id_type == x) return true;
        return false;
    };
    auto file_name = [&] {
        const size_t sl = path.find_last_of('/');
        return sl == std::string::npos ? path : path.substr(sl + 1);
    };
    auto file_stem = [&] { // Path::file_stem: the name up to its last '.', unless that is its first character
        std::string f = file_name();
        const size_t dot = f.find_last_of('.');
        if (dot != std::string::npos && dot > 0) f = f.substr(0, dot);
        return f;
    };
    // leftmost match of the regex AF-.+-model_v\d, `.+` greedy: from the first "AF-" that can match, up to the LAST
    // "-model_v<digit>" at least one character later
    auto afdb = [&](const std::string &f, size_t *b, size_t *e) {
        for (size_t st = f.find("AF-"); st != std::string::npos; st = f.find("AF-", st + 1))
            for (size_t m = f.rfind("-model_v"); m != std::string::npos && m >= st + 4; m = m ? f.rfind("-model_v", m - 1) : std::string::npos)
                if (m + 8 < f.size() && isdigit((unsigned char)f[m + 8])) {
                    *b = st;
                    *e = m + 9;
                    return true;
                }
        return false;
    };
    std::string id;
    if (is({"Pdb", "PDB", "pdb"})) {
        const std::string st = file_stem();
        id = st.compare(0, 3, "pdb") == 0 ? st.substr(3) : st;
    } else if (is({"Afdb", "AFDB", "afdb"}) || is({"Uniprot", "UniProt", "uniprot"})) {
        const std::string st = file_stem();
        size_t b = 0, e = 0;
        if (!afdb(st, &b, &e)) {
            id = st;
        } else if (is({"Afdb", "AFDB", "afdb"})) {
            id = st.substr(b, e - b);
        } else { // the second '-'-separated field of the match (mode.rs:105-107)
            const std::string m = st.substr(b, e - b);
            const size_t d1 = m.find('-'), d2 = m.find('-', d1 + 1);
            id = m.substr(d1 + 1, d2 == std::string::npos ? std::string::npos : d2 - d1 - 1);
        }
    } else if (is({"BasenameWithoutExt", "basename_without_ext", "basename_no_ext", "filename"})) {
        id = file_stem();
    } else if (is({"BasenameWithExt", "basename_with_ext", "basename", "file"})) {
        id = file_name();
    } else if (is({"AbsPath", "Abspath", "abspath", "absolute_path", "path"})) {
        char buf[PATH_MAX];
        id = realpath(path.c_str(), buf) ? std::string(buf) : path;
    } else {
        id = path; // RelPath / Other
    }
    if (out && cap > id.size()) memcpy(out, id.c_str(), id.size() + 1);
    return (int64_t)id.size();
}

// ---- store ----
fdh_store *fdh_store_new(void) { return new fdh_store(); }
int64_t fdh_store_add(fdh_store *s, const fdh_compact *c, const char *name) {
    s->n.insert(s->n.end(), c->n.begin(), c->n.end());
    s->ca.insert(s->ca.end(), c->ca.begin(), c->ca.end());
    s->cb.insert(s->cb.end(), c->cb.begin(), c->cb.end());
    s->aa.insert(s->aa.end(), c->aa.begin(), c->aa.end());
    s->cb_valid.insert(s->cb_valid.end(), c->cb_valid.begin(), c->cb_valid.end());
    s->chain.insert(s->chain.end(), c->chain.begin(), c->chain.end());
    s->serial.insert(s->serial.end(), c->serial.begin(), c->serial.end());
    s->row_offsets.push_back(s->row_offsets.back() + c->nres());
    s->names.push_back(name ? name : "");
    s->plddt.push_back(fdh_compact_avg_plddt(c));
    return (int64_t)s->names.size() - 1;
}
int64_t fdh_store_add_soa(fdh_store *s, uint64_t S, const uint64_t *ro, const float *nx, const float *cax,
                          const float *cbx, const uint8_t *aa, const char *prefix) {
    const uint64_t R = ro[S];
    const uint64_t base = s->row_offsets.back();
    s->n.insert(s->n.end(), nx, nx + 3 * R);
    s->ca.insert(s->ca.end(), cax, cax + 3 * R);
    s->cb.insert(s->cb.end(), cbx, cbx + 3 * R);
    s->aa.insert(s->aa.end(), aa, aa + R);
    s->cb_valid.insert(s->cb_valid.end(), R, 1);
    s->chain.insert(s->chain.end(), R, (uint8_t)'A');
    const int64_t first = (int64_t)s->names.size();
    for (uint64_t k = 0; k < S; k++) {
        for (uint64_t r = ro[k]; r < ro[k + 1]; r++) s->serial.push_back(r - ro[k] + 1);
        s->row_offsets.push_back(base + ro[k + 1]);
        s->names.push_back(std::string(prefix ? prefix : "s") + std::to_string(first + (int64_t)k));
        s->plddt.push_back(0.f);
    }
    return first;
}
// ---- PREFIX.store: on-disk companion of the index holding the compact structures (SoA), so that a query run needs
// no per-candidate file parse (the reference re-reads and re-parses every candidate, retrieve.rs:375).  Layout:
//   "FDB2STR1" | u64 n_structs | u64 n_residues | row_offsets u64[S+1] | n, ca, cb f32[3R] each | aa u8[R] |
//   cb_valid u8[R] | chain u8[R] | serial u64[R] | plddt f32[S] | names: S x (u32 length, bytes)
int fdh_store_save(const fdh_store *s, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) {
        set_err(std::string("cannot write ") + path);
        return FD_ERR_ARG;
    }
    const uint64_t S = s->names.size(), R = s->row_offsets.back();
    bool ok = fwrite("FDB2STR1", 1, 8, f) == 8 && fwrite(&S, 8, 1, f) == 1 && fwrite(&R, 8, 1, f) == 1;
    auto wr = [&](const void *p, size_t bytes) { ok = ok && (bytes == 0 || fwrite(p, 1, bytes, f) == bytes); };
    wr(s->row_offsets.data(), 8 * (S + 1));
    wr(s->n.data(), 12 * R);
    wr(s->ca.data(), 12 * R);
    wr(s->cb.data(), 12 * R);
    wr(s->aa.data(), R);
    wr(s->cb_valid.data(), R);
    wr(s->chain.data(), R);
    wr(s->serial.data(), 8 * R);
    wr(s->plddt.data(), 4 * S);
    for (uint64_t k = 0; k < S; k++) {
        const uint32_t len = (uint32_t)s->names[k].size();
        wr(&len, 4);
        wr(s->names[k].data(), len);
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) {
        set_err(std::string("write error on ") + path);
        return FD_ERR_ARG;
    }
    return FD_OK;
}

fdh_store *fdh_store_load(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) {
        set_err(std::string("cannot open ") + path);
        return nullptr;
    }
    char magic[8];
    uint64_t S = 0, R = 0;
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "FDB2STR1", 8) == 0 && fread(&S, 8, 1, f) == 1 &&
              fread(&R, 8, 1, f) == 1;
    if (ok) { // the header must be consistent with the file size before anything is allocated from it
        struct stat st;
        const bool sane = S < (1ull << 40) && R < (1ull << 40);
        const uint64_t need = sane ? 24 + 8 * (S + 1) + 36 * R + 3 * R + 8 * R + 4 * S + 4 * S : 0;
        ok = sane && fstat(fileno(f), &st) == 0 && (uint64_t)st.st_size >= need;
    }
    fdh_store *s = new fdh_store();
    auto rd = [&](void *p, size_t bytes) { ok = ok && (bytes == 0 || fread(p, 1, bytes, f) == bytes); };
    if (ok) {
        s->row_offsets.resize(S + 1);
        s->n.resize(3 * R);
        s->ca.resize(3 * R);
        s->cb.resize(3 * R);
        s->aa.resize(R);
        s->cb_valid.resize(R);
        s->chain.resize(R);
        s->serial.resize(R);
        s->plddt.resize(S);
        s->names.resize(S);
        rd(s->row_offsets.data(), 8 * (S + 1));
        rd(s->n.data(), 12 * R);
        rd(s->ca.data(), 12 * R);
        rd(s->cb.data(), 12 * R);
        rd(s->aa.data(), R);
        rd(s->cb_valid.data(), R);
        rd(s->chain.data(), R);
        rd(s->serial.data(), 8 * R);
        rd(s->plddt.data(), 4 * S);
        for (uint64_t k = 0; k < S && ok; k++) {
            uint32_t len = 0;
            rd(&len, 4);
            if (ok && len > (1u << 20)) ok = false;
            if (ok) {
                s->names[k].resize(len);
                rd(&s->names[k][0], len);
            }
        }
        ok = ok && s->row_offsets[0] == 0 && s->row_offsets[S] == R;
        for (uint64_t k = 0; ok && k < S; k++) ok = s->row_offsets[k] <= s->row_offsets[k + 1]; // labels are indexed through them
    }
    fclose(f);
    if (!ok) {
        set_err(std::string("not a structure store (or truncated): ") + path);
        delete s;
        return nullptr;
    }
    return s;
}
uint64_t fdh_store_size(const fdh_store *s) { return s->names.size(); }
uint64_t fdh_store_num_residues(const fdh_store *s) { return s->row_offsets.back(); }
void fdh_store_get_lookup(const fdh_store *s, uint32_t *nres, float *plddt) {
    for (size_t k = 0; k < s->names.size(); k++) {
        if (nres) nres[k] = (uint32_t)(s->row_offsets[k + 1] - s->row_offsets[k]);
        if (plddt) plddt[k] = s->plddt[k];
    }
}
const char *fdh_store_name(const fdh_store *s, uint64_t id) { return id < s->names.size() ? s->names[id].c_str() : ""; }
int fdh_store_batch(const fdh_store *s, fd_struct_batch *out) {
    out->n_structs = s->names.size();
    out->row_offsets = s->row_offsets.data();
    out->n_xyz = s->n.data();
    out->ca_xyz = s->ca.data();
    out->cb_xyz = s->cb.data();
    out->aa = s->aa.data();
    out->cb_valid = s->cb_valid.data();
    return FD_OK;
}
int fdh_store_attach(const fdh_store *s, fd_ctx *ctx) {
    fd_struct_batch b;
    fdh_store_batch(s, &b);
    int rc = fd_store_attach(ctx, &b);
    if (rc == FD_OK && s->chain.size() == s->aa.size() && s->serial.size() == s->aa.size() && !s->aa.empty())
        rc = fd_store_attach_labels(ctx, s->chain.data(), s->serial.data(), s->aa.size());
    if (rc != FD_OK) set_err(fd_last_error(ctx));
    return rc;
}
void fdh_store_free(fdh_store *s) { delete s; }

// ---- index ----
// Folddisco::collect_and_count .. add_entries on the GPU.  The sort/encode stage needs ~50 bytes of device memory
// per (hash, structure) posting, so large databases are built in hash-range chunks (every chunk re-hashes the
// structures, which is cheap, and keeps only its range); the chunks concatenate to the index byte for byte.
fdh_index *fdh_index_build(fd_ctx *ctx, const fdh_store *s, const fd_hash_params *params) {
    fd_struct_batch b;
    fdh_store_batch(s, &b);
    const uint64_t S = b.n_structs, R = S ? b.row_offsets[S] : 0;
    uint64_t chunk_keys = 600ull * 1000 * 1000;
    if (const char *e = getenv("FD_BUILD_CHUNK_KEYS")) chunk_keys = std::max<uint64_t>(1000, strtoull(e, nullptr, 10));
    std::vector<uint64_t> bounds{0, 1ull << 32};
    if (R * 100 > chunk_keys) { // ~100 postings per residue: more than one chunk is likely, so plan from a sample
        const uint64_t n_sample = std::min<uint64_t>(S, 256);
        fd_struct_batch sb = b;
        sb.n_structs = n_sample;
        uint32_t *sh = nullptr;
        uint64_t *sro = nullptr;
        if (fd_hash_structures(ctx, &sb, params, &sh, &sro) != FD_OK) {
            set_err(fd_last_error(ctx));
            return nullptr;
        }
        const uint64_t n_sh = sro[n_sample], r_sample = std::max<uint64_t>(1, b.row_offsets[n_sample]);
        std::vector<uint64_t> hist(4096, 0); // by hash >> 20
        for (uint64_t k = 0; k < n_sh; k++) hist[sh[k] >> 20]++;
        fd_free(sh);
        fd_free(sro);
        const double scale = (double)R / (double)r_sample;
        bounds.assign(1, 0);
        double acc = 0;
        for (uint32_t p = 0; p < 4096; p++) {
            const double add = hist[p] * scale;
            if (acc > 0 && acc + add > (double)chunk_keys) {
                bounds.push_back((uint64_t)p << 20);
                acc = 0;
            }
            acc += add;
        }
        bounds.push_back(1ull << 32);
    }
    fd_index_buffers all;
    memset(&all, 0, sizeof(all));
    std::vector<uint32_t> hashes;
    std::vector<uint64_t> offsets(1, 0);
    std::vector<uint8_t> values;
    const bool single = bounds.size() == 2;
    for (size_t c = 0; c + 1 < bounds.size(); c++) {
        fd_index_buffers out;
        if (fd_build_index(ctx, &b, params, 0, bounds[c], bounds[c + 1], &out) != FD_OK) {
            set_err(fd_last_error(ctx));
            return nullptr;
        }
        if (single) {
            fdh_index *ix = fdh_index_from_buffers(&out, s, params);
            fd_free_index_buffers(&out);
            return ix;
        }
        hashes.insert(hashes.end(), out.hashes, out.hashes + out.count);
        const uint64_t base = values.size();
        for (uint64_t k = 1; k <= out.count; k++) offsets.push_back(base + out.offsets[k]);
        values.insert(values.end(), out.values, out.values + out.value_bytes);
        fd_free_index_buffers(&out);
    }
    all.count = hashes.size();
    all.hashes = hashes.data();
    all.offsets = offsets.data();
    all.value_bytes = values.size();
    all.values = values.data();
    return fdh_index_from_buffers(&all, s, params);
}

fdh_index *fdh_index_from_buffers(const fd_index_buffers *out, const fdh_store *s, const fd_hash_params *params) {
    fdh_index *ix = new fdh_index();
    ix->own_hashes.assign(out->hashes, out->hashes + out->count);
    ix->own_offsets.assign(out->offsets, out->offsets + out->count + 1);
    ix->own_values.assign(out->values, out->values + out->value_bytes);
    ix->hashes = ix->own_hashes.data();
    ix->offsets = ix->own_offsets.data();
    ix->values = ix->own_values.data();
    ix->count = ix->own_hashes.size();
    ix->value_bytes = ix->own_values.size();
    ix->names = s->names;
    ix->nres.resize(s->names.size());
    ix->plddt.resize(s->names.size());
    fdh_store_get_lookup(s, ix->nres.data(), ix->plddt.data());
    ix->params = *params;
    return ix;
}

// HashType::get_with_str / to_string (src/geometry/core.rs:42-75) -> FD_HASH_*; -1 = unknown
int fdh_hash_type_from_string(const char *name) {
    const std::string t = name ? name : "";
    auto is = [&](std::initializer_list<const char *> l) {
        for (const char *x : l)
            if (t == x) return true;
        return false;
    };
    if (is({"0", "PDBMotif", "pyscomotif", "orig_pdb"})) return FD_HASH_PDBMOTIF;
    if (is({"1", "PDBMotifSinCos", "pdb"})) return FD_HASH_PDBMOTIFSINCOS;
    if (is({"2", "TrRosetta", "trrosetta", "tr"})) return FD_HASH_TRROSETTA;
    if (is({"3", "PDBTrRosetta", "pdbtr", "default", "folddisco"})) return FD_HASH_PDBTRROSETTA;
    if (is({"4", "PointPairFeature", "ppf"})) return FD_HASH_POINTPAIRFEATURE;
    if (is({"5", "TertiaryInteraction", "tertiary", "3di"})) return FD_HASH_TERTIARYINTERACTION;
    if (is({"6", "Hybrid", "hybrid"})) return FD_HASH_HYBRID;
    if (is({"7", "FolddiscoAngle", "angle", "folddisco_angle"})) return FD_HASH_FOLDDISCOANGLE;
    if (is({"8", "FolddiscoDist", "distance", "dist", "folddisco_dist"})) return FD_HASH_FOLDDISCODIST;
    return -1;
}
const char *hash_type_name(uint32_t t) {
    switch (t) {
        case FD_HASH_PDBMOTIF: return "PDBMotif";
        case FD_HASH_PDBMOTIFSINCOS: return "PDBMotifSinCos";
        case FD_HASH_TRROSETTA: return "TrRosetta";
        case FD_HASH_POINTPAIRFEATURE: return "PointPairFeature";
        case FD_HASH_TERTIARYINTERACTION: return "TertiaryInteraction";
        case FD_HASH_HYBRID: return "Hybrid";
        case FD_HASH_FOLDDISCOANGLE: return "FolddiscoAngle";
        case FD_HASH_FOLDDISCODIST: return "FolddiscoDist";
        default: return "PDBTrRosetta";
    }
}
const char *fdh_hash_type_name(uint32_t hash_type) { return hash_type_name(hash_type); }

int fdh_index_save(const fdh_index *ix, const fdh_store *s, const char *prefix, uint64_t max_residue,
                   const char *foldcomp_db) {
    auto wr = [&](const std::string &path, const void *p, size_t n, FILE *f) {
        return n == 0 || fwrite(p, 1, n, f) == n;
    };
    { // PREFIX: posting bytes (indextable.rs:239-264)
        FILE *f = fopen(prefix, "wb");
        if (!f || !wr(prefix, ix->values, ix->value_bytes, f)) {
            set_err(std::string("cannot write ") + prefix);
            if (f) fclose(f);
            return FD_ERR_ARG;
        }
        fclose(f);
    }
    { // PREFIX.offset: u64 count | u32 hashes | u64 offsets (indextable.rs:297-326)
        std::string p = std::string(prefix) + ".offset";
        FILE *f = fopen(p.c_str(), "wb");
        uint64_t count = ix->count;
        if (!f || !wr(p, &count, 8, f) || !wr(p, ix->hashes, 4 * count, f) || !wr(p, ix->offsets, 8 * (count + 1), f)) {
            set_err("cannot write " + p);
            if (f) fclose(f);
            return FD_ERR_ARG;
        }
        fclose(f);
    }
    { // PREFIX.lookup: id \t name \t nres \t plddt \t db_key (lookup.rs:17-58)
        std::string p = std::string(prefix) + ".lookup";
        FILE *f = fopen(p.c_str(), "wb");
        if (!f) {
            set_err("cannot write " + p);
            return FD_ERR_ARG;
        }
        (void)s;
        for (size_t i = 0; i < ix->names.size(); i++)
            fprintf(f, "%zu\t%s\t%u\t%s\t%llu\n", i, ix->names[i].c_str(), ix->nres[i], rust_f32(ix->plddt[i]).c_str(),
                    (unsigned long long)(i < ix->db_key.size() ? ix->db_key[i] : (uint64_t)i));
        fclose(f);
    }
    { // PREFIX.type: TOML with sorted keys (cli/config.rs:64-97)
        std::string p = std::string(prefix) + ".type";
        FILE *f = fopen(p.c_str(), "wb");
        if (!f) {
            set_err("cannot write " + p);
            return FD_ERR_ARG;
        }
        char buf[64];
        auto r = std::to_chars(buf, buf + sizeof(buf), (double)ix->params.dist_cutoff, std::chars_format::fixed);
        std::string g(buf, r.ptr);
        if (g.find('.') == std::string::npos) g += ".0";
        fprintf(f, "chunk_size = %zu\n", ix->names.size());
        if (foldcomp_db) fprintf(f, "foldcomp_db = \"%s\"\n", foldcomp_db);
        fprintf(f, "grid_width = %s\nhash_type = \"%s\"\ninput_format = \"%s\"\nmax_residue = %llu\n", g.c_str(),
                hash_type_name(ix->params.hash_type), ix->fcz_input ? "FCZDB" : "PDB", (unsigned long long)max_residue);
        if (ix->params.n_multiple_bins) { // multiple_bin = [[16, 4], [8, 3]] (config.rs:78-85)
            fprintf(f, "multiple_bin = [");
            for (uint32_t k = 0; k < ix->params.n_multiple_bins; k++)
                fprintf(f, "%s[%u, %u]", k ? ", " : "", ix->params.multiple_bins[2 * k], ix->params.multiple_bins[2 * k + 1]);
            fprintf(f, "]\n");
        }
        fprintf(f, "num_bin_angle = %u\nnum_bin_dist = %u\n", ix->params.nbin_angle, ix->params.nbin_dist);
        fclose(f);
    }
    return FD_OK;
}

fdh_index *fdh_index_load(const char *prefix) {
    auto map_file = [](const std::string &p, void **addr, size_t *len) {
        int fd = open(p.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) {
            close(fd);
            return false;
        }
        *len = (size_t)st.st_size;
        *addr = *len ? mmap(nullptr, *len, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
        close(fd);
        return *len == 0 || *addr != MAP_FAILED;
    };
    fdh_index *ix = new fdh_index();
    std::string vp = std::string(prefix) + ".value";
    if (access(vp.c_str(), F_OK) != 0) vp = prefix; // indextable.rs:333-337
    if (!map_file(vp, &ix->map_val, &ix->map_val_len) || !map_file(std::string(prefix) + ".offset", &ix->map_off, &ix->map_off_len)) {
        set_err(std::string("cannot open index ") + prefix);
        ix->map_off = ix->map_val = nullptr;
        delete ix;
        return nullptr;
    }
    const uint8_t *o = (const uint8_t *)ix->map_off;
    if (ix->map_off_len < 8) {
        set_err("offset file too short");
        delete ix;
        return nullptr;
    }
    memcpy(&ix->count, o, 8);
    // a corrupt count must not wrap the size computation (12 bytes per hash + the 16 bytes around them)
    if (ix->count > (ix->map_off_len - 8) / 12) {
        set_err("Offset file appears to be in old format or corrupted");
        delete ix;
        return nullptr;
    }
    const size_t need = 8 + ix->count * 4 + (ix->count + 1) * 8;
    if (ix->map_off_len < need) { // indextable.rs:347-355
        set_err("Offset file appears to be in old format or corrupted");
        delete ix;
        return nullptr;
    }
    ix->hashes = (const uint32_t *)(o + 8);
    ix->own_offsets.resize(ix->count + 1); // unaligned when count is odd: copy
    memcpy(ix->own_offsets.data(), o + 8 + ix->count * 4, (ix->count + 1) * 8);
    ix->offsets = ix->own_offsets.data();
    ix->values = (const uint8_t *)ix->map_val;
    ix->value_bytes = ix->map_val_len;
    // the offsets index the value file: a pair of files that do not belong together must not reach the device, and
    // an offset that runs backwards would send the scan out of bounds
    bool monotone = true;
    for (uint64_t k = 0; k < ix->count && monotone; k++) monotone = ix->offsets[k] <= ix->offsets[k + 1];
    if (!monotone) {
        set_err(std::string("index files are corrupt: the offsets of ") + prefix + ".offset are not ascending");
        delete ix;
        return nullptr;
    }
    if (ix->offsets[0] != 0 || ix->offsets[ix->count] != ix->value_bytes) {
        set_err(std::string("index files are inconsistent: the offsets of ") + prefix + ".offset end at byte " +
                std::to_string(ix->offsets[ix->count]) + " but the value file has " + std::to_string(ix->value_bytes));
        delete ix;
        return nullptr;
    }
    { // lookup
        std::ifstream in(std::string(prefix) + ".lookup");
        std::string line;
        while (std::getline(in, line)) {
            size_t a = line.find('\t'), b = line.find('\t', a + 1), c = line.find('\t', b + 1);
            if (a == std::string::npos || b == std::string::npos || c == std::string::npos) continue;
            size_t d = line.find('\t', c + 1);
            ix->names.push_back(line.substr(a + 1, b - a - 1));
            ix->nres.push_back((uint32_t)strtoul(line.substr(b + 1, c - b - 1).c_str(), nullptr, 10));
            ix->plddt.push_back(strtof(line.substr(c + 1, d == std::string::npos ? std::string::npos : d - c - 1).c_str(), nullptr));
            ix->db_key.push_back(d == std::string::npos ? (uint64_t)(ix->names.size() - 1)
                                                        : (uint64_t)strtoull(line.c_str() + d + 1, nullptr, 10));
        }
    }
    { // type (TOML subset written by the reference)
        std::ifstream in(std::string(prefix) + ".type");
        std::string line;
        while (std::getline(in, line)) {
            size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            std::string k = line.substr(0, eq), v = line.substr(eq + 1);
            while (!k.empty() && k.back() == ' ') k.pop_back();
            if (k == "num_bin_dist") ix->params.nbin_dist = (uint32_t)atoi(v.c_str());
            else if (k == "num_bin_angle") ix->params.nbin_angle = (uint32_t)atoi(v.c_str());
            else if (k == "grid_width") ix->params.dist_cutoff = strtof(v.c_str(), nullptr);
            else if (k == "hash_type") {
                const size_t a = v.find('"'), b = v.rfind('"');
                const int t = a != std::string::npos && b > a ? fdh_hash_type_from_string(v.substr(a + 1, b - a - 1).c_str()) : -1;
                if (t < 0) {
                    set_err("unknown hash_type in " + std::string(prefix) + ".type:" + v);
                    delete ix;
                    return nullptr;
                }
                ix->params.hash_type = (uint32_t)t;
            } else if (k == "input_format" || k == "foldcomp_db") {
                const size_t a = v.find('"'), b = v.rfind('"');
                const std::string val = a != std::string::npos && b > a ? v.substr(a + 1, b - a - 1) : "";
                if (k == "foldcomp_db") ix->foldcomp_db = val;
                else ix->fcz_input = val == "FCZDB" || val == "fczdb" || val == "4"; // StructureFileFormat::get_with_string
            } else if (k == "multiple_bin") {
                std::vector<uint32_t> nums;
                for (size_t pos = 0; pos < v.size();) {
                    if (isdigit((unsigned char)v[pos])) {
                        char *end = nullptr;
                        nums.push_back((uint32_t)strtoul(v.c_str() + pos, &end, 10));
                        pos = (size_t)(end - v.c_str());
                    } else {
                        pos++;
                    }
                }
                if (nums.size() % 2 != 0 || nums.size() / 2 > FD_MAX_MULTIPLE_BINS) {
                    set_err("multiple_bin in " + std::string(prefix) + ".type: expected at most 8 [dist, angle] pairs");
                    delete ix;
                    return nullptr;
                }
                ix->params.n_multiple_bins = (uint32_t)(nums.size() / 2);
                for (size_t q = 0; q < nums.size(); q++) ix->params.multiple_bins[q] = nums[q];
            }
        }
    }
    return ix;
}
int fdh_index_get(const fdh_index *ix, fd_index_buffers *v) {
    v->count = ix->count;
    v->hashes = const_cast<uint32_t *>(ix->hashes);
    v->offsets = const_cast<uint64_t *>(ix->offsets);
    v->value_bytes = ix->value_bytes;
    v->values = const_cast<uint8_t *>(ix->values);
    return FD_OK;
}
uint64_t fdh_index_num_structs(const fdh_index *ix) { return ix->nres.size(); }
void fdh_index_get_lookup(const fdh_index *ix, uint32_t *nres, float *plddt) {
    if (nres) memcpy(nres, ix->nres.data(), 4 * ix->nres.size());
    if (plddt) memcpy(plddt, ix->plddt.data(), 4 * ix->plddt.size());
}
uint64_t fdh_index_db_key(const fdh_index *ix, uint64_t id) { return id < ix->db_key.size() ? ix->db_key[id] : id; }
// "" unless PREFIX.type names a Foldcomp database AND the index was built from it (input_format = "FCZDB":
// query_pdb.rs:322 `using_foldcomp`)
const char *fdh_index_foldcomp_db(const fdh_index *ix) { return ix->fcz_input ? ix->foldcomp_db.c_str() : ""; }
// index built from a Foldcomp database: the database keys of the structures (Folddisco::numeric_db_key_vec, mod.rs:151;
// written as the 5th lookup column, lookup.rs:36-40) -- fdh_index_save then records input_format = "FCZDB"
int fdh_index_set_db_keys(fdh_index *ix, const uint64_t *keys, uint64_t n) {
    if (n != ix->names.size()) {
        set_err("fdh_index_set_db_keys: one key per structure");
        return FD_ERR_ARG;
    }
    ix->db_key.assign(keys, keys + n);
    ix->fcz_input = true;
    return FD_OK;
}
const char *fdh_index_name(const fdh_index *ix, uint64_t id) { return id < ix->names.size() ? ix->names[id].c_str() : ""; }
void fdh_index_get_params(const fdh_index *ix, fd_hash_params *p) { *p = ix->params; }
int fdh_index_attach(fd_ctx *ctx, const fdh_index *ix) {
    return fd_index_attach(ctx, ix->hashes, ix->offsets, ix->count, ix->values, ix->value_bytes, ix->nres.size(),
                           ix->nres.data(), ix->plddt.data());
}
void fdh_index_free(fdh_index *ix) { delete ix; }

// ---- queries ----
int64_t fdh_parse_query_string(const char *q, uint8_t default_chain, uint8_t *chains, uint64_t *serials,
                               int64_t *subs_off, int64_t *subs_end, uint8_t *subs, int64_t cap_res, int64_t cap_subs) {
    ParsedQuery pq;
    if (!parse_query(q, default_chain, pq)) return -1;
    int64_t ns = 0;
    if ((int64_t)pq.chains.size() > cap_res) return -2;
    for (size_t i = 0; i < pq.chains.size(); i++) {
        chains[i] = pq.chains[i];
        serials[i] = pq.serials[i];
        if (pq.has_sub[i]) {
            subs_off[i] = ns;
            for (uint8_t v : pq.subs[i]) {
                if (ns >= cap_subs) return -2;
                subs[ns++] = v;
            }
            subs_end[i] = ns;
        } else {
            subs_off[i] = -1;
            subs_end[i] = -1;
        }
    }
    return (int64_t)pq.chains.size();
}

fdh_queries *fdh_queries_new(const fdh_query_params *p) {
    {
        fdg::TypedParams tp;
        if (const char *why = fdg::typed_params_from(&p->hash, &tp)) {
            set_err(why);
            return nullptr;
        }
    }
    fdh_queries *qs = new fdh_queries();
    qs->p = *p;
    qs->dist_thr.assign(p->dist_thr, p->dist_thr + p->n_dist_thr);
    qs->angle_thr.assign(p->angle_thr, p->angle_thr + p->n_angle_thr);
    qs->p.dist_thr = nullptr;
    qs->p.angle_thr = nullptr;
    return qs;
}
static bool prepare_query(const fdh_queries *qs, std::shared_ptr<const fdh_compact> st, const char *query_string,
                          Query &Q, std::string &err) {
    ParsedQuery pq;
    const int fc = fdh_compact_first_chain(st.get());
    if (!parse_query(query_string, fc < 0 ? (uint8_t)'A' : (uint8_t)fc, pq)) {
        err = std::string("Invalid residue in query string: ") + query_string;
        return false;
    }
    if (pq.chains.empty()) {
        // an empty query string makes every residue of the structure a query residue (query.rs:226-233)
        for (size_t i = 0; i < st->nres(); i++) {
            pq.chains.push_back(st->chain[i]);
            pq.serials.push_back(st->serial[i]);
            pq.has_sub.push_back(0);
            pq.subs.emplace_back();
        }
    }
    Q.st = std::move(st);
    Q.qstring = query_string;
    Q.residue_count = (uint32_t)pq.chains.size();
    if (!build_query_map(Q, pq, *qs)) {
        err = "query has more than 65535 edges (residue pairs with a feature); whole-structure queries are limited to chains of a few hundred residues";
        return false;
    }
    return true;
}


int64_t fdh_queries_add(fdh_queries *qs, const fdh_compact *st, const char *query_string) {
    Query Q;
    std::string err;
    if (!prepare_query(qs, share_compact(st), query_string, Q, err)) {
        set_err(err);
        return -1;
    }
    qs->q.push_back(std::move(Q));
    qs->finalized = false;
    qs->generation++;
    return (int64_t)qs->q.size() - 1;
}

// make_query_map for n (structure, query string) pairs at once, query-parallel (query_pdb.rs:348 into_par_iter).
// Query k uses structs[which_struct[k]] and strings[which_string[k]] (NULL index arrays = k itself).
static int64_t add_many_impl(fdh_queries *qs, const fdh_compact *const *structs, int64_t n_structs,
                             const char *const *strings, int64_t n_strings, const uint32_t *which_struct,
                             const uint32_t *which_string, int64_t n, int threads) {
    for (int64_t k = 0; k < n; k++)
        if ((which_struct && which_struct[k] >= n_structs) || (which_string && which_string[k] >= n_strings)) {
            set_err("fdh_queries_add_many_indexed: index out of range");
            return -1;
        }
    // the structures are shared with the handles, not copied (share_compact)
    std::vector<std::shared_ptr<const fdh_compact>> uniq((size_t)n_structs);
    for (int64_t k = 0; k < n; k++) {
        const int64_t u = which_struct ? which_struct[k] : k;
        if (!uniq[u]) uniq[u] = share_compact(structs[u]);
    }
    int nt = threads > 0 ? threads : fd_default_host_threads();
    nt = std::max(1, std::min<int>(nt, 64));
    std::vector<Query> out((size_t)n);
    std::vector<std::string> errs((size_t)n);
    std::vector<uint8_t> ok((size_t)n, 0);
    std::atomic<int64_t> next{0};
    auto worker = [&] {
        for (int64_t k; (k = next.fetch_add(1)) < n;)
            ok[k] = prepare_query(qs, uniq[which_struct ? which_struct[k] : k], strings[which_string ? which_string[k] : k],
                                  out[k], errs[k]);
    };
    fd_parallel(nt, [&](int) { worker(); });
    for (int64_t k = 0; k < n; k++)
        if (!ok[k]) {
            set_err(errs[k]);
            return -1;
        }
    const int64_t first = (int64_t)qs->q.size();
    qs->q.reserve(qs->q.size() + (size_t)n);
    for (auto &Q : out) qs->q.push_back(std::move(Q));
    qs->finalized = false;
    qs->generation++;
    return first;
}
int64_t fdh_queries_add_many(fdh_queries *qs, const fdh_compact *const *structures, const char *const *query_strings,
                             int64_t n, int threads) {
    return add_many_impl(qs, structures, n, query_strings, n, nullptr, nullptr, n, threads);
}
int64_t fdh_queries_add_many_indexed(fdh_queries *qs, const fdh_compact *const *structures, int64_t n_structures,
                                     const char *const *query_strings, int64_t n_strings, const uint32_t *which_structure,
                                     const uint32_t *which_string, int64_t n, int threads) {
    if (!which_structure || !which_string) {
        set_err("fdh_queries_add_many_indexed: NULL index array");
        return -1;
    }
    return add_many_impl(qs, structures, n_structures, query_strings, n_strings, which_structure, which_string, n, threads);
}
int64_t fdh_queries_size(const fdh_queries *qs) { return (int64_t)qs->q.size(); }

static int ensure_verify_prepared(fd_ctx *ctx, const fdh_queries *qs);
int fdh_queries_finalize(fdh_queries *qs, fd_ctx *ctx) {
    // calculate_idf_for_hash (query.rs:17-32) for the observed hash of every query pair, one device call
    std::vector<uint32_t> all;
    for (auto &Q : qs->q) all.insert(all.end(), Q.pair_hash.begin(), Q.pair_hash.end());
    std::vector<uint32_t> counts(all.size());
    if (!all.empty()) {
        int rc = fd_posting_counts(ctx, all.data(), all.size(), counts.data());
        if (rc != FD_OK) {
            set_err(fd_last_error(ctx));
            return rc;
        }
    }
    // total_structures = lookup.len() as f32
    const int rc = fdh_queries_finalize_with_counts(qs, counts.data(), fd_index_num_structs(ctx));
    if (rc != FD_OK) return rc;
    if (!fdg::ht_default_route(&qs->p.hash)) return FD_OK; // verified through the general path: no fused-kernel tables
    return ensure_verify_prepared(ctx, qs);
}
int64_t fdh_queries_num_hashes(const fdh_queries *qs, int64_t q) { return (int64_t)qs->q[q].entries.size(); }
void fdh_queries_get_map(const fdh_queries *qs, int64_t q, uint32_t *hash, int64_t *qi, int64_t *qj,
                         uint8_t *primary, float *idf) {
    const Query &Q = qs->q[q];
    for (size_t k = 0; k < Q.entries.size(); k++) {
        if (hash) hash[k] = Q.entries[k].hash;
        if (qi) qi[k] = Q.entries[k].qi;
        if (qj) qj[k] = Q.entries[k].qj;
        if (primary) primary[k] = Q.entries[k].primary;
        if (idf) idf[k] = Q.entries[k].idf;
    }
}
int64_t fdh_queries_num_indices(const fdh_queries *qs, int64_t q) { return (int64_t)qs->q[q].indices.size(); }
int64_t fdh_queries_residue_count(const fdh_queries *qs, int64_t q) { return (int64_t)qs->q[q].residue_count; }
void fdh_queries_get_indices(const fdh_queries *qs, int64_t q, int64_t *indices) {
    for (size_t k = 0; k < qs->q[q].indices.size(); k++) indices[k] = qs->q[q].indices[k];
}
// A batch of 1 024 query maps owns ~15 k heap blocks and seven device tables: tearing it down costs ~1 ms, which a
// serving loop would pay on its critical path between two batches.  The handle is queued to a background thread
// instead (started on first use, detached; batches still queued at process exit are simply not freed).
namespace {
struct Reaper {
    std::mutex m;
    std::condition_variable cv;
    std::vector<fdh_queries *> q;
    Reaper() {
        std::thread([this] {
            for (;;) {
                std::vector<fdh_queries *> take;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [this] { return !q.empty(); });
                    take.swap(q);
                }
                for (fdh_queries *x : take) delete x;
            }
        }).detach();
    }
    void push(fdh_queries *x) {
        {
            std::lock_guard<std::mutex> lk(m);
            q.push_back(x);
        }
        cv.notify_one();
    }
};
} // namespace
void fdh_queries_free(fdh_queries *qs) {
    if (!qs) return;
    if (getenv("FD_SYNC_FREE")) {
        delete qs;
        return;
    }
    static Reaper *reaper = new Reaper(); // leaked on purpose: the thread outlives static destruction
    reaper->push(qs);
}

} // extern "C"

// ---- search ----
namespace {

struct FinalMatch { // one verified component of one candidate
    uint64_t cand;
    uint32_t node_count;
    float idf, rmsd;
    float U[9], t[3];
    uint32_t n_res;
    uint32_t res16[16];            // per query residue: target residue index + 1, 0 = none
    std::vector<uint32_t> res_big; // only when the query has more than 16 residues (general path)
    const uint32_t *res() const { return n_res <= 16 ? res16 : res_big.data(); }
};

// General verification path for an arbitrary candidate list: K4 (fd_candidate_edges_batch) -> host graph /
// mapping / rescue (candidate-parallel) -> K5 (fd_kabsch_store_batch).  Handles what the fused kernel cannot.
int verify_general(fd_ctx *ctx, const fdh_queries *qs, uint32_t q_begin, uint32_t nq, const fdh_search_params *p,
                   const std::vector<uint32_t> &cand_q, const std::vector<uint32_t> &cand_n,
                   const std::vector<uint64_t> &cand_global, std::vector<FinalMatch> &out, fdh_results *R,
                   double *host_ms) {
    const uint64_t n_cand = cand_q.size();
    if (n_cand == 0) return FD_OK;
    std::vector<fd_retrieval_query> rq(nq);
    for (uint32_t q = 0; q < nq; q++) {
        const Query &Q = qs->q[q_begin + q];
        rq[q] = fd_retrieval_query{(uint32_t)Q.hashes_sorted.size(), Q.hashes_sorted.data(), (uint32_t)Q.aad.size(),
                                   Q.aad_aa1.data(), Q.aad_aa2.data(), Q.aad_dist.data(), Q.aad_qi.data()};
        R->h2d_bytes += 4ull * Q.hashes_sorted.size() + 12ull * Q.aad.size() + 28;
    }
    std::vector<fd_cand_edge> all_edges;
    std::vector<fd_cand_pair> all_pairs;
    const uint64_t CH = (1ull << 24) - 1; // the kernel takes at most 2^24-1 candidates per call
    for (uint64_t c0 = 0; c0 < n_cand; c0 += CH) {
        const uint64_t cn = std::min(CH, n_cand - c0);
        fd_cand_edge *edges = nullptr;
        fd_cand_pair *pairs = nullptr;
        uint64_t ne = 0, np = 0;
        if (fd_candidate_edges_batch(ctx, rq.data(), nq, cand_q.data() + c0, cand_n.data() + c0, cn, &qs->p.hash,
                                     p->ca_dist_cutoff, &edges, &ne, &pairs, &np) != FD_OK) {
            set_err(fd_last_error(ctx));
            return FD_ERR_CUDA;
        }
        R->d2h_bytes += ne * sizeof(fd_cand_edge) + np * sizeof(fd_cand_pair) + 16;
        R->h2d_bytes += 8ull * cn;
        for (uint64_t k = 0; k < ne; k++) {
            edges[k].cand += (uint32_t)c0;
            all_edges.push_back(edges[k]);
        }
        for (uint64_t k = 0; k < np; k++) {
            pairs[k].cand += (uint32_t)c0;
            all_pairs.push_back(pairs[k]);
        }
        fd_free(edges);
        fd_free(pairs);
    }
    auto t0 = std::chrono::steady_clock::now();
    std::vector<uint64_t> e_begin(n_cand + 1, 0), p_begin(n_cand + 1, 0);
    for (auto &e : all_edges) e_begin[e.cand + 1]++;
    for (auto &e : all_pairs) p_begin[e.cand + 1]++;
    for (uint64_t c = 0; c < n_cand; c++) {
        e_begin[c + 1] += e_begin[c];
        p_begin[c + 1] += p_begin[c];
    }
    int nt = p->host_threads > 0 ? p->host_threads : fd_default_host_threads();
    nt = std::max(1, std::min(nt, 64));
    std::vector<std::vector<MatchTmp>> per_thread(nt);
    std::atomic<uint64_t> next{0};
    // few, heavy candidates (the ones beyond the fused kernels' limits) must still spread over the threads
    const uint64_t GRAIN = std::max<uint64_t>(1, std::min<uint64_t>(256, n_cand / (4 * (uint64_t)nt)));
    struct Chunk {
        uint64_t c0;
        int thread;
        size_t begin, end;
    };
    std::vector<std::vector<Chunk>> chunks(nt);
    auto worker = [&](int tid) {
        std::vector<fd_cand_edge> ce;
        std::vector<std::vector<uint32_t>> comps;
        Graph g;
        std::vector<MatchTmp> &outv = per_thread[tid];
        for (;;) {
            const uint64_t c0 = next.fetch_add(GRAIN);
            if (c0 >= n_cand) break;
            const size_t begin = outv.size();
            for (uint64_t c = c0; c < std::min(n_cand, c0 + GRAIN); c++) {
                const uint64_t eb = e_begin[c], ee = e_begin[c + 1];
                if (ee == eb) continue;
                const Query &Q = qs->q[q_begin + cand_q[c]];
                ce.assign(all_edges.begin() + eb, all_edges.begin() + ee);
                g.node_res.clear(); // create_index_graph (graph.rs:16-27): node ids by first appearance
                g.e.clear();
                auto node_of = [&](uint32_t r) {
                    for (uint32_t k = 0; k < g.node_res.size(); k++)
                        if (g.node_res[k] == r) return k;
                    g.node_res.push_back(r);
                    return (uint32_t)g.node_res.size() - 1;
                };
                for (auto &e : ce) {
                    const uint32_t a = node_of(e.i);
                    const uint32_t b = node_of(e.j);
                    g.e.push_back({a, b});
                }
                graph_components(g, comps);
                for (auto &comp : comps) {
                    std::vector<uint32_t> sub;
                    for (uint32_t k = 0; k < ce.size(); k++) {
                        const bool ia = std::find(comp.begin(), comp.end(), g.e[k].first) != comp.end();
                        const bool ib = std::find(comp.begin(), comp.end(), g.e[k].second) != comp.end();
                        if (ia && ib) sub.push_back(k);
                    }
                    outv.emplace_back();
                    MatchTmp &m = outv.back();
                    m.cand = (uint32_t)c;
                    match_component(Q, ce, sub, all_pairs.data() + p_begin[c], p_begin[c + 1] - p_begin[c],
                                    (uint32_t)comp.size(), p->skip_ca_match != 0, m);
                }
            }
            chunks[tid].push_back(Chunk{c0, tid, begin, outv.size()});
        }
    };
    fd_parallel(nt, [&](int t) { worker(t); });
    std::vector<MatchTmp> mt;
    std::vector<Chunk> allc;
    for (auto &v : chunks) allc.insert(allc.end(), v.begin(), v.end());
    std::sort(allc.begin(), allc.end(), [](const Chunk &a, const Chunk &b) { return a.c0 < b.c0; });
    for (auto &ch : allc)
        for (size_t k = ch.begin; k < ch.end; k++) mt.push_back(std::move(per_thread[ch.thread][k]));
    *host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    const uint32_t n_align = (uint32_t)mt.size();
    if (!n_align) return FD_OK;
    std::vector<float> rmsd(n_align), U(9 * (size_t)n_align), T(3 * (size_t)n_align);
    std::vector<uint64_t> q_res_off(nq + 1, 0);
    for (uint32_t q = 0; q < nq; q++) q_res_off[q + 1] = q_res_off[q] + qs->q[q_begin + q].st->nres();
    std::vector<float> q_ca(3 * q_res_off[nq]), q_cb(3 * q_res_off[nq]);
    for (uint32_t q = 0; q < nq; q++) {
        const Query &Q = qs->q[q_begin + q];
        memcpy(q_ca.data() + 3 * q_res_off[q], Q.st->ca.data(), 12 * Q.st->nres());
        memcpy(q_cb.data() + 3 * q_res_off[q], Q.st->cb.data(), 12 * Q.st->nres());
    }
    std::vector<uint32_t> a_nid(n_align), a_off(n_align + 1, 0), pq_, pt_;
    for (uint32_t a = 0; a < n_align; a++) {
        const MatchTmp &m = mt[a];
        a_nid[a] = cand_n[m.cand];
        const uint64_t qb = q_res_off[cand_q[m.cand]];
        for (size_t k = 0; k < m.aq.size(); k++) {
            pq_.push_back((uint32_t)(qb + m.aq[k]));
            pt_.push_back(m.at[k]);
        }
        a_off[a + 1] = (uint32_t)pq_.size();
    }
    R->h2d_bytes += 4ull * (q_ca.size() + q_cb.size()) + 4ull * (a_nid.size() + a_off.size() + pq_.size() + pt_.size());
    R->d2h_bytes += 52ull * n_align;
    // `--partial-fit`: LMS-QCP above three matched residues (retrieve.rs:773-814), else Kabsch
    if ((p->partial_fit ? fd_lmsqcp_store_batch : fd_kabsch_store_batch)(ctx, q_ca.data(), q_cb.data(), q_res_off[nq], a_nid.data(),
                                                                         a_off.data(), n_align, pq_.data(), pt_.data(),
                                                                         rmsd.data(), U.data(), T.data()) != FD_OK) {
        set_err(fd_last_error(ctx));
        return FD_ERR_CUDA;
    }
    for (uint32_t a = 0; a < n_align; a++) {
        FinalMatch fm;
        fm.cand = cand_global[mt[a].cand];
        fm.node_count = mt[a].node_count_final;
        fm.idf = mt[a].idf;
        fm.rmsd = rmsd[a];
        memcpy(fm.U, &U[9 * (size_t)a], sizeof(fm.U));
        memcpy(fm.t, &T[3 * (size_t)a], sizeof(fm.t));
        fm.n_res = (uint32_t)mt[a].res_final.size();
        if (fm.n_res <= 16) memcpy(fm.res16, mt[a].res_final.data(), 4 * fm.n_res);
        else fm.res_big = mt[a].res_final;
        out.push_back(std::move(fm));
    }
    return FD_OK;
}

} // namespace

extern "C" {

// fd_query descriptors of queries [q_begin, q_end) (views into qs)
static void make_fd_queries(const fdh_queries *qs, uint32_t q_begin, uint32_t q_end, std::vector<fd_query> &fq,
                            uint64_t *h2d_bytes) {
    fq.resize(q_end - q_begin);
    for (uint32_t q = q_begin; q < q_end; q++) {
        const Query &Q = qs->q[q];
        if (!qs->shard_bounds.empty())
            fq[q - q_begin] = fd_query{(uint32_t)Q.hashes_flat.size(), Q.hashes_flat.data(), Q.s_edge_of_hash.data(),
                                       (uint32_t)Q.s_edge_node.size(), Q.s_edge_node.data(), Q.n_nodes,
                                       Q.residue_count, Q.s_edge_group.data()};
        else
            fq[q - q_begin] = fd_query{(uint32_t)Q.hashes_flat.size(), Q.hashes_flat.data(), Q.edge_of_hash.data(),
                                       (uint32_t)Q.edge_node.size(), Q.edge_node.data(), Q.n_nodes,
                                       Q.residue_count, nullptr};
        if (h2d_bytes) *h2d_bytes += 6ull * Q.hashes_flat.size() + 2ull * Q.edge_node.size() + 24;
    }
}

// verification tables of the whole batch (cand_query = index into qs->q), cached on qs
static int ensure_verify_prepared(fd_ctx *ctx, const fdh_queries *qs) {
    const int dev = fd_device(ctx);
    if (qs->vprep && qs->vprep_device == dev) return FD_OK;
    fd_verify_prepared_free(qs->vprep);
    qs->vprep = nullptr;
    const uint32_t nq = (uint32_t)qs->q.size();
    std::vector<fd_verify_query> vq(nq);
    for (uint32_t q = 0; q < nq; q++) {
        const Query &Q = qs->q[q];
        vq[q] = fd_verify_query{(uint32_t)Q.hashes_sorted.size(), Q.hashes_sorted.data(), Q.vs_qi.data(),
                                Q.vs_qj.data(), Q.vs_idf.data(), Q.vs_sym.data(), (uint32_t)Q.aad.size(),
                                Q.aad_aa1.data(), Q.aad_aa2.data(), Q.aad_dist.data(), Q.aad_qi.data(),
                                (uint32_t)Q.indices.size(), Q.indices.data(), (uint32_t)Q.st->nres(), Q.st->ca.data(),
                                Q.st->cb.data()};
    }
    const int rc = fd_verify_prepare(ctx, vq.data(), nq, &qs->vprep);
    if (rc != FD_OK) {
        set_err(fd_last_error(ctx));
        return rc;
    }
    qs->vprep_device = dev;
    return FD_OK;
}

// query_pdb.rs:348-452 for queries [q_begin, q_end) of the batch.  votes == nullptr: count_query on this
// context's index (fd_count_query_batch); otherwise finish count_query from merged dense votes (fd_votes_select).
static fdh_results *search_impl(fd_ctx *ctx, const fdh_queries *qs, const fdh_search_params *p, const fdh_store *labels,
                                uint32_t q_begin, uint32_t q_end, const fd_votes_layout *layout, const uint32_t *votes,
                                bool sharded = false) {
    if (!qs->finalized) {
        set_err("fdh_search: call fdh_queries_finalize first");
        return nullptr;
    }
    if (q_begin > q_end || q_end > qs->q.size()) {
        set_err("fdh_search: query range out of bounds");
        return nullptr;
    }
    const uint32_t nq = q_end - q_begin;
    fdh_results *R = new fdh_results();
    R->struct_off.assign(nq + 1, 0);
    R->match_off.assign(nq + 1, 0);
    if (nq == 0) return R;
    // --- K3: count_query + filter + sort + top ---
    std::vector<fd_query> fq;
    fd_struct_hit *hits = nullptr;
    uint64_t *hoff = nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    const auto t_all = now();
    auto t_stage = now();
    int rc;
    if (sharded) {
        // id-range shards: this rank scans its local index for the whole batch; the owners get the global top n
        const fdh_queries::ShardedBatch &S = qs->sh;
        const uint32_t nb = (uint32_t)S.q_nh.size();
        fq.resize(nb);
        size_t hb = 0, eb = 0;
        for (uint32_t q = 0; q < nb; q++) {
            fq[q] = fd_query{S.q_nh[q], S.hashes.data() + hb, S.edge_of_hash.data() + hb, S.q_ne[q], S.edge_node.data() + eb,
                             S.q_nn[q], S.q_expected[q], nullptr};
            hb += S.q_nh[q];
            eb += S.q_ne[q];
        }
        R->h2d_bytes += 6ull * hb + 2ull * eb + 24ull * nb + 4ull * hb;
        rc = fd_count_query_sharded_id(ctx, fq.data(), nb, &p->prefilter, S.gcounts.data(), S.total_structs, S.first_id,
                                       S.slice_begin.data(), qs->batch_id(), &hits, &hoff);
    } else if (votes) {
        make_fd_queries(qs, 0, (uint32_t)qs->q.size(), fq, nullptr);
        rc = fd_votes_select(ctx, fq.data(), (uint32_t)fq.size(), &p->prefilter, layout, votes, q_begin, q_end, &hits, &hoff);
    } else {
        make_fd_queries(qs, q_begin, q_end, fq, &R->h2d_bytes);
        // (a sub-range of the batch is a different batch: the id covers the whole one)
        rc = fd_count_query_batch_id(ctx, fq.data(), nq, &p->prefilter, q_begin == 0 && q_end == qs->q.size() ? qs->batch_id() : 0,
                                     &hits, &hoff);
    }
    if (rc != FD_OK) {
        set_err(fd_last_error(ctx));
        delete R;
        return nullptr;
    }
    const uint64_t n_cand = hoff[nq];
    R->wall_ms[0] = ms_since(t_stage);
    t_stage = now();
    R->d2h_bytes += n_cand * sizeof(fd_struct_hit) + (nq + 1) * 8ull + nq * 16ull;
    const bool any_match_filter = p->connected_node_count > 0 || p->connected_node_ratio > 0.f ||
                                  p->prefilter.idf_score_cutoff > 0.f || p->rmsd_cutoff > 0.f;
    const bool any_struct_filter = !p->skip_match && (p->max_matching_node_count > 0 || p->max_matching_node_ratio > 0.f ||
                                                      p->rmsd_cutoff > 0.f);
    std::vector<FinalMatch> fm; // matches of the candidates that took the general path, grouped by candidate
    // matches of everything else: views of pinned staging buffers, one set per verification lane (a lane = a
    // contiguous range of queries verified by one host thread on its own context / stream)
    struct Lane {
        uint64_t c0 = 0, c1 = 0; // candidate range
        const fd_match_record *recs = nullptr;
        const uint32_t *first = nullptr;
        const uint8_t *flags = nullptr;
        uint64_t n_recs = 0;
        int rc = FD_OK;
        std::string err;
    };
    std::vector<Lane> lanes;
    std::vector<uint8_t> all_general;
    std::vector<uint32_t> cand_q;
    double host_ms = 0.0;
    auto fail = [&]() {
        fd_free(hits);
        fd_free(hoff);
        delete R;
        return (fdh_results *)nullptr;
    };
    if (!p->skip_match && n_cand) {
        cand_q.resize(n_cand);
        std::vector<uint32_t> cand_n(n_cand);
        for (uint32_t q = 0; q < nq; q++)
            for (uint64_t k = hoff[q]; k < hoff[q + 1]; k++) {
                cand_q[k] = q;
                cand_n[k] = hits[k].nid;
            }
        // --- K6: verification on the device; candidates beyond its limits come back flagged ---
        // the fused kernels superpose with Kabsch only: a partial-fit search takes the general path (K4 -> host graph
        // step -> fd_lmsqcp_store_batch)
        const bool general_only = p->verify_mode == 1 || !fdg::ht_default_route(&qs->p.hash) || p->partial_fit != 0;
        if (!general_only) {
            if (ensure_verify_prepared(ctx, qs) != FD_OK) return fail();
            R->h2d_bytes += fd_verify_prepared_bytes(qs->vprep) * nq / std::max<size_t>(1, qs->q.size());
        }
        uint64_t n_recs = 0;
        // other encodings / `--multiple-bins`: the fused kernels and the pair table are PDBTrRosetta single-bin only
        if (general_only) { // general path for everything
            all_general.assign(n_cand, 1);
            lanes.resize(1);
            lanes[0].c1 = n_cand;
            lanes[0].flags = all_general.data();
        } else {
            // Optional lanes (FD_VERIFY_LANES=2..4): the queries are split into ranges verified by separate host
            // threads on forked contexts (own stream, own staging), so one lane's host-side steps, synchronisations
            // and record copy overlap another lane's kernels.  Measured on the bench workload: no gain (7.4 vs 7.2 ms,
            // the verification kernels already fill the GPU), so the default is one lane.
            int n_lanes = 1;
            if (const char *e = getenv("FD_VERIFY_LANES")) n_lanes = std::max(1, std::min(atoi(e), (int)std::min<uint32_t>(nq, 4)));
            std::vector<fd_ctx *> lane_ctx(n_lanes, ctx);
            for (int l = 1; l < n_lanes; l++)
                if (fd_lane(ctx, l - 1, &lane_ctx[l]) != FD_OK) {
                    set_err(fd_last_error(ctx));
                    return fail();
                }
            lanes.resize(n_lanes);
            std::vector<uint32_t> cand_q_global(n_cand); // index into the prepared tables of the whole batch
            for (uint64_t c = 0; c < n_cand; c++) cand_q_global[c] = cand_q[c] + q_begin;
            std::vector<uint32_t> q_split(n_lanes + 1);
            for (int l = 0; l <= n_lanes; l++) q_split[l] = (uint32_t)((uint64_t)nq * l / n_lanes);
            for (int l = 0; l < n_lanes; l++) {
                lanes[l].c0 = hoff[q_split[l]];
                lanes[l].c1 = hoff[q_split[l + 1]];
            }
            // Row assembly on the device (fd_verify_rows): the records stay in HBM and the finished row arrays are
            // copied straight into this result's page-locked blocks.  Taken with the default flags (no after-match
            // filter), one lane, labels on the device iff the caller passed them; a batch with candidates on the
            // general path falls back to the host assembly below.
            const bool device_rows = n_lanes == 1 && !any_match_filter && !any_struct_filter && !p->want_metrics &&
                                     (labels != nullptr) == (fd_store_has_labels(ctx) != 0) &&
                                     !(getenv("FD_DEVICE_ROWS") && atoi(getenv("FD_DEVICE_ROWS")) == 0);
            bool rows_done = false;
            auto run_lane = [&](int l) {
                fd_ctx *lc = lane_ctx[l];
                Lane &L = lanes[l];
                if (device_rows) {
                    // one pipelined call: chunks cut at query boundaries, the rows of a finished chunk are assembled
                    // and copied while the next chunks are verified
                    std::vector<uint32_t> n_res_q(nq);
                    uint32_t max_res = 1;
                    for (uint32_t q = 0; q < nq; q++) {
                        n_res_q[q] = (uint32_t)qs->q[q_begin + q].indices.size();
                        max_res = std::max(max_res, n_res_q[q]);
                    }
                    const uint64_t mcap = fd_verify_match_capacity(lc, n_cand);
                    std::vector<uint64_t> res_off(nq + 1, 0);
                    std::vector<uint8_t> needs(nq, 0);
                    if (!R->structs.alloc(n_cand) || !R->matches.alloc(mcap) || !R->match_order.alloc(mcap) ||
                        !R->residues.alloc(mcap * max_res)) {
                        L.rc = FD_ERR_NOMEM;
                        L.err = "fdh_search: host allocation failed";
                        return;
                    }
                    fd_rows_plan plan{nq, hoff, hits, n_res_q.data(), mcap, mcap * max_res, R->structs.data(), R->matches.data(),
                                      R->match_order.data(), R->residues.data(), needs.data(), res_off.data(), 0};
                    L.rc = fd_verify_candidates_rows(lc, qs->vprep, cand_q_global.data() + L.c0, cand_n.data() + L.c0,
                                                     L.c1 - L.c0, &qs->p.hash, p->ca_dist_cutoff, p->skip_ca_match, &plan,
                                                     &L.n_recs, &L.first, &L.flags);
                    if (L.rc == FD_OK) {
                        bool general = false;
                        for (uint64_t c = 0; c < n_cand && !general; c++) general = L.flags[c] != 0;
                        bool have_rows = plan.done != 0;
                        if (!have_rows && !general) {
                            // a capacity was too small: the records are still on the device, assemble in one piece
                            for (uint32_t q = 0; q < nq; q++)
                                res_off[q + 1] = res_off[q] + (uint64_t)(L.first[hoff[q + 1]] - L.first[hoff[q]]) * n_res_q[q];
                            if (!R->matches.alloc(L.n_recs) || !R->match_order.alloc(L.n_recs) || !R->residues.alloc(res_off[nq])) {
                                L.rc = FD_ERR_NOMEM;
                                L.err = "fdh_search: host allocation failed";
                                return;
                            }
                            fd_rows_request rq{nq, hoff, hits, res_off.data(), R->structs.data(), R->matches.data(),
                                               R->match_order.data(), R->residues.data(), needs.data()};
                            L.rc = fd_verify_rows(lc, &rq);
                            have_rows = L.rc == FD_OK;
                        }
                        if (have_rows) {
                            rows_done = true;
                            R->matches.shrink(L.n_recs);
                            R->match_order.shrink(L.n_recs);
                            R->residues.shrink(res_off[nq]);
                            for (uint32_t q = 0; q < nq; q++) {
                                R->struct_off[q + 1] = hoff[q + 1];
                                R->match_off[q + 1] = L.first[hoff[q + 1]];
                            }
                            for (uint32_t q = 0; q < nq; q++) { // queries the kernel left in emission order
                                if (!needs[q]) continue;
                                fdh_struct_row *S0 = R->structs.data() + R->struct_off[q], *S1 = R->structs.data() + R->struct_off[q + 1];
                                std::stable_sort(S0, S1, [](const fdh_struct_row &a, const fdh_struct_row &b) {
                                    if (a.idf != b.idf) return a.idf > b.idf;
                                    return a.min_rmsd_with_max_match < b.min_rmsd_with_max_match;
                                });
                                const fdh_match_row *M = R->matches.data();
                                uint64_t *O0 = R->match_order.data() + R->match_off[q], *O1 = R->match_order.data() + R->match_off[q + 1];
                                std::stable_sort(O0, O1, [&](uint64_t a, uint64_t b) {
                                    const fdh_match_row &x = M[a], &y = M[b];
                                    if (x.idf != y.idf) return x.idf > y.idf;
                                    return x.rmsd < y.rmsd;
                                });
                            }
                        } else if (L.rc == FD_OK) {
                            L.rc = fd_verify_records_fetch(lc, &L.recs);
                        }
                    }
                } else {
                    L.rc = fd_verify_candidates_prepared(lc, qs->vprep, cand_q_global.data() + L.c0, cand_n.data() + L.c0,
                                                         L.c1 - L.c0, &qs->p.hash, p->ca_dist_cutoff, p->skip_ca_match,
                                                         &L.recs, &L.n_recs, &L.first, &L.flags);
                }
                if (L.rc != FD_OK && L.err.empty()) L.err = fd_last_error(lc);
            };
            std::vector<std::thread> th;
            for (int l = 1; l < n_lanes; l++) th.emplace_back(run_lane, l);
            run_lane(0);
            for (auto &t : th) t.join();
            fd_lanes_fold_stats(ctx);
            for (auto &L : lanes) {
                if (L.rc != FD_OK) {
                    set_err(L.err);
                    return fail();
                }
                n_recs += L.n_recs;
            }
            if (rows_done) {
                R->h2d_bytes += 8ull * n_cand + n_cand * sizeof(fd_struct_hit) + 4ull * n_cand + 16ull * nq;
                R->d2h_bytes += n_cand * sizeof(fdh_struct_row) + n_recs * (sizeof(fdh_match_row) + 8) +
                                R->residues.size() * sizeof(fdh_residue_match) + 5ull * n_cand + nq;
                R->wall_ms[1] = ms_since(t_stage);
                R->wall_ms[2] = 0.0;
                R->wall_ms[3] = ms_since(t_all);
                R->host_ms = 0.0;
                fd_free(hits);
                fd_free(hoff);
                return R;
            }
        }
        R->h2d_bytes += 8ull * n_cand;
        R->d2h_bytes += n_recs * sizeof(fd_match_record) + 5ull * n_cand;
        auto t0 = std::chrono::steady_clock::now();
        std::vector<uint32_t> fq_, fn_;
        std::vector<uint64_t> fglobal;
        uint64_t general_reason[5] = {0, 0, 0, 0, 0};
        for (auto &L : lanes)
            for (uint64_t c = L.c0; c < L.c1; c++)
                if (L.flags[c - L.c0]) {
                    fq_.push_back(cand_q[c]);
                    fn_.push_back(cand_n[c]);
                    fglobal.push_back(c);
                    // why the candidate left the fused kernels (fd_verify.cu): 1 query beyond their limits, 2 more than
                    // 256 matching edges, 4 more than 64 graph nodes, 8 more than 16 components, 16 prefilter list cap
                    for (int b = 0; b < 5; b++)
                        if (L.flags[c - L.c0] & (1u << b)) general_reason[b]++;
                }
        host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fd_note_general_path(ctx, fglobal.size());
        {
            static const char *names[5] = {"general_reason_query", "general_reason_edges", "general_reason_nodes",
                                           "general_reason_components", "general_reason_lists"};
            for (int b = 0; b < 5; b++)
                for (uint64_t k = 0; k < general_reason[b]; k++) fd_note_host_ms(ctx, names[b], 0.0);
        }
        const auto t_general = std::chrono::steady_clock::now();
        if (!fglobal.empty()) {
            if (verify_general(ctx, qs, q_begin, nq, p, fq_, fn_, fglobal, fm, R, &host_ms) != FD_OK) return fail();
            std::stable_sort(fm.begin(), fm.end(), [](const FinalMatch &a, const FinalMatch &b) { return a.cand < b.cand; });
            fd_note_host_ms(ctx, "general_path_wall", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_general).count());
        }
    }
    // --- assemble rows: per-candidate summary (retrieve.rs:539-551), filter_after_matching (filter.rs:103-116),
    //     MatchFilter (:194-235), default sorts; query-parallel, then concatenated in query order ---
    R->wall_ms[1] = ms_since(t_stage);
    auto t1 = std::chrono::steady_clock::now();
    std::vector<size_t> fm_begin; // general-path matches of candidate c: fm[fm_begin[c] .. fm_begin[c+1])
    if (!fm.empty()) {
        fm_begin.assign(n_cand + 1, 0);
        for (auto &m : fm) fm_begin[m.cand + 1]++;
        for (uint64_t c = 0; c < n_cand; c++) fm_begin[c + 1] += fm_begin[c];
    }
    struct MatchView { // one verified component, from either source
        uint32_t node_count;
        float idf, rmsd;
        const float *U, *t;
        const uint32_t *res;
    };
    // number of matches of candidate c and the a-th of them
    auto lane_of = [&](uint64_t c) -> const Lane & {
        size_t l = 0;
        while (l + 1 < lanes.size() && c >= lanes[l].c1) l++;
        return lanes[l];
    };
    auto match_count = [&](uint64_t c) -> size_t {
        if (lanes.empty()) return 0;
        const Lane &L = lane_of(c);
        if (L.flags[c - L.c0]) return fm_begin.empty() ? 0 : fm_begin[c + 1] - fm_begin[c];
        return L.first ? (size_t)(L.first[c - L.c0 + 1] - L.first[c - L.c0]) : 0;
    };
    auto match_at = [&](uint64_t c, size_t a) -> MatchView {
        const Lane &L = lane_of(c);
        if (L.flags[c - L.c0]) {
            const FinalMatch &m = fm[fm_begin[c] + a];
            return MatchView{m.node_count, m.idf, m.rmsd, m.U, m.t, m.res()};
        }
        const fd_match_record &r = L.recs[L.first[c - L.c0] + a];
        return MatchView{r.node_count, r.idf, r.rmsd, r.U, r.t, r.res};
    };
    // per-candidate summary (retrieve.rs:539-551) + filter_after_matching (filter.rs:103-116)
    auto summarize = [&](uint64_t c, size_t na, float expected, uint32_t *max_node_out, float *min_rmsd_out) -> bool {
        uint32_t max_node = 0;
        float min_rmsd = 0.f;
        for (size_t a = 0; a < na; a++) {
            const MatchView m = match_at(c, a);
            if (m.node_count > max_node) {
                max_node = m.node_count;
                min_rmsd = m.rmsd;
            } else if (m.node_count == max_node && m.rmsd < min_rmsd) {
                min_rmsd = m.rmsd;
            }
        }
        *max_node_out = max_node;
        *min_rmsd_out = min_rmsd;
        if (p->skip_match) return true;
        bool pass = true;
        if (p->max_matching_node_count > 0) pass = pass && max_node >= p->max_matching_node_count;
        if (p->max_matching_node_ratio > 0.f) pass = pass && (float)max_node / expected >= p->max_matching_node_ratio;
        if (p->rmsd_cutoff > 0.f) pass = pass && min_rmsd <= p->rmsd_cutoff;
        return pass;
    };
    auto match_passes = [&](const MatchView &m, float expected) -> bool { // MatchFilter (filter.rs:194-235)
        bool pass = true;
        if (p->connected_node_count > 0) pass = pass && m.node_count >= p->connected_node_count;
        if (p->connected_node_ratio > 0.f) pass = pass && (float)m.node_count / expected >= p->connected_node_ratio;
        if (p->prefilter.idf_score_cutoff > 0.f) pass = pass && m.idf >= p->prefilter.idf_score_cutoff;
        if (p->rmsd_cutoff > 0.f) pass = pass && m.rmsd <= p->rmsd_cutoff;
        return pass;
    };
    std::vector<uint64_t> res_off(nq + 1, 0);
    // pass 1: rows per query
    auto count_query_rows = [&](uint32_t q) {
        const Query &Q = qs->q[q_begin + q];
        const float expected = (float)Q.residue_count;
        uint64_t ns = 0, nm = 0;
        if (!any_struct_filter && !any_match_filter) { // default flags: every candidate and every match is a row
            ns = hoff[q + 1] - hoff[q];
            for (uint64_t c = hoff[q]; c < hoff[q + 1]; c++) nm += match_count(c);
            R->struct_off[q + 1] = ns;
            R->match_off[q + 1] = nm;
            res_off[q + 1] = nm * Q.indices.size();
            return;
        }
        for (uint64_t c = hoff[q]; c < hoff[q + 1]; c++) {
            const size_t na = match_count(c);
            uint32_t max_node;
            float min_rmsd;
            if (!summarize(c, na, expected, &max_node, &min_rmsd)) continue;
            ns++;
            if (!any_match_filter) {
                nm += na;
            } else {
                for (size_t a = 0; a < na; a++) nm += match_passes(match_at(c, a), expected) ? 1 : 0;
            }
        }
        R->struct_off[q + 1] = ns;
        R->match_off[q + 1] = nm;
        res_off[q + 1] = nm * Q.indices.size();
    };
    // pass 2: rows written in place, then the default sorts inside the query's ranges
    const bool prof = getenv("FD_ASSEMBLE_PROFILE") != nullptr;
    std::atomic<uint64_t> prof_rows{0}, prof_ssort{0}, prof_msort{0};
    auto tick = [] { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    auto build_query = [&](uint32_t q) {
        const uint64_t pt0 = prof ? tick() : 0;
        const Query &Q = qs->q[q_begin + q];
        const float expected = (float)Q.residue_count;
        const uint32_t n_res = (uint32_t)Q.indices.size();
        fdh_struct_row *S0 = R->structs.data() + R->struct_off[q], *S = S0;
        const uint64_t mb = R->match_off[q];
        fdh_match_row *M = R->matches.data();
        fdh_residue_match *RM = R->residues.data();
        uint64_t mpos = mb, rpos = res_off[q];
        for (uint64_t c = hoff[q]; c < hoff[q + 1]; c++) {
            const size_t na = match_count(c);
            uint32_t max_node;
            float min_rmsd;
            if (!summarize(c, na, expected, &max_node, &min_rmsd)) continue;
            fdh_struct_row sr{hits[c].nid, hits[c].match_count, hits[c].node_count, hits[c].edge_count, hits[c].idf,
                              max_node, min_rmsd, 0, 0};
            sr.match_begin = mpos;
            for (size_t a = 0; a < na; a++) {
                const MatchView m = match_at(c, a);
                if (any_match_filter && !match_passes(m, expected)) continue;
                fdh_match_row &mr = M[mpos++];
                mr.nid = hits[c].nid;
                mr.node_count = m.node_count;
                mr.idf = m.idf;
                mr.rmsd = m.rmsd;
                memcpy(mr.U, m.U, sizeof(mr.U));
                memcpy(mr.t, m.t, sizeof(mr.t));
                mr.res_begin = rpos;
                const uint32_t *res = m.res;
                for (uint32_t k = 0; k < n_res; k++) {
                    const uint32_t v = res[k];
                    fdh_residue_match rm{(uint8_t)(v != 0), 0, v ? (uint64_t)(v - 1) : 0};
                    if (v && labels && hits[c].nid < labels->names.size()) { // (chain, residue number) of the target
                        const uint64_t r = labels->row_offsets[hits[c].nid] + (v - 1);
                        rm.chain = labels->chain[r];
                        rm.serial = labels->serial[r];
                    }
                    if (!R->res_index.empty()) R->res_index[rpos] = v;
                    RM[rpos++] = rm;
                }
            }
            sr.match_end = mpos;
            *S++ = sr;
        }
        const uint64_t pt1 = prof ? tick() : 0;
        // StructureSortStrategy::default: idf desc, min_rmsd asc (sort.rs:454-458), stable.  The rows arrive in
        // count_query order (idf desc, nid asc), so only runs of equal idf can need reordering.
        {
            bool sorted = true;
            for (fdh_struct_row *a = S0; a + 1 < S && sorted; a++) sorted = a[0].idf > a[1].idf;
            if (!sorted)
                std::stable_sort(S0, S, [](const fdh_struct_row &a, const fdh_struct_row &b) {
                    if (a.idf != b.idf) return a.idf > b.idf;
                    return a.min_rmsd_with_max_match < b.min_rmsd_with_max_match;
                });
        }
        // MatchSortStrategy::default: idf desc, rmsd asc (sort.rs:218-222), stable over emission order.  Sorted as
        // (order-preserving integer image of (-idf, rmsd), emission index) records: the same total order as the
        // stable comparison sort, without the two indirect float loads per comparison.  Values that the integer image
        // would order differently from the float comparison (NaN, -0.0) take the comparison sort.
        const uint64_t pt2 = prof ? tick() : 0;
        uint64_t *O = R->match_order.data() + mb;
        const uint64_t nm = mpos - mb;
        auto sortable = [](float f) -> uint32_t {
            uint32_t u;
            memcpy(&u, &f, 4);
            return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        };
        bool plain = true;
        for (uint64_t k = mb; k < mpos && plain; k++)
            plain = M[k].idf == M[k].idf && M[k].rmsd == M[k].rmsd && !(M[k].idf == 0.f && std::signbit(M[k].idf)) &&
                    !(M[k].rmsd == 0.f && std::signbit(M[k].rmsd));
        if (plain && nm <= 4096) {
            struct Key {
                uint64_t key;
                uint32_t idx;
            };
            Key local[256];
            std::vector<Key> big;
            Key *keys = local;
            if (nm > 256) {
                big.resize(nm);
                keys = big.data();
            }
            for (uint64_t k = 0; k < nm; k++)
                keys[k] = Key{((uint64_t)(~sortable(M[mb + k].idf)) << 32) | sortable(M[mb + k].rmsd), (uint32_t)k};
            std::sort(keys, keys + nm, [](const Key &a, const Key &b) { return a.key != b.key ? a.key < b.key : a.idx < b.idx; });
            for (uint64_t k = 0; k < nm; k++) O[k] = mb + keys[k].idx;
        } else {
            for (uint64_t k = 0; k < nm; k++) O[k] = mb + k;
            std::stable_sort(O, O + nm, [&](uint64_t a, uint64_t b) {
                const fdh_match_row &x = M[a], &y = M[b];
                if (x.idf != y.idf) return x.idf > y.idf;
                return x.rmsd < y.rmsd;
            });
        }
        if (prof) {
            const uint64_t pt3 = tick();
            prof_rows += pt1 - pt0;
            prof_ssort += pt2 - pt1;
            prof_msort += pt3 - pt2;
        }
    };
    {
        int nt = p->host_threads > 0 ? p->host_threads : fd_default_host_threads();
        nt = std::max(1, std::min(nt, 64));
        auto run_parallel = [&](const std::function<void(uint32_t)> &fn) {
            std::atomic<uint32_t> next{0};
            auto worker = [&] {
                for (uint32_t q0; (q0 = next.fetch_add(8)) < nq;)
                    for (uint32_t q = q0; q < std::min(nq, q0 + 8); q++) fn(q);
            };
            fd_parallel(nt, [&](int) { worker(); });
        };
        run_parallel(count_query_rows);
        for (uint32_t q = 0; q < nq; q++) {
            R->struct_off[q + 1] += R->struct_off[q];
            R->match_off[q + 1] += R->match_off[q];
            res_off[q + 1] += res_off[q];
        }
        if (!R->structs.alloc(R->struct_off[nq]) || !R->matches.alloc(R->match_off[nq]) ||
            !R->match_order.alloc(R->match_off[nq]) || !R->residues.alloc(res_off[nq])) {
            set_err("fdh_search: host allocation failed");
            return fail();
        }
        if (p->want_metrics) R->res_index.assign(res_off[nq], 0);
        const auto tb0 = std::chrono::steady_clock::now();
        run_parallel(build_query);
        // StructureSimilarityMetrics of every match row (retrieve.rs:776-831) on the device, from the rows' own
        // superpositions and matched residues
        if (p->want_metrics && R->match_off[nq] > 0) {
            const uint64_t NM = R->match_off[nq];
            std::vector<uint64_t> q_res_off(nq + 1, 0);
            for (uint32_t q = 0; q < nq; q++) q_res_off[q + 1] = q_res_off[q] + qs->q[q_begin + q].st->nres();
            std::vector<float> q_ca(3 * q_res_off[nq]), q_cb(3 * q_res_off[nq]);
            for (uint32_t q = 0; q < nq; q++) {
                const Query &Q = qs->q[q_begin + q];
                memcpy(q_ca.data() + 3 * q_res_off[q], Q.st->ca.data(), 12 * Q.st->nres());
                memcpy(q_cb.data() + 3 * q_res_off[q], Q.st->cb.data(), 12 * Q.st->nres());
            }
            std::vector<uint32_t> a_nid(NM), a_off(NM + 1, 0), pq_, pt_;
            std::vector<float> U(9 * NM), T(3 * NM);
            for (uint32_t q = 0; q < nq; q++) {
                const Query &Q = qs->q[q_begin + q];
                for (uint64_t m = R->match_off[q]; m < R->match_off[q + 1]; m++) {
                    const fdh_match_row &mr = R->matches[m];
                    a_nid[m] = mr.nid;
                    memcpy(&U[9 * m], mr.U, sizeof(mr.U));
                    memcpy(&T[3 * m], mr.t, sizeof(mr.t));
                    for (size_t k = 0; k < Q.indices.size(); k++) {
                        const uint32_t v = R->res_index[mr.res_begin + k];
                        if (!v) continue;
                        pq_.push_back((uint32_t)(q_res_off[q] + Q.indices[k]));
                        pt_.push_back(v - 1);
                    }
                    a_off[m + 1] = (uint32_t)pq_.size();
                }
            }
            R->metrics.assign(5 * NM, 0.f);
            R->h2d_bytes += 4ull * (q_ca.size() + q_cb.size() + a_nid.size() + a_off.size() + pq_.size() + pt_.size() + U.size() + T.size());
            R->d2h_bytes += 20ull * NM;
            if (fd_metrics_store_batch(ctx, q_ca.data(), q_cb.data(), q_res_off[nq], a_nid.data(), a_off.data(), (uint32_t)NM,
                                       pq_.data(), pt_.data(), U.data(), T.data(), R->metrics.data()) != FD_OK) {
                set_err(fd_last_error(ctx));
                return fail();
            }
        }
        if (prof)
            fprintf(stderr, "assemble: pass1+alloc %.3f ms, pass2 wall %.3f ms (%d threads); cpu: rows %.3f ms, struct sort %.3f ms, match sort %.3f ms\n",
                    std::chrono::duration<double, std::milli>(tb0 - t1).count(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count(), nt,
                    prof_rows.load() * 1e-6, prof_ssort.load() * 1e-6, prof_msort.load() * 1e-6);
    }
    host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    R->host_ms = host_ms;
    R->wall_ms[2] = ms_since(t1);
    R->wall_ms[3] = ms_since(t_all);
    fd_free(hits);
    fd_free(hoff);
    return R;
}

fdh_results *fdh_search(fd_ctx *ctx, const fdh_queries *qs, const fdh_search_params *p, const fdh_store *labels) {
    return search_impl(ctx, qs, p, labels, 0, (uint32_t)qs->q.size(), nullptr, nullptr);
}

// ---- hash-range sharded index: the same search in three steps around the caller's collective ----
// The index is split into `world` hash ranges [bounds[r], bounds[r+1]).  count_query's edge bit masks must merge
// by addition, so every (query edge, rank owning at least one of the edge's hashes) pair gets its own vote bit;
// bits of one edge are consecutive, never cross a 32-bit word and are counted once (fd_query.edge_group).
int fdh_queries_set_shards(fdh_queries *qs, const uint64_t *bounds, int world) {
    if (world < 1 || world > 64 || !bounds) {
        set_err("fdh_queries_set_shards: world must be in 1..64");
        return FD_ERR_ARG;
    }
    for (int r = 0; r < world; r++)
        if (bounds[r] > bounds[r + 1]) {
            set_err("fdh_queries_set_shards: bounds must be ascending");
            return FD_ERR_ARG;
        }
    std::vector<uint64_t> b(bounds, bounds + world + 1);
    qs->generation++;
    for (auto &Q : qs->q) {
        const size_t E = Q.edge_node.size();
        std::vector<uint64_t> owners(E, 0);
        std::vector<uint8_t> rank_of(Q.hashes_flat.size());
        for (size_t k = 0; k < Q.hashes_flat.size(); k++) {
            const uint64_t h = Q.hashes_flat[k];
            int r = (int)(std::upper_bound(b.begin() + 1, b.end() - 1, h) - (b.begin() + 1)); // ranges cover everything
            rank_of[k] = (uint8_t)r;
            owners[Q.edge_of_hash[k]] |= 1ull << r;
        }
        Q.s_edge_node.clear();
        Q.s_edge_group.clear();
        std::vector<uint32_t> first_bit(E, 0);
        uint16_t gid = 0;
        for (size_t e = 0; e < E; e++) {
            const uint32_t g = std::max(1, __builtin_popcountll(owners[e]));
            while ((Q.s_edge_node.size() & 31) + g > 32) { // padding bit: never voted for, its own group
                Q.s_edge_node.push_back(Q.edge_node[e]);
                Q.s_edge_group.push_back(gid++);
            }
            first_bit[e] = (uint32_t)Q.s_edge_node.size();
            for (uint32_t k = 0; k < g; k++) {
                Q.s_edge_node.push_back(Q.edge_node[e]);
                Q.s_edge_group.push_back(gid);
            }
            gid++;
        }
        if (Q.s_edge_node.size() > 256) {
            set_err("fdh_queries_set_shards: query needs more than 256 vote bits");
            return FD_ERR_LIMIT;
        }
        Q.s_edge_of_hash.resize(Q.hashes_flat.size());
        for (size_t k = 0; k < Q.hashes_flat.size(); k++) {
            const uint32_t e = Q.edge_of_hash[k];
            const uint64_t below = owners[e] & ((1ull << rank_of[k]) - 1);
            Q.s_edge_of_hash[k] = (uint16_t)(first_bit[e] + __builtin_popcountll(below));
        }
    }
    qs->shard_bounds = std::move(b);
    return FD_OK;
}
int64_t fdh_queries_num_vote_bits(const fdh_queries *qs, int64_t q) {
    return (int64_t)(qs->shard_bounds.empty() ? qs->q[q].edge_node.size() : qs->q[q].s_edge_node.size());
}
void fdh_queries_get_vote_bits(const fdh_queries *qs, int64_t q, uint32_t *hashes, uint16_t *bit_of_hash,
                               uint16_t *bit_node, uint16_t *bit_group) {
    const Query &Q = qs->q[q];
    const bool sh = !qs->shard_bounds.empty();
    const std::vector<uint16_t> &eoh = sh ? Q.s_edge_of_hash : Q.edge_of_hash;
    const std::vector<uint16_t> &en = sh ? Q.s_edge_node : Q.edge_node;
    for (size_t k = 0; k < Q.hashes_flat.size(); k++) {
        if (hashes) hashes[k] = Q.hashes_flat[k];
        if (bit_of_hash) bit_of_hash[k] = eoh[k];
    }
    for (size_t e = 0; e < en.size(); e++) {
        if (bit_node) bit_node[e] = en[e];
        if (bit_group) bit_group[e] = sh ? Q.s_edge_group[e] : (uint16_t)e;
    }
}
// Flat scan inputs of the batch for an all-gather between ranks (each rank builds the query maps of its own
// slice only): per query the number of hashes, vote bits and pairs; then the concatenated arrays.
void fdh_queries_scan_sizes(const fdh_queries *qs, uint64_t *n_hashes, uint64_t *n_pairs) {
    uint64_t h = 0, pr = 0;
    for (auto &Q : qs->q) {
        h += Q.hashes_flat.size();
        pr += Q.pair_hash.size();
    }
    *n_hashes = h;
    *n_pairs = pr;
}
void fdh_queries_scan_arrays(const fdh_queries *qs, uint32_t *per_query /* 3 per query: hashes, bits, pairs */,
                             uint32_t *hashes, uint32_t *bit_of_hash, uint32_t *pair_hashes) {
    const bool sh = !qs->shard_bounds.empty();
    size_t h = 0, pr = 0;
    for (size_t q = 0; q < qs->q.size(); q++) {
        const Query &Q = qs->q[q];
        const std::vector<uint16_t> &eoh = sh ? Q.s_edge_of_hash : Q.edge_of_hash;
        per_query[3 * q] = (uint32_t)Q.hashes_flat.size();
        per_query[3 * q + 1] = (uint32_t)(sh ? Q.s_edge_node.size() : Q.edge_node.size());
        per_query[3 * q + 2] = (uint32_t)Q.pair_hash.size();
        for (size_t k = 0; k < Q.hashes_flat.size(); k++) {
            hashes[h] = Q.hashes_flat[k];
            bit_of_hash[h++] = eoh[k];
        }
        for (uint32_t v : Q.pair_hash) pair_hashes[pr++] = v;
    }
}
// fd_votes_scan_sparse over flat arrays (the whole batch, gathered from all ranks)
int fdh_votes_scan_sparse_flat(fd_ctx *ctx, uint32_t nq, const uint32_t *per_query, const uint32_t *hashes,
                               const uint32_t *bit_of_hash, const fd_prefilter_params *prefilter,
                               const uint32_t *slice_begin, uint32_t world, fd_votes_layout *layout,
                               uint32_t **d_records, uint64_t *region_offset, uint64_t *region_count) {
    std::vector<fd_query> fq(nq);
    size_t total = 0, max_bits = 1;
    for (uint32_t q = 0; q < nq; q++) {
        total += per_query[3 * q];
        max_bits = std::max<size_t>(max_bits, per_query[3 * q + 1]);
    }
    std::vector<uint16_t> eoh(total), zeros(max_bits, 0);
    for (size_t k = 0; k < total; k++) eoh[k] = (uint16_t)bit_of_hash[k];
    size_t h = 0;
    for (uint32_t q = 0; q < nq; q++) {
        // the pack epilogue does not look at nodes / groups: one node, no groups
        fq[q] = fd_query{per_query[3 * q], hashes + h, eoh.data() + h, per_query[3 * q + 1], zeros.data(), 1, 1, nullptr};
        h += per_query[3 * q];
    }
    const int rc = fd_votes_scan_sparse(ctx, fq.data(), nq, prefilter, slice_begin, world, layout, d_records,
                                        region_offset, region_count);
    if (rc != FD_OK) set_err(fd_last_error(ctx));
    return rc;
}
int64_t fdh_queries_num_pairs(const fdh_queries *qs) {
    int64_t n = 0;
    for (auto &Q : qs->q) n += (int64_t)Q.pair_hash.size();
    return n;
}
int fdh_queries_pair_counts(const fdh_queries *qs, fd_ctx *ctx, uint32_t *out_counts) {
    std::vector<uint32_t> all;
    for (auto &Q : qs->q) all.insert(all.end(), Q.pair_hash.begin(), Q.pair_hash.end());
    if (all.empty()) return FD_OK;
    const int rc = fd_posting_counts(ctx, all.data(), all.size(), out_counts);
    if (rc != FD_OK) set_err(fd_last_error(ctx));
    return rc;
}
int fdh_queries_finalize_with_counts(fdh_queries *qs, const uint32_t *counts, uint64_t total_structures) {
    fd_verify_prepared_free(qs->vprep); // the tables carry the per-hash idf
    qs->vprep = nullptr;
    const float total = (float)total_structures;
    const size_t nq = qs->q.size();
    std::vector<size_t> base(nq + 1, 0);
    for (size_t q = 0; q < nq; q++) base[q + 1] = base[q] + qs->q[q].pair_hash.size();
    std::atomic<size_t> next{0};
    fd_parallel(nq >= 64 ? std::min(fd_default_host_threads(), 16) : 1, [&](int) {
        for (size_t q0; (q0 = next.fetch_add(16)) < nq;)
            for (size_t q = q0; q < std::min(nq, q0 + 16); q++) {
                Query &Q = qs->q[q];
                for (auto &e : Q.entries) {
                    const uint32_t c = counts[base[q] + e.pair];
                    e.idf = c > 0 ? log2f(total / (float)c) : 0.0f;
                }
                for (size_t k = 0; k < Q.vs_entry.size(); k++) Q.vs_idf[k] = Q.entries[Q.vs_entry[k]].idf;
            }
    });
    qs->finalized = true;
    qs->generation++;
    return FD_OK;
}
int fdh_votes_scan(fd_ctx *ctx, const fdh_queries *qs, const fd_prefilter_params *prefilter, fd_votes_layout *layout,
                   uint32_t **d_votes) {
    std::vector<fd_query> fq;
    make_fd_queries(qs, 0, (uint32_t)qs->q.size(), fq, nullptr);
    const int rc = fd_votes_scan(ctx, fq.data(), (uint32_t)fq.size(), prefilter, layout, d_votes);
    if (rc != FD_OK) set_err(fd_last_error(ctx));
    return rc;
}
fdh_results *fdh_search_from_votes(fd_ctx *ctx, const fdh_queries *qs, const fdh_search_params *p,
                                   const fdh_store *labels, const fd_votes_layout *layout, const uint32_t *d_votes,
                                   uint32_t q_begin, uint32_t q_end) {
    if (!layout || !d_votes) {
        set_err("fdh_search_from_votes: NULL votes");
        return nullptr;
    }
    return search_impl(ctx, qs, p, labels, q_begin, q_end, layout, d_votes);
}

// ---- id-range shards over NCCL (fd_comm_*): gather the ranks' query descriptors, all-reduce the list lengths ----
int fdh_queries_finalize_sharded(fdh_queries *qs, fd_ctx *ctx, uint64_t first_id, uint64_t total_structs) {
    qs->generation++;
    const int world = fd_comm_world(ctx), rank = fd_comm_rank(ctx);
    auto t_mark = std::chrono::steady_clock::now();
    auto mark = [&](const char *name) { // host wall time of the step that just ended -> fd_stage_ms(ctx, name)
        const auto t = std::chrono::steady_clock::now();
        fd_note_host_ms(ctx, name, std::chrono::duration<double, std::milli>(t - t_mark).count());
        t_mark = t;
    };
    // this rank's blob: n_queries | per query {n_hashes, n_edges, n_nodes, residue_count, n_pairs} | hashes | pair
    // hashes | edge_of_hash | edge_node (u16 arrays last, padded to 4 bytes)
    std::vector<uint32_t> blob;
    const uint32_t nq = (uint32_t)qs->q.size();
    blob.push_back(nq);
    for (auto &Q : qs->q) {
        blob.push_back((uint32_t)Q.hashes_flat.size());
        blob.push_back((uint32_t)Q.edge_node.size());
        blob.push_back(Q.n_nodes);
        blob.push_back(Q.residue_count);
        blob.push_back((uint32_t)Q.pair_hash.size());
    }
    for (auto &Q : qs->q) blob.insert(blob.end(), Q.hashes_flat.begin(), Q.hashes_flat.end());
    for (auto &Q : qs->q) blob.insert(blob.end(), Q.pair_hash.begin(), Q.pair_hash.end());
    std::vector<uint16_t> u16;
    for (auto &Q : qs->q) u16.insert(u16.end(), Q.edge_of_hash.begin(), Q.edge_of_hash.end());
    for (auto &Q : qs->q) u16.insert(u16.end(), Q.edge_node.begin(), Q.edge_node.end());
    if (u16.size() & 1) u16.push_back(0);
    const size_t words32 = blob.size();
    blob.resize(words32 + u16.size() / 2);
    memcpy(blob.data() + words32, u16.data(), u16.size() * 2);
    mark("fs_flatten");
    uint64_t my_bytes = blob.size() * 4;
    std::vector<uint64_t> sizes(world);
    if (fd_comm_allgather(ctx, &my_bytes, 8, sizes.data()) != FD_OK) {
        set_err(fd_last_error(ctx));
        return FD_ERR_CUDA;
    }
    uint64_t mx = 0;
    for (uint64_t b : sizes) mx = std::max(mx, b);
    blob.resize(mx / 4, 0);
    std::vector<uint32_t> all((size_t)world * (mx / 4));
    if (fd_comm_allgather(ctx, blob.data(), mx, all.data()) != FD_OK) {
        set_err(fd_last_error(ctx));
        return FD_ERR_CUDA;
    }
    mark("fs_allgather");
    fdh_queries::ShardedBatch &S = qs->sh;
    S = fdh_queries::ShardedBatch();
    S.world = world;
    S.rank = rank;
    S.first_id = first_id;
    S.total_structs = total_structs;
    S.slice_begin.assign(world + 1, 0);
    std::vector<uint32_t> pair_all;       // pair hashes of every rank, rank by rank
    std::vector<size_t> pair_begin(world + 1, 0);
    for (int r = 0; r < world; r++) {
        const uint32_t *b = all.data() + (size_t)r * (mx / 4);
        const uint32_t n = b[0];
        S.slice_begin[r + 1] = S.slice_begin[r] + n;
        size_t nh = 0, ne = 0, np = 0;
        for (uint32_t q = 0; q < n; q++) {
            S.q_nh.push_back(b[1 + 5 * q]);
            S.q_ne.push_back(b[2 + 5 * q]);
            S.q_nn.push_back(b[3 + 5 * q]);
            S.q_expected.push_back(b[4 + 5 * q]);
            nh += b[1 + 5 * q];
            ne += b[2 + 5 * q];
            np += b[5 + 5 * q];
        }
        const uint32_t *hp = b + 1 + 5 * (size_t)n;
        S.hashes.insert(S.hashes.end(), hp, hp + nh);
        pair_all.insert(pair_all.end(), hp + nh, hp + nh + np);
        pair_begin[r + 1] = pair_all.size();
        const uint16_t *up = reinterpret_cast<const uint16_t *>(hp + nh + np);
        S.edge_of_hash.insert(S.edge_of_hash.end(), up, up + nh);
        S.edge_node.insert(S.edge_node.end(), up + nh, up + nh + ne);
    }
    // global list lengths: local counts of [batch hashes | pair hashes], summed over the ranks
    std::vector<uint32_t> probe(S.hashes);
    probe.insert(probe.end(), pair_all.begin(), pair_all.end());
    std::vector<uint32_t> counts(probe.size());
    mark("fs_parse");
    if (!probe.empty()) {
        if (fd_posting_counts(ctx, probe.data(), probe.size(), counts.data()) != FD_OK) {
            set_err(fd_last_error(ctx));
            return FD_ERR_CUDA;
        }
        mark("fs_counts");
        if (fd_comm_allreduce_u32(ctx, counts.data(), counts.size()) != FD_OK) {
            set_err(fd_last_error(ctx));
            return FD_ERR_CUDA;
        }
        mark("fs_allreduce");
    }
    S.gcounts.assign(counts.begin(), counts.begin() + S.hashes.size());
    // calculate_idf_for_hash (query.rs:17-32) of this rank's own query pairs, from the global lengths
    const int rc = fdh_queries_finalize_with_counts(qs, counts.data() + S.hashes.size() + pair_begin[rank], total_structs);
    if (rc != FD_OK) return rc;
    const int rc2 = ensure_verify_prepared(ctx, qs);
    mark("fs_tables");
    return rc2;
}

fdh_results *fdh_search_sharded(fd_ctx *ctx, const fdh_queries *qs, const fdh_search_params *p, const fdh_store *labels) {
    if (qs->sh.world == 0 || qs->sh.slice_begin.empty()) {
        set_err("fdh_search_sharded: call fdh_queries_finalize_sharded first");
        return nullptr;
    }
    if (qs->sh.slice_begin[qs->sh.rank + 1] - qs->sh.slice_begin[qs->sh.rank] != qs->q.size()) {
        set_err("fdh_search_sharded: the batch changed since fdh_queries_finalize_sharded");
        return nullptr;
    }
    return search_impl(ctx, qs, p, labels, 0, (uint32_t)qs->q.size(), nullptr, nullptr, true);
}

uint64_t fdh_results_num_queries(const fdh_results *r) { return r->struct_off.size() - 1; }
const uint64_t *fdh_results_struct_offsets(const fdh_results *r) { return r->struct_off.data(); }
const fdh_struct_row *fdh_results_struct_rows(const fdh_results *r) { return r->structs.data(); }
const uint64_t *fdh_results_match_offsets(const fdh_results *r) { return r->match_off.data(); }
const fdh_match_row *fdh_results_match_rows(const fdh_results *r) { return r->matches.data(); }
const uint64_t *fdh_results_match_order(const fdh_results *r) { return r->match_order.data(); }
const fdh_residue_match *fdh_results_residues(const fdh_results *r) { return r->residues.data(); }
uint64_t fdh_results_num_residues(const fdh_results *r) { return r->residues.size(); }
const float *fdh_results_metrics(const fdh_results *r) { return r->metrics.empty() ? nullptr : r->metrics.data(); }
const uint32_t *fdh_results_residue_index(const fdh_results *r) { return r->res_index.empty() ? nullptr : r->res_index.data(); }
int fdh_store_get_ca(const fdh_store *s, uint64_t id, uint64_t residue, float *xyz) {
    if (id >= s->names.size() || residue >= s->row_offsets[id + 1] - s->row_offsets[id]) return FD_ERR_ARG;
    memcpy(xyz, &s->ca[3 * (s->row_offsets[id] + residue)], 12);
    return FD_OK;
}
double fdh_results_host_ms(const fdh_results *r) { return r->host_ms; }
uint64_t fdh_results_h2d_bytes(const fdh_results *r) { return r->h2d_bytes; }
double fdh_results_wall_ms(const fdh_results *r, int which) { return which >= 0 && which < 4 ? r->wall_ms[which] : -1.0; }
uint64_t fdh_results_d2h_bytes(const fdh_results *r) { return r->d2h_bytes; }
void fdh_results_free(fdh_results *r) { delete r; }

} // extern "C"
