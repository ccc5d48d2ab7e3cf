// fd_cli.cpp -- `folddisco-b200 index | query`: the reference's two hot-path sub-commands on top of the host C ABI
// (include/folddisco_b200_host.h), with the reference's flag surface and TSV outputs so that results diff against a
// `folddisco` run elsewhere.
//
//   flags and defaults      src/cli/main.rs:26-141
//   index workflow          src/cli/workflows/build_index.rs (files in sorted path order; the reference uses read_dir
//                           order, loader.rs:12, which is unspecified -- sorted is what its README outputs show)
//   query workflow          src/cli/workflows/query_pdb.rs:186-512 (query file = lines of path \t residues \t output)
//   TSV columns, precision  src/controller/result.rs:213-353, src/utils/formatter.rs:7, 117-183
//   ids                     src/controller/mode.rs:70-125
//
// Not here (fails loudly): `benchmark` / `analyze`, Foldcomp databases.  There is no CPU path: without a CUDA device fd_create fails and the command exits non-zero.
#include <dirent.h>
#include <limits.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/folddisco_b200_host.h"

namespace {

[[noreturn]] void die(const std::string &msg) {
    fprintf(stderr, "[FAIL] %s\n", msg.c_str());
    exit(1);
}

struct Args { // pico_args-like: options may appear anywhere after the sub-command
    std::vector<std::string> v;
    std::vector<bool> used;
    Args(int argc, char **argv) {
        for (int i = 2; i < argc; i++) v.push_back(argv[i]);
        used.assign(v.size(), false);
    }
    bool flag(std::initializer_list<const char *> names) {
        bool found = false;
        for (size_t i = 0; i < v.size(); i++)
            for (const char *n : names)
                if (!used[i] && v[i] == n) used[i] = found = true;
        return found;
    }
    bool value(std::initializer_list<const char *> names, std::string &out) {
        for (size_t i = 0; i < v.size(); i++)
            for (const char *n : names) {
                if (used[i]) continue;
                const std::string key = n;
                if (v[i] == key) {
                    if (i + 1 >= v.size()) die("option " + key + " needs a value");
                    used[i] = used[i + 1] = true;
                    out = v[i + 1];
                    return true;
                }
                if (v[i].compare(0, key.size() + 1, key + "=") == 0) {
                    used[i] = true;
                    out = v[i].substr(key.size() + 1);
                    return true;
                }
            }
        return false;
    }
    std::string str(std::initializer_list<const char *> names, const std::string &dflt) {
        std::string s;
        return value(names, s) ? s : dflt;
    }
    double num(std::initializer_list<const char *> names, double dflt) {
        std::string s;
        if (!value(names, s)) return dflt;
        char *end = nullptr;
        const double x = strtod(s.c_str(), &end);
        if (end == s.c_str() || *end) die("invalid number '" + s + "'");
        return x;
    }
    void reject(std::initializer_list<const char *> names, const char *why) {
        for (const char *n : names)
            for (size_t i = 0; i < v.size(); i++)
                if (v[i] == n || v[i].compare(0, strlen(n) + 1, std::string(n) + "=") == 0)
                    die(std::string(n) + " is not supported by folddisco-b200: " + why);
    }
    void finish() {
        for (size_t i = 0; i < v.size(); i++)
            if (!used[i]) die("unknown or repeated argument: " + v[i]);
    }
};

bool ends_with(const std::string &s, const char *suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

bool is_dir(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
bool is_file(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

void list_files(const std::string &dir, bool recursive, std::vector<std::string> &out) { // loader.rs:9-37
    DIR *d = opendir(dir.c_str());
    if (!d) die("cannot read directory " + dir);
    std::vector<std::string> names;
    while (dirent *e = readdir(d)) {
        const std::string n = e->d_name;
        if (n == "." || n == "..") continue;
        names.push_back(n);
    }
    closedir(d);
    std::sort(names.begin(), names.end());
    const std::string base = !dir.empty() && dir.back() == '/' ? dir : dir + "/";
    for (auto &n : names) {
        const std::string p = base + n;
        if (is_dir(p)) {
            if (recursive) list_files(p, true, out);
        } else {
            out.push_back(p);
        }
    }
}

std::string make_id(const std::string &path, const std::string &id_type) { // mode.rs:19-31, 70-125
    std::string out((size_t)fdh_parse_path_by_id_type(path.c_str(), id_type.c_str(), nullptr, 0), '\0');
    std::vector<char> buf(out.size() + 1);
    fdh_parse_path_by_id_type(path.c_str(), id_type.c_str(), buf.data(), buf.size());
    return std::string(buf.data());
}

std::vector<float> parse_thresholds(const std::string &s) { // comma-separated list (query_pdb.rs parse_threshold_string)
    std::vector<float> out;
    size_t a = 0;
    while (a <= s.size()) {
        const size_t b = s.find(',', a);
        const std::string tok = s.substr(a, b == std::string::npos ? std::string::npos : b - a);
        if (!tok.empty()) {
            char *end = nullptr;
            const float x = strtof(tok.c_str(), &end);
            if (end == tok.c_str()) die("invalid threshold '" + tok + "'");
            out.push_back(x);
        }
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return out;
}

const char *HELP =
    "usage: folddisco-b200 <command> [<args>]\n\n"
    "subcommands:\n"
    "  index     Create a new index table from a directory of PDB files (GPU)\n"
    "  query     Query a motif from an index table (GPU)\n"
    "  version   Print version information\n\n"
    "index:  -p/--pdbs DIR|FOLDCOMP_DB  -i/--index PREFIX  [-t N] [-d NBIN_DIST] [-a NBIN_ANGLE] [-g GRID] [-n MAX_RESIDUE]\n"
    "        [-y/--type default|pdbtr|pdb|orig_pdb|tr|ppf|3di|hybrid|angle|dist] [--multiple-bins D1-A1,D2-A2,..]\n"
    "        [-r] [--id relpath|abspath|basename|filename|pdb] [--no-store] [-v]\n"
    "query:  -p/--pdb FILE -q/--query RESIDUES | -q FILE.txt|.tsv   -i/--index PREFIX  [-t N]\n"
    "        [-d DIST_THR[,..]] [-a ANGLE_THR[,..]] [--ca-distance X] [--total-match N] [--covered-node N]\n"
    "        [--covered-node-ratio X] [--max-node N] [--max-node-ratio X] [--score X] [--connected-node N]\n"
    "        [--connected-node-ratio X] [--num-residue N] [--plddt X] [--rmsd X] [--top N] [--sampling-count N]\n"
    "        [--sampling-ratio X] [--freq-filter X] [--length-penalty X] [--per-structure|--per-match] [--skip-match]\n"
    "        [--skip-ca-match] [--serial-index] [--sort-by KEY[:asc|desc],..] [--format-output COL,..] [--header]\n"
    "        [--tm-score X] [--gdt-ts X] [--gdt-ha X] [--chamfer X] [--hausdorff X] [--superpose] [--web] [--partial-fit] [-o FILE] [-v]\n"
    "        match columns: qid tid nid db_key node_count idf rmsd e_value u_matrix t_vector matching_residues\n"
    "                       matching_coordinates query_residues tm_score gdt_ts gdt_ha chamfer_distance hausdorff_distance\n";

int cmd_index(Args &a) {
    a.reject({"--mmap-on-disk"}, "the index is built in GPU memory");
    const std::string multiple_bins = a.str({"--multiple-bins"}, "");
    const std::string dir = a.str({"-p", "--pdbs"}, "");
    const std::string type = a.str({"-y", "--type"}, "default");
    const std::string prefix = a.str({"-i", "--index"}, "");
    const int threads = (int)a.num({"-t", "--threads"}, 1);
    fd_hash_params hp;
    memset(&hp, 0, sizeof(hp));
    hp.nbin_dist = (uint32_t)a.num({"-d", "--distance"}, 0);
    hp.nbin_angle = (uint32_t)a.num({"-a", "--angle"}, 0);
    hp.dist_cutoff = (float)a.num({"-g", "--grid"}, 20.0);
    const uint64_t max_residue = (uint64_t)a.num({"-n", "--residue"}, 50000);
    const bool recursive = a.flag({"-r", "--recursive"});
    const std::string id_type = a.str({"--id"}, "relpath");
    const bool verbose = a.flag({"-v", "--verbose"});
    const bool no_store = a.flag({"--no-store"});
    if (a.flag({"-h", "--help"})) {
        fputs(HELP, stdout);
        return 0;
    }
    a.finish();
    { // HashType::get_with_str (geometry/core.rs:42-57)
        const int t = fdh_hash_type_from_string(type.c_str());
        if (t < 0) die("unknown hash type '" + type + "'");
        hp.hash_type = (uint32_t)t;
    }
    if (!multiple_bins.empty()) { // parse_distance_angle_pairs (utils/cli.rs:1-16): "16-4,8-3"; malformed pairs are skipped
        size_t pos = 0;
        while (pos <= multiple_bins.size()) {
            const size_t comma = multiple_bins.find(',', pos);
            std::string pr = multiple_bins.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
            pos = comma == std::string::npos ? multiple_bins.size() + 1 : comma + 1;
            while (!pr.empty() && pr.front() == ' ') pr.erase(pr.begin());
            while (!pr.empty() && pr.back() == ' ') pr.pop_back();
            const size_t dash = pr.find('-');
            if (dash == std::string::npos || pr.find('-', dash + 1) != std::string::npos) continue;
            char *e1 = nullptr, *e2 = nullptr;
            const std::string a1 = pr.substr(0, dash), a2 = pr.substr(dash + 1);
            const unsigned long d = strtoul(a1.c_str(), &e1, 10), g = strtoul(a2.c_str(), &e2, 10);
            if (a1.empty() || a2.empty() || *e1 || *e2) continue;
            if (hp.n_multiple_bins >= FD_MAX_MULTIPLE_BINS) die("--multiple-bins: at most 8 pairs");
            hp.multiple_bins[2 * hp.n_multiple_bins] = (uint32_t)d;
            hp.multiple_bins[2 * hp.n_multiple_bins + 1] = (uint32_t)g;
            hp.n_multiple_bins++;
        }
    }
    if (dir.empty() || prefix.empty()) die("index needs -p DIR and -i PREFIX");
    if (threads > 0) setenv("FD_HOST_THREADS", std::to_string(threads).c_str(), 0);
    // -p is a directory of structure files or a Foldcomp database (build_index.rs:104-123)
    std::vector<std::string> files;
    std::vector<uint64_t> db_keys;
    fdh_fcz_db *fcz = nullptr;
    if (is_dir(dir)) {
        list_files(dir, recursive, files);
    } else {
        fcz = fdh_fcz_db_open(dir.c_str());
        if (!fcz) die(dir + " is neither a directory nor a Foldcomp database: " + fdh_last_error());
        for (int64_t k = 0; k < fdh_fcz_db_size(fcz); k++) {
            files.push_back(fdh_fcz_db_name(fcz, k));
            db_keys.push_back(fdh_fcz_db_key(fcz, k));
        }
    }
    if (files.empty()) die("no input files in " + dir);
    if (!fcz)
        for (auto &f : files) { // read_structure_from_path (controller/io.rs:337-379)
            const std::string b = ends_with(f, ".gz") ? f.substr(0, f.size() - 3) : f;
            if (!(ends_with(b, ".pdb") || ends_with(b, ".ent") || ends_with(b, ".PDB") || ends_with(b, ".cif")))
                die("unsupported input format (.pdb / .ent / .cif, optionally .gz): " + f);
        }
    if (verbose) fprintf(stderr, "[INFO] Indexing %zu files with %s\n", files.size(), fdh_hash_type_name(hp.hash_type));
    // parse on the host (file-parallel like the reference, mod.rs:298), keep file order
    std::vector<fdh_compact *> comps(files.size(), nullptr);
    {
        std::vector<std::string> errs(files.size());
        const int nt = std::max(1, std::min(fd_default_host_threads(), 64));
        std::atomic<size_t> next_file{0};
        auto work = [&](int) { // files differ in size by orders of magnitude: the threads take them one at a time
            for (size_t k; (k = next_file.fetch_add(1)) < files.size();) {
                comps[k] = fcz ? fdh_fcz_db_read(fcz, (int64_t)k) : fdh_compact_read_structure(files[k].c_str());
                if (!comps[k]) errs[k] = fdh_last_error();
            }
        };
        { // the library's worker pool is internal; plain threads are fine for a once-per-run parse
            std::vector<std::thread> th;
            for (int t = 1; t < nt; t++) th.emplace_back(work, t);
            work(0);
            for (auto &x : th) x.join();
        }
        for (size_t k = 0; k < files.size(); k++)
            if (!comps[k]) die("Failed to read structure " + files[k] + ": " + errs[k]);
    }
    if (fcz) fdh_fcz_db_close(fcz);
    fdh_store *store = fdh_store_new();
    fdh_compact *empty = fdh_compact_from_soa(0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    for (size_t k = 0; k < files.size(); k++) {
        const std::string id = make_id(files[k], id_type);
        if ((uint64_t)fdh_compact_num_residues_raw(comps[k]) > max_residue) { // mod.rs:313-319: skipped, id slot kept
            fprintf(stderr, "[WARN] %s has too many residues. Skipping\n", files[k].c_str());
            if (!empty) die("internal: cannot create an empty structure");
            fdh_store_add(store, empty, id.c_str());
        } else {
            fdh_store_add(store, comps[k], id.c_str());
        }
        fdh_compact_free(comps[k]);
    }
    if (empty) fdh_compact_free(empty);
    fd_ctx *ctx = nullptr;
    if (fd_create(&ctx, 0) != FD_OK) die(fd_last_error(nullptr));
    fdh_index *ix = fdh_index_build(ctx, store, &hp);
    if (!ix) die(fdh_last_error());
    if (!db_keys.empty() && fdh_index_set_db_keys(ix, db_keys.data(), db_keys.size()) != FD_OK) die(fdh_last_error());
    // the reference's default build (feature "foldcomp") records the -p argument as foldcomp_db for every input
    // (build_index.rs:228-232); input_format says whether it is a Foldcomp database
    if (fdh_index_save(ix, store, prefix.c_str(), max_residue, dir.c_str()) != FD_OK) die(fdh_last_error());
    if (!no_store) {
        if (fdh_store_save(store, (prefix + ".store").c_str()) != FD_OK) die(fdh_last_error());
    } else {
        // a store left by an earlier build at this prefix would describe other coordinates: remove it
        remove((prefix + ".store").c_str());
    }
    if (verbose) {
        fd_index_buffers b;
        fdh_index_get(ix, &b);
        fprintf(stderr, "[INFO] %llu structures, %llu residues, %llu hashes, %llu posting bytes -> %s\n",
                (unsigned long long)fdh_store_size(store), (unsigned long long)fdh_store_num_residues(store),
                (unsigned long long)b.count, (unsigned long long)b.value_bytes, prefix.c_str());
    }
    fdh_index_free(ix);
    fdh_store_free(store);
    fd_destroy(ctx);
    return 0;
}

struct QueryJob {
    std::string pdb, residues, output;
};

std::string residues_of(const fdh_results *R, const fdh_match_row &m, int64_t n_res) {
    const fdh_residue_match *res = fdh_results_residues(R) + m.res_begin;
    std::string s;
    for (int64_t k = 0; k < n_res; k++) {
        if (k) s += ',';
        if (res[k].some) {
            s += (char)res[k].chain;
            s += std::to_string((unsigned long long)res[k].serial);
        } else {
            s += '_';
        }
    }
    return s;
}

std::string escape_tsv(std::string s) { // formatter.rs:186-188
    for (char &c : s)
        if (c == '\t' || c == '\n') c = ' ';
    return s;
}


std::string lower_trim(std::string s) {
    size_t a = s.find_first_not_of(" \t"), b = s.find_last_not_of(" \t");
    s = a == std::string::npos ? "" : s.substr(a, b - a + 1);
    for (char &c : s) c = (char)tolower((unsigned char)c);
    return s;
}
std::vector<std::string> split(const std::string &s, char sep) {
    std::vector<std::string> out;
    size_t a = 0;
    for (;;) {
        const size_t b = s.find(sep, a);
        out.push_back(s.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return out;
}

// --sort-by: "key[:asc|desc],..." (sort.rs:47-66, 117-205 for match rows; :297-330, 380-440 for structure rows).
// A key is (column id, descending?); values compare as the reference's extract_value does, ties keep the prior order.
enum MatchKey { MK_NODE, MK_IDF, MK_RMSD, MK_EVALUE, MK_TM, MK_GDT_TS, MK_GDT_HA, MK_CHAMFER, MK_HAUSDORFF };
enum StructKey { SK_MAXNODE, SK_NODE, SK_IDF, SK_MINRMSD, SK_TOTAL, SK_EDGE, SK_NRES, SK_PLDDT };
struct SortSpec {
    std::vector<std::pair<int, bool>> keys; // (key, descending)
};
bool parse_order(const std::string &o, bool *desc) {
    if (o == "asc" || o == "ascending" || o == "a") {
        *desc = false;
        return true;
    }
    if (o == "desc" || o == "descending" || o == "d") {
        *desc = true;
        return true;
    }
    return false;
}
SortSpec parse_sort(const std::string &arg, bool structure_mode) {
    SortSpec sp;
    for (const std::string &part0 : split(arg, ',')) {
        const std::string part = lower_trim(part0);
        if (part.empty()) continue;
        const std::vector<std::string> kv = split(part, ':');
        if (kv.size() > 2) die("Error parsing --sort-by: Invalid format: '" + part + "'. Use 'key:order' or just 'key'");
        const std::string k = lower_trim(kv[0]);
        int key = -1;
        bool desc = true;
        auto is = [&](std::initializer_list<const char *> a) {
            for (const char *x : a)
                if (k == x) return true;
            return false;
        };
        if (structure_mode) {
            if (is({"max_node_count", "max-node-count", "max_node", "max-node", "max_nodes", "max-nodes"})) key = SK_MAXNODE;
            else if (is({"node_count", "node-count", "nodes", "node", "n"})) key = SK_NODE;
            else if (is({"idf", "score"})) key = SK_IDF;
            else if (is({"min_rmsd", "min-rmsd", "rmsd"})) key = SK_MINRMSD, desc = false;
            else if (is({"total_match_count", "total-match-count", "total_match", "total-match", "matches", "match"})) key = SK_TOTAL;
            else if (is({"edge_count", "edge-count", "edges", "edge", "e"})) key = SK_EDGE;
            else if (is({"nres", "num_residues", "num-residues", "length", "residues", "residue", "l"})) key = SK_NRES;
            else if (is({"plddt"})) key = SK_PLDDT;
            else die("Error parsing --sort-by: Unknown structure sort key: '" + k + "'. Valid keys: max_node_count, node_count, idf, min_rmsd, total_match_count, edge_count, nres, plddt");
        } else {
            if (is({"node_count", "node-count", "nodes", "node", "n"})) key = MK_NODE;
            else if (is({"idf", "score"})) key = MK_IDF;
            else if (is({"rmsd"})) key = MK_RMSD, desc = false;
            else if (is({"evalue", "e_value", "e-value"})) key = MK_EVALUE, desc = false; // sort.rs:52-60, 78-85
            else if (is({"tm_score", "tm-score", "tmscore", "tm"})) key = MK_TM;
            else if (is({"gdt_ts", "gdt-ts", "gdtts", "gdt"})) key = MK_GDT_TS;
            else if (is({"gdt_ha", "gdt-ha", "gdtha"})) key = MK_GDT_HA;
            else if (is({"chamfer", "chamfer-distance", "chamfer_distance"})) key = MK_CHAMFER, desc = false;
            else if (is({"hausdorff", "hausdorff-distance", "hausdorff_distance"})) key = MK_HAUSDORFF, desc = false;
            else die("Error parsing --sort-by: Unknown sort key: '" + k + "'");
        }
        if (kv.size() == 2 && !parse_order(lower_trim(kv[1]), &desc))
            die("Error parsing --sort-by: Unknown sort order: '" + kv[1] + "'. Use 'asc' or 'desc'");
        sp.keys.emplace_back(key, desc);
    }
    return sp;
}
// partial_cmp(..).unwrap_or(Equal) on f32 / f64 values
int cmp_val(double a, double b, bool desc) {
    if (desc) std::swap(a, b);
    return a < b ? -1 : (a > b ? 1 : 0);
}

// evalue_fitting (src/controller/result.rs:357-378): x = match idf, m = structures in the index, l = query residues
double evalue_fitting(float x, float m, float l) {
    const double x_d = x, m_d = m, l_d = l;
    const double mu = 4.2161 * std::exp(l_d * 0.0489) + 3.6661;
    const double lam = 0.2894 * std::exp(l_d * -0.0762) + 0.0316;
    const double k_val = std::exp(lam * mu) / 10546.0;
    const double e_val_raw = k_val * m_d * l_d * std::exp(-lam * x_d);
    return (e_val_raw * m_d) / (e_val_raw + m_d);
}
// Rust `{:.4e}`: mantissa with four decimals, exponent without sign padding ("1.2345e-3", "9.8765e2")
std::string rust_sci4(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[64];
    snprintf(buf, sizeof(buf), "%.4e", v);
    std::string t = buf;
    const size_t e = t.find('e');
    if (e == std::string::npos) return t;
    const int ex = atoi(t.c_str() + e + 1);
    return t.substr(0, e) + "e" + std::to_string(ex);
}

int cmd_query(Args &a) {
    // QueryMode::Web (controller/mode.rs:230-246, query_pdb.rs:481-493): per-match rows with the superposition columns,
    // at most MAX_NUM_LINES_FOR_WEB = 1000 of them; takes precedence over --skip-match / --per-structure
    const bool web = a.flag({"--web"});
    const bool superpose = a.flag({"--superpose"}) || web;
    // MatchFilter cutoffs over the similarity metrics (src/cli/workflows/query_pdb.rs:80-83, filter.rs:217-236); 0 = off
    const float tm_cut = (float)a.num({"--tm-score"}, 0.0), gdt_ts_cut = (float)a.num({"--gdt-ts"}, 0.0),
                gdt_ha_cut = (float)a.num({"--gdt-ha"}, 0.0), chamfer_cut = (float)a.num({"--chamfer"}, 0.0),
                hausdorff_cut = (float)a.num({"--hausdorff"}, 0.0);
    const std::string pdb_path = a.str({"-p", "--pdb"}, "");
    const std::string query_string = a.str({"-q", "--query"}, "");
    const int threads = (int)a.num({"-t", "--threads"}, 1);
    const std::string prefix = a.str({"-i", "--index"}, "");
    fdh_search_params sp;
    memset(&sp, 0, sizeof(sp));
    sp.skip_match = a.flag({"--skip-match"}) ? 1 : 0;
    const std::vector<float> dist_thr = parse_thresholds(a.str({"-d", "--distance"}, "0.5"));
    const std::vector<float> angle_thr = parse_thresholds(a.str({"-a", "--angle"}, "5"));
    sp.ca_dist_cutoff = (float)a.num({"--ca-distance"}, 1.0);
    sp.prefilter.total_match_count = (uint64_t)a.num({"--total-match"}, 0);
    sp.prefilter.covered_node_count = (uint64_t)a.num({"--covered-node"}, 0);
    sp.prefilter.covered_node_ratio = (float)a.num({"--covered-node-ratio"}, 0.0);
    sp.max_matching_node_count = (uint64_t)a.num({"--max-node"}, 0);
    sp.max_matching_node_ratio = (float)a.num({"--max-node-ratio"}, 0.0);
    sp.prefilter.idf_score_cutoff = (float)a.num({"--score"}, 0.0);
    sp.connected_node_count = (uint64_t)a.num({"--connected-node"}, 0);
    sp.connected_node_ratio = (float)a.num({"--connected-node-ratio"}, 0.0);
    sp.prefilter.num_res_cutoff = (uint64_t)a.num({"--num-residue"}, 50000);
    sp.prefilter.plddt_cutoff = (float)a.num({"--plddt"}, 0.0);
    sp.rmsd_cutoff = (float)a.num({"--rmsd"}, 0.0);
    const double top = a.num({"--top"}, -1);
    sp.prefilter.top_n = top < 0 ? UINT64_MAX : (uint64_t)top;
    sp.prefilter.sampling_count = (int64_t)a.num({"--sampling-count"}, -1);
    sp.prefilter.sampling_ratio = (float)a.num({"--sampling-ratio"}, -1.0);
    sp.prefilter.freq_filter = (float)a.num({"--freq-filter"}, -1.0);
    sp.prefilter.length_penalty = (float)a.num({"--length-penalty"}, 0.5);
    bool per_structure = a.flag({"--per-structure"});
    const bool per_match = a.flag({"--per-match"});
    sp.skip_ca_match = a.flag({"--skip-ca-match"}) ? 1 : 0;
    sp.partial_fit = a.flag({"--partial-fit"}) ? 1 : 0; // LMS-QCP superposition of matches above three residues
    const bool header = a.flag({"--header"});
    const bool serial_query = a.flag({"--serial-index"});
    const std::string output = a.str({"-o", "--output"}, "");
    const std::string sort_by = a.str({"--sort-by"}, "");
    std::string format_output;
    const bool has_format = a.value({"--format-output"}, format_output);
    const bool verbose = a.flag({"-v", "--verbose"});
    if (a.flag({"-h", "--help"})) {
        fputs(HELP, stdout);
        return 0;
    }
    a.finish();
    if (per_structure && per_match)
        die("Cannot print output per structure and per match at the same time. Use either --per-structure or --per-match");
    if (sp.skip_match) per_structure = true; // QueryMode::SkipMatch prints structure rows (query_pdb.rs:186-203)
    if (web) per_structure = false;
    const SortSpec sort_spec = parse_sort(sort_by, per_structure); // query_pdb.rs:212-245
    std::vector<std::string> columns;                              // --format-output (query_pdb.rs:248-254)
    if (has_format)
        for (const std::string &c : split(format_output, ',')) columns.push_back(lower_trim(c));
    if (prefix.empty()) die("query needs -i PREFIX");
    if (threads > 0) setenv("FD_HOST_THREADS", std::to_string(threads).c_str(), 0);
    // the similarity metrics (one small kernel per search, fd_metrics_store_batch) and the residue indices behind
    // matching_coordinates are computed only when a filter, sort key or column asks for them
    bool need_metrics = tm_cut > 0.f || gdt_ts_cut > 0.f || gdt_ha_cut > 0.f || chamfer_cut > 0.f || hausdorff_cut > 0.f;
    for (auto &kd : sort_spec.keys) need_metrics = need_metrics || (!per_structure && kd.first >= MK_TM);
    for (auto &c : columns)
        need_metrics = need_metrics || c == "tm_score" || c == "gdt_ts" || c == "gdt_ha" || c == "chamfer_distance" ||
                       c == "hausdorff_distance" || c == "matching_coordinates";
    if (superpose && !per_structure) need_metrics = true; // MATCH_RESULT_SUPERPOSE_COLUMNS has matching_coordinates
    sp.want_metrics = need_metrics && !sp.skip_match ? 1 : 0;

    std::vector<QueryJob> jobs; // query_pdb.rs:297-315
    if (ends_with(query_string, ".txt") || ends_with(query_string, ".tsv")) {
        std::ifstream in(query_string);
        if (!in) die("Failed to open query file: " + query_string);
        std::string line;
        while (std::getline(in, line)) {
            if (line.empty()) continue;
            QueryJob j;
            const size_t t1 = line.find('\t');
            j.pdb = line.substr(0, t1);
            if (t1 != std::string::npos) {
                const size_t t2 = line.find('\t', t1 + 1);
                j.residues = line.substr(t1 + 1, t2 == std::string::npos ? std::string::npos : t2 - t1 - 1);
                if (t2 != std::string::npos) {
                    const size_t t3 = line.find('\t', t2 + 1);
                    j.output = line.substr(t2 + 1, t3 == std::string::npos ? std::string::npos : t3 - t2 - 1);
                }
            }
            jobs.push_back(j);
        }
    } else {
        jobs.push_back(QueryJob{pdb_path, query_string, output});
    }
    if (jobs.empty()) die("no queries");
    for (auto &j : jobs) {
        if (j.pdb.empty()) die("query needs -p PDB (or a query file)");
        // (an empty -q makes every residue a query residue, query.rs:226-233: count_query answers it through its
        // global-memory path, the verification through the general path)
    }

    fd_ctx *ctx = nullptr;
    if (fd_create(&ctx, 0) != FD_OK) die(fd_last_error(nullptr));
    fdh_index *ix = fdh_index_load(prefix.c_str());
    if (!ix) die(fdh_last_error());
    if (fdh_index_attach(ctx, ix) != FD_OK) die(fd_last_error(ctx));
    const uint64_t S = fdh_index_num_structs(ix);
    std::vector<uint32_t> nres(S);
    std::vector<float> plddt(S);
    fdh_index_get_lookup(ix, nres.data(), plddt.data());
    // structures for the verification: PREFIX.store, else the files named by the lookup (index built by `folddisco`)
    fdh_store *store = nullptr;
    if (!sp.skip_match) {
        const std::string sp_path = prefix + ".store";
        if (is_file(sp_path)) {
            store = fdh_store_load(sp_path.c_str());
            if (!store) die(fdh_last_error());
            if (fdh_store_size(store) != S) die(sp_path + " does not belong to this index (structure count differs)");
            // same count is not enough (an index rebuilt at this prefix, e.g. by the reference binary, leaves the old
            // store behind): every structure must agree with the lookup in residue count and name
            {
                std::vector<uint32_t> s_nres(S);
                std::vector<float> s_plddt(S);
                fdh_store_get_lookup(store, s_nres.data(), s_plddt.data());
                for (uint64_t k = 0; k < S; k++)
                    if (s_nres[k] != nres[k] || strcmp(fdh_store_name(store, k), fdh_index_name(ix, k)) != 0)
                        die(sp_path + " does not belong to this index (structure " + std::to_string(k) + ": " +
                            fdh_store_name(store, k) + " / " + std::to_string(s_nres[k]) + " residues, lookup says " +
                            fdh_index_name(ix, k) + " / " + std::to_string(nres[k]) + "); re-run `index` or delete it");
            }
        } else {
            if (verbose) fprintf(stderr, "[INFO] %s not found: parsing the %llu structures named in the lookup\n", sp_path.c_str(), (unsigned long long)S);
            store = fdh_store_new();
            const size_t sl = prefix.find_last_of('/');
            const std::string index_dir = sl == std::string::npos ? "" : prefix.substr(0, sl + 1);
            // an index built from a Foldcomp database reads its structures from that database (query_pdb.rs:320-341):
            // the recorded path, else PREFIX-without-_folddisco + "_foldcomp", else PREFIX-without-_folddisco
            // (get_foldcomp_db_path_with_prefix, controller/io.rs:423-448)
            fdh_fcz_db *fcz = nullptr;
            if (*fdh_index_foldcomp_db(ix)) {
                std::string dbp = fdh_index_foldcomp_db(ix);
                if (!is_file(dbp)) {
                    std::string base = prefix;
                    if (ends_with(base, "_folddisco")) base = base.substr(0, base.size() - 10);
                    for (const std::string &cand : {base + "_foldcomp", base})
                        if (is_file(cand) && is_file(cand + ".index") && is_file(cand + ".lookup")) {
                            dbp = cand;
                            break;
                        }
                }
                fcz = fdh_fcz_db_open(dbp.c_str());
                if (!fcz) die(std::string("cannot open the Foldcomp database of this index: ") + fdh_last_error());
            }
            // the structures are read in parallel (the reference re-reads every candidate inside its rayon loop,
            // retrieve.rs:375), a few thousand at a time, and added to the store in id order
            const uint64_t CHUNK = 4096;
            const int nt = std::max(1, std::min(fd_default_host_threads(), 64));
            for (uint64_t k0 = 0; k0 < S; k0 += CHUNK) {
                const uint64_t kn = std::min(CHUNK, S - k0);
                std::vector<fdh_compact *> got(kn, nullptr);
                std::vector<std::string> errs(kn);
                std::atomic<uint64_t> next_k{0};
                auto work = [&](int) {
                    for (uint64_t j; (j = next_k.fetch_add(1)) < kn;) {
                        std::string p = fdh_index_name(ix, k0 + j);
                        if (fcz) {
                            const int64_t e = fdh_fcz_db_find(fcz, p.c_str());
                            if (e < 0) {
                                errs[j] = "Entry with name " + p + " not found.";
                                continue;
                            }
                            got[j] = fdh_fcz_db_read(fcz, e);
                        } else {
                            if (!is_file(p)) p = index_dir + p; // resolve_tid_path_from_index_prefix (controller/io.rs:488-528)
                            got[j] = fdh_compact_read_structure(p.c_str());
                        }
                        if (!got[j]) errs[j] = fdh_last_error();
                    }
                };
                {
                    std::vector<std::thread> th;
                    for (int t = 1; t < nt && (uint64_t)t < kn; t++) th.emplace_back(work, t);
                    work(0);
                    for (auto &x : th) x.join();
                }
                for (uint64_t j = 0; j < kn; j++) {
                    if (!got[j]) die(std::string("Failed to read structure ") + fdh_index_name(ix, k0 + j) + ": " + errs[j]);
                    fdh_store_add(store, got[j], fdh_index_name(ix, k0 + j));
                    fdh_compact_free(got[j]);
                }
            }
            if (fcz) fdh_fcz_db_close(fcz);
        }
        if (fdh_store_attach(store, ctx) != FD_OK) die(fdh_last_error());
    }

    fdh_query_params qp;
    fdh_index_get_params(ix, &qp.hash);
    qp.dist_thr = dist_thr.data();
    qp.n_dist_thr = (int)dist_thr.size();
    qp.angle_thr = angle_thr.data();
    qp.n_angle_thr = (int)angle_thr.size();
    qp.serial_query = serial_query ? 1 : 0;
    fdh_queries *qs = fdh_queries_new(&qp);
    if (!qs) die(fdh_last_error()); // an index of an encoding that is not built (TertiaryInteraction, Hybrid)
    {
        std::vector<std::pair<std::string, fdh_compact *>> cache; // a query file usually repeats structures
        for (auto &j : jobs) {
            fdh_compact *c = nullptr;
            for (auto &e : cache)
                if (e.first == j.pdb) c = e.second;
            if (!c) {
                c = fdh_compact_read_structure(j.pdb.c_str());
                if (!c) die("Failed to read structure: " + j.pdb + ": " + fdh_last_error());
                cache.emplace_back(j.pdb, c);
            }
            if (fdh_queries_add(qs, c, j.residues.c_str()) < 0) die(fdh_last_error());
        }
        for (auto &e : cache) fdh_compact_free(e.second);
    }
    if (fdh_queries_finalize(qs, ctx) != FD_OK) die(fdh_last_error());
    if (verbose) fprintf(stderr, "[INFO] Querying %zu motif(s) to %s\n", jobs.size(), prefix.c_str());
    fdh_results *R = fdh_search(ctx, qs, &sp, store);
    if (!R) die(fdh_last_error());

    const uint64_t *soff = fdh_results_struct_offsets(R), *moff = fdh_results_match_offsets(R);
    const fdh_struct_row *srows = fdh_results_struct_rows(R);
    const fdh_match_row *mrows = fdh_results_match_rows(R);
    const uint64_t *morder = fdh_results_match_order(R);
    for (size_t q = 0; q < jobs.size(); q++) {
        FILE *out = stdout;
        if (!jobs[q].output.empty()) {
            out = fopen(jobs[q].output.c_str(), "wb");
            if (!out) die("Failed to create file: " + jobs[q].output);
        }
        const int64_t n_res = fdh_queries_num_indices(qs, (int64_t)q);
        const std::string qres = escape_tsv(jobs[q].residues), qid = escape_tsv(jobs[q].pdb);
        char buf[64];
        auto f4 = [&](double v) {
            snprintf(buf, sizeof(buf), "%.4f", v);
            return std::string(buf);
        };
        if (per_structure) {
            // rows: default order from the library (idf desc, min_rmsd asc); --sort-by re-sorts them, stable
            std::vector<uint64_t> order;
            for (uint64_t k = soff[q]; k < soff[q + 1]; k++) order.push_back(k);
            if (!sort_spec.keys.empty()) {
                auto val = [&](const fdh_struct_row &r, int key) -> double {
                    switch (key) {
                        case SK_MAXNODE: return (float)r.max_matching_node_count;
                        case SK_NODE: return (float)r.node_count;
                        case SK_IDF: return r.idf;
                        case SK_MINRMSD: return r.min_rmsd_with_max_match;
                        case SK_TOTAL: return (float)r.total_match_count;
                        case SK_EDGE: return (float)r.edge_count;
                        case SK_NRES: return (float)nres[r.nid];
                        default: return plddt[r.nid];
                    }
                };
                std::stable_sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) {
                    for (auto &kd : sort_spec.keys) {
                        const int c = cmp_val(val(srows[x], kd.first), val(srows[y], kd.first), kd.second);
                        if (c) return c < 0;
                    }
                    return false;
                });
            }
            // STRUCTURE_RESULT_DEFAULT_COLUMNS (result.rs:300-313); unknown names are skipped like the reference's
            // filter_map over its column registry (result.rs:317-321)
            std::vector<std::string> cols = has_format ? columns
                                                       : std::vector<std::string>{"tid", "idf", "total_match_count", "node_count", "edge_count",
                                                                                  "max_node_cov", "min_rmsd", "nres", "plddt",
                                                                                  "matching_residues", "db_key", "query_residues"};
            const char *known[] = {"qid", "tid", "nid", "db_key", "total_match_count", "node_count", "edge_count", "idf", "nres",
                                   "plddt", "max_node_cov", "min_rmsd", "matching_residues", "query_residues"};
            std::vector<std::string> use;
            for (auto &c : cols)
                for (const char *k : known)
                    if (c == k) use.push_back(c);
            if (header) {
                for (size_t i = 0; i < use.size(); i++) fprintf(out, "%s%s", i ? "\t" : "", use[i].c_str());
                fputc('\n', out);
            }
            for (uint64_t k : order) {
                const fdh_struct_row &r = srows[k];
                for (size_t i = 0; i < use.size(); i++) {
                    const std::string &c = use[i];
                    std::string v;
                    if (c == "qid") v = qid;
                    else if (c == "tid") v = escape_tsv(fdh_index_name(ix, r.nid));
                    else if (c == "nid") v = std::to_string(r.nid);
                    else if (c == "db_key") v = std::to_string((unsigned long long)fdh_index_db_key(ix, r.nid));
                    else if (c == "total_match_count") v = std::to_string(r.total_match_count);
                    else if (c == "node_count") v = std::to_string(r.node_count);
                    else if (c == "edge_count") v = std::to_string(r.edge_count);
                    else if (c == "idf") v = f4(r.idf);
                    else if (c == "nres") v = std::to_string(nres[r.nid]);
                    else if (c == "plddt") {
                        snprintf(buf, sizeof(buf), "%.2f", (double)plddt[r.nid]);
                        v = buf;
                    } else if (c == "max_node_cov") v = std::to_string(r.max_matching_node_count);
                    else if (c == "min_rmsd") v = f4(r.min_rmsd_with_max_match);
                    else if (c == "matching_residues") {
                        for (uint64_t m = r.match_begin; m < r.match_end; m++) {
                            if (!v.empty()) v += ';';
                            v += residues_of(R, mrows[m], n_res) + ":" + f4(mrows[m].rmsd);
                        }
                        if (v.empty()) v = "NA";
                    } else v = qres;
                    fprintf(out, "%s%s", i ? "\t" : "", v.c_str());
                }
                fputc('\n', out);
            }
        } else {
            // rows in the default order (idf desc, rmsd asc) from the library, or emission order re-sorted by --sort-by
            // (stable, like par_sort_by over the candidate-ordered vector, result.rs:456-464)
            std::vector<uint64_t> order;
            const float *metrics = fdh_results_metrics(R);          // 5 per match row, or NULL
            const uint32_t *res_index = fdh_results_residue_index(R); // per residue entry, or NULL
            const float n_query_res = (float)fdh_queries_residue_count(qs, (int64_t)q);
            auto evalue_of = [&](const fdh_match_row &m) { return evalue_fitting(m.idf, (float)S, n_query_res); };
            // match_results.retain(|v| match_filter.filter(v)) for the metric cutoffs (query_pdb.rs:474, filter.rs:217-236)
            auto passes = [&](uint64_t k) {
                if (!metrics) return true;
                const float *mt = metrics + 5 * k;
                bool pass = true;
                if (tm_cut > 0.f) pass = pass && mt[0] >= tm_cut;
                if (gdt_ts_cut > 0.f) pass = pass && mt[1] >= gdt_ts_cut;
                if (gdt_ha_cut > 0.f) pass = pass && mt[2] >= gdt_ha_cut;
                if (chamfer_cut > 0.f) pass = pass && mt[3] <= chamfer_cut;
                if (hausdorff_cut > 0.f) pass = pass && mt[4] <= hausdorff_cut;
                return pass;
            };
            if (sort_spec.keys.empty()) {
                for (uint64_t k = moff[q]; k < moff[q + 1]; k++)
                    if (passes(morder[k])) order.push_back(morder[k]);
            } else {
                for (uint64_t k = moff[q]; k < moff[q + 1]; k++)
                    if (passes(k)) order.push_back(k);
                auto val = [&](uint64_t k, int key) -> double {
                    const fdh_match_row &m = mrows[k];
                    switch (key) {
                        case MK_NODE: return (double)m.node_count;
                        case MK_IDF: return (double)m.idf;
                        case MK_RMSD: return (double)m.rmsd;
                        case MK_EVALUE: return evalue_of(m);
                        default: return metrics ? (double)metrics[5 * k + (key - MK_TM)] : 0.0;
                    }
                };
                std::stable_sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) {
                    for (auto &kd : sort_spec.keys) {
                        const int c = cmp_val(val(x, kd.first), val(y, kd.first), kd.second);
                        if (c) return c < 0;
                    }
                    return false;
                });
            }
            const uint64_t print_limit = web ? 1000 : sp.prefilter.top_n; // result.rs:466-471, query_pdb.rs:489
            if (order.size() > print_limit) order.resize(print_limit);
            // MATCH_RESULT_DEFAULT_COLUMNS (result.rs:330-338)
            // MATCH_RESULT_SUPERPOSE_COLUMNS with --superpose (result.rs:341-352, :476-484)
            std::vector<std::string> cols =
                has_format ? columns
                : superpose ? std::vector<std::string>{"tid", "node_count", "idf", "rmsd", "matching_residues", "u_matrix",
                                                       "t_vector", "matching_coordinates", "db_key", "query_residues"}
                            : std::vector<std::string>{"tid", "node_count", "idf", "rmsd", "matching_residues", "query_residues"};
            const char *known[] = {"qid", "tid", "nid", "db_key", "node_count", "idf", "rmsd", "e_value", "u_matrix", "t_vector",
                                   "matching_residues", "matching_coordinates", "query_residues", "tm_score", "gdt_ts",
                                   "gdt_ha", "chamfer_distance", "hausdorff_distance"};
            static const char *metric_cols[5] = {"tm_score", "gdt_ts", "gdt_ha", "chamfer_distance", "hausdorff_distance"};
            std::vector<std::string> use;
            for (auto &c : cols)
                for (const char *k : known)
                    if (c == k) use.push_back(c);
            if (header) {
                for (size_t i = 0; i < use.size(); i++) fprintf(out, "%s%s", i ? "\t" : "", use[i].c_str());
                fputc('\n', out);
            }
            for (uint64_t k : order) {
                const fdh_match_row &m = mrows[k];
                for (size_t i = 0; i < use.size(); i++) {
                    const std::string &c = use[i];
                    std::string v;
                    if (c == "qid") v = qid;
                    else if (c == "tid") v = escape_tsv(fdh_index_name(ix, m.nid));
                    else if (c == "nid") v = std::to_string(m.nid);
                    else if (c == "db_key") v = std::to_string((unsigned long long)fdh_index_db_key(ix, m.nid));
                    else if (c == "node_count") v = std::to_string(m.node_count);
                    else if (c == "idf") v = f4(m.idf);
                    else if (c == "rmsd") v = f4(m.rmsd);
                    else if (c == "u_matrix") {
                        for (int j = 0; j < 9; j++) v += (j ? "," : "") + f4(m.U[j]);
                    } else if (c == "t_vector") {
                        for (int j = 0; j < 3; j++) v += (j ? "," : "") + f4(m.t[j]);
                    } else if (c == "matching_residues") v = residues_of(R, m, n_res);
                    else if (c == "e_value") v = rust_sci4(evalue_of(m));
                    else if (c == "matching_coordinates") { // C-alpha of the matched target residues (retrieve.rs:769-771)
                        bool first = true;
                        for (int64_t r = 0; r < n_res && res_index && store; r++) {
                            const uint32_t idx = res_index[m.res_begin + r];
                            float xyz[3];
                            if (!idx || fdh_store_get_ca(store, m.nid, idx - 1, xyz) != FD_OK) continue;
                            for (int j = 0; j < 3; j++) {
                                v += (first ? "" : ",") + f4(xyz[j]);
                                first = false;
                            }
                        }
                    } else if (c == "query_residues") v = qres;
                    else {
                        for (int j = 0; j < 5; j++)
                            if (c == metric_cols[j]) v = f4(metrics ? metrics[5 * k + j] : 0.0);
                    }
                    fprintf(out, "%s%s", i ? "\t" : "", v.c_str());
                }
                fputc('\n', out);
            }
        }
        if (out != stdout) fclose(out);
    }
    fdh_results_free(R);
    fdh_queries_free(qs);
    if (store) fdh_store_free(store);
    fdh_index_free(ix);
    fd_destroy(ctx);
    return 0;
}

} // namespace

int main(int argc, char **argv) {
    if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) {
        fputs(HELP, stdout);
        return argc < 2 ? 1 : 0;
    }
    const std::string cmd = argv[1];
    Args a(argc, argv);
    if (cmd == "index") return cmd_index(a);
    if (cmd == "query") return cmd_query(a);
    if (cmd == "version") {
        printf("%s\n", fd_version());
        return 0;
    }
    if (cmd == "benchmark" || cmd == "analyze")
        die("sub-command '" + cmd + "' is outside the ported path (use the reference binary on the produced TSV)");
    die("Invalid subcommand");
}
