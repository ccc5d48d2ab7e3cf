/* folddisco_b200_host.h -- host side above the C ABI of folddisco_b200.h.
 *
 * The reference's host orchestration is Rust; no Rust toolchain exists in this environment, so the host side
 * is C++ compiled into the same shared library and exported with C linkage.  It mirrors the reference's own
 * interface for the hot path (same names, argument meaning and error behaviour) so a Rust host can either
 * bind the kernels directly (folddisco_b200.h) or these orchestration entry points:
 *
 *   reference                                               here
 *   read_structure_from_path + Structure::to_compact        fdh_compact_read_pdb / fdh_compact_from_atoms
 *     (src/controller/io.rs:337-379, src/structure/core.rs:70-214)
 *   parse_query_string (src/controller/query.rs:331-384)    fdh_parse_query_string
 *   Folddisco::{collect_and_count, add_entries} + writers   fdh_store_* + fdh_index_build / fdh_index_save
 *     (src/controller/mod.rs:274-441, src/index/indextable.rs, src/index/lookup.rs, src/cli/config.rs)
 *   load_folddisco_index + load_lookup_from_file            fdh_index_load
 *   make_query_map + count_query + retrieval_wrapper        fdh_query_batch_* (one call per batch of queries)
 *     (src/cli/workflows/query_pdb.rs:348-452)
 *
 * Every compute step goes through the CUDA kernels of folddisco_b200.h; nothing here hashes structures,
 * scans postings or superposes on the CPU except the per-query feature of the k(k-1) query residue pairs
 * (make_query_map, microseconds) which uses the same fd_geom.cuh source as the kernels.
 */
#ifndef FOLDDISCO_B200_HOST_H
#define FOLDDISCO_B200_HOST_H

#include "folddisco_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdh_compact fdh_compact; /* CompactStructure */
typedef struct fdh_store fdh_store;     /* an ordered set of CompactStructures = the database / lookup */
typedef struct fdh_index fdh_index;     /* host copy (or mmap) of PREFIX / PREFIX.offset / .lookup / .type */
typedef struct fdh_queries fdh_queries; /* a batch of parsed queries with their query maps */
typedef struct fdh_results fdh_results; /* per-structure and per-match rows of a batch */

const char *fdh_last_error(void);

/* ---- structures ---- */
/* read_structure_from_path (src/controller/io.rs:337-379) + Structure::to_compact: .pdb / .ent (src/structure/io/pdb.rs)
 * and .cif (src/structure/io/cif.rs: the atom_site loop, first model), each optionally gzip-compressed (.gz).
 * NULL + fdh_last_error() on failure.  fdh_compact_read_pdb is the older name of the same function. */
fdh_compact *fdh_compact_read_structure(const char *path);
fdh_compact *fdh_compact_read_pdb(const char *path);
fdh_compact *fdh_compact_from_atoms(int64_t n_atoms, const float *x, const float *y, const float *z,
                                    const uint8_t *atom_name4, const uint8_t *chain, const uint8_t *res_name3,
                                    const uint64_t *res_serial, const float *b_factor);
/* aa uses the fd_struct_batch convention (128 + code = modified residue) */
fdh_compact *fdh_compact_from_soa(int64_t n, const float *n_xyz, const float *ca_xyz, const float *cb_xyz,
                                  const uint8_t *cb_valid, const uint8_t *aa, const uint8_t *chain,
                                  const uint64_t *serial, const float *b_factor);
/* ---- Foldcomp input (src/structure/io/fcz.rs) ----
 * The codec itself is the third-party Foldcomp library, bound at run time through the C ABI the reference binds
 * (lib/foldcomp/foldcompffi.h:18-21; fcz.rs:80-94): dlopen of $FD_FOLDCOMP_LIB, else "libfoldcomp_ffi.so".  Without it
 * every function below that decodes an entry fails with a message (NULL + fdh_last_error()).
 * fdh_compact_read_structure also accepts "DB:name" (an entry of a database, controller/io.rs:303-334) and "*.fcz"
 * (one entry in a file of its own). */
typedef struct fdh_fcz_db fdh_fcz_db; /* FoldcompDbReader (fcz.rs:21-136): PATH + PATH.index + PATH.lookup */
fdh_fcz_db *fdh_fcz_db_open(const char *path);
void fdh_fcz_db_close(fdh_fcz_db *db);
/* the named entries in ascending key order = get_paths / get_db_key_vector (fcz.rs:123-136, 208-219) */
int64_t fdh_fcz_db_size(const fdh_fcz_db *db);
const char *fdh_fcz_db_name(const fdh_fcz_db *db, int64_t k);
uint64_t fdh_fcz_db_key(const fdh_fcz_db *db, int64_t k);
int64_t fdh_fcz_db_find(const fdh_fcz_db *db, const char *name); /* position of the entry with this name, -1 if none */
fdh_compact *fdh_fcz_db_read(const fdh_fcz_db *db, int64_t k);   /* read_single_structure_by_id + to_compact */
fdh_compact *fdh_compact_from_fcz(const uint8_t *bytes, uint64_t n_bytes); /* one compressed entry */

int64_t fdh_compact_nres(const fdh_compact *c);
int64_t fdh_compact_num_residues_raw(const fdh_compact *c); /* Structure.num_residues (serial changes) */
int fdh_compact_first_chain(const fdh_compact *c);
float fdh_compact_avg_plddt(const fdh_compact *c);
void fdh_compact_get(const fdh_compact *c, float *n_xyz, float *ca_xyz, float *cb_xyz, uint8_t *cb_valid,
                     uint8_t *aa, uint8_t *chain, uint64_t *serial, float *b_factor);
void fdh_compact_free(fdh_compact *c);

/* parse_path_by_id_type (src/controller/mode.rs:19-31, 70-125): the id recorded in PREFIX.lookup for an input path under
 * `--id TYPE` (pdb, afdb, uniprot, filename, basename, abspath, relpath / anything else = the path itself).  Returns the
 * id's length; the id is written to out (0-terminated) when cap > length. */
int64_t fdh_parse_path_by_id_type(const char *path, const char *id_type, char *out, uint64_t cap);

/* ---- store ---- */
fdh_store *fdh_store_new(void);
/* takes a copy; returns the structure id (position) */
int64_t fdh_store_add(fdh_store *s, const fdh_compact *c, const char *name);
/* bulk add of n structures from one SoA (synthetic data); names are "prefix%llu" */
int64_t fdh_store_add_soa(fdh_store *s, uint64_t n_structs, const uint64_t *row_offsets, const float *n_xyz,
                          const float *ca_xyz, const float *cb_xyz, const uint8_t *aa, const char *name_prefix);
/* PREFIX.store: binary companion of the index with the compact structures, so `query` needs no per-candidate file
 * parse (the reference re-reads every candidate structure, retrieve.rs:375).  Layout in fd_host.cpp. */
int fdh_store_save(const fdh_store *s, const char *path);
fdh_store *fdh_store_load(const char *path); /* NULL + fdh_last_error() on failure */
uint64_t fdh_store_size(const fdh_store *s);
uint64_t fdh_store_num_residues(const fdh_store *s);
void fdh_store_get_lookup(const fdh_store *s, uint32_t *nres, float *plddt);
const char *fdh_store_name(const fdh_store *s, uint64_t id);
/* view usable with fd_build_index / fd_store_attach; valid until the store changes */
int fdh_store_batch(const fdh_store *s, fd_struct_batch *out);
/* fd_store_attach of the store's structures + fd_store_attach_labels of its (chain, residue number) labels: with the
 * labels on the device a search assembles its result rows there (fd_verify_rows) */
int fdh_store_attach(const fdh_store *s, fd_ctx *ctx);
void fdh_store_free(fdh_store *s);

/* ---- index files ---- */
/* Folddisco::collect_and_count .. save_offset_to_file on the GPU (fd_build_index) */
fdh_index *fdh_index_build(fd_ctx *ctx, const fdh_store *s, const fd_hash_params *params);
/* wraps existing arrays (copied) with the lookup of store s; for callers that built the index elsewhere */
fdh_index *fdh_index_from_buffers(const fd_index_buffers *b, const fdh_store *s, const fd_hash_params *params);
/* writes PREFIX, PREFIX.offset, PREFIX.lookup, PREFIX.type exactly like `folddisco index` */
int fdh_index_save(const fdh_index *ix, const fdh_store *s, const char *prefix, uint64_t max_residue,
                   const char *foldcomp_db /* NULL omits the key */);
/* load_folddisco_index (mmap) + load_lookup_from_file + read_index_config_from_file */
fdh_index *fdh_index_load(const char *prefix);
int fdh_index_get(const fdh_index *ix, fd_index_buffers *view); /* pointers owned by ix */
uint64_t fdh_index_num_structs(const fdh_index *ix);
void fdh_index_get_lookup(const fdh_index *ix, uint32_t *nres, float *plddt);
const char *fdh_index_name(const fdh_index *ix, uint64_t id);
uint64_t fdh_index_db_key(const fdh_index *ix, uint64_t id); /* 5th column of PREFIX.lookup */
/* keys of an index built from a Foldcomp database (mod.rs:151, lookup.rs:36-40); save then writes input_format "FCZDB" */
int fdh_index_set_db_keys(fdh_index *ix, const uint64_t *keys, uint64_t n);
/* the Foldcomp database a loaded index was built from (PREFIX.type: input_format = "FCZDB" + foldcomp_db), "" if none */
const char *fdh_index_foldcomp_db(const fdh_index *ix);
void fdh_index_get_params(const fdh_index *ix, fd_hash_params *params);
/* HashType::get_with_str / to_string (src/geometry/core.rs:42-75): a `--type` spelling ("default", "pdb", "ppf", "3",
 * "PDBMotifSinCos" ...) -> FD_HASH_* (6 / 7 for the two encodings that are not built), -1 if unknown; and back to the
 * name PREFIX.type stores */
int fdh_hash_type_from_string(const char *name);
const char *fdh_hash_type_name(uint32_t hash_type);
/* fd_index_attach with this index and its lookup */
int fdh_index_attach(fd_ctx *ctx, const fdh_index *ix);
void fdh_index_free(fdh_index *ix);

/* ---- queries ---- */
/* Returns number of residues or <0; see query.rs:331-384.  subs_off[i] < 0 = no substitution for residue i,
 * else subs[subs_off[i] .. subs_end[i]). */
int64_t fdh_parse_query_string(const char *q, uint8_t default_chain, uint8_t *chains, uint64_t *serials,
                               int64_t *subs_off, int64_t *subs_end, uint8_t *subs, int64_t cap_res,
                               int64_t cap_subs);

typedef struct {
    fd_hash_params hash;    /* from the index .type */
    const float *dist_thr;  /* -d, default {0.5} */
    int n_dist_thr;
    const float *angle_thr; /* -a, default {5.0} (degrees) */
    int n_angle_thr;
    int serial_query;       /* --serial-index */
} fdh_query_params;

fdh_queries *fdh_queries_new(const fdh_query_params *p);
/* adds one query: structure + query string (an empty string makes every residue a query residue, query.rs:226-233).
 * The batch shares the structure with the handle (structures are immutable and reference counted inside the library),
 * so the caller may fdh_compact_free it right after the call.  Returns the query number or <0. */
int64_t fdh_queries_add(fdh_queries *qs, const fdh_compact *structure, const char *query_string);
/* n queries at once, query maps built on `threads` host threads (0 = all cores); returns the first query number */
int64_t fdh_queries_add_many(fdh_queries *qs, const fdh_compact *const *structures, const char *const *query_strings,
                             int64_t n, int threads);
/* the same for batches that reuse a few structures / query strings (a query file usually does): query k is
 * (structures[which_structure[k]], query_strings[which_string[k]]); every query map is still built on its own */
int64_t fdh_queries_add_many_indexed(fdh_queries *qs, const fdh_compact *const *structures, int64_t n_structures,
                                     const char *const *query_strings, int64_t n_strings, const uint32_t *which_structure,
                                     const uint32_t *which_string, int64_t n, int threads);
int64_t fdh_queries_size(const fdh_queries *qs);
/* finishes make_query_map for the whole batch: one fd_posting_counts call supplies the per-edge idf
 * (calculate_idf_for_hash, query.rs:17-32).  Needs an attached index. */
int fdh_queries_finalize(fdh_queries *qs, fd_ctx *ctx);
int64_t fdh_queries_num_hashes(const fdh_queries *qs, int64_t q);
/* query-map entries of query q in insertion order */
void fdh_queries_get_map(const fdh_queries *qs, int64_t q, uint32_t *hash, int64_t *qi, int64_t *qj,
                         uint8_t *primary, float *idf);
int64_t fdh_queries_num_indices(const fdh_queries *qs, int64_t q);
/* residue_count of src/cli/workflows/query_pdb.rs:355-359: the PARSED query residues, resolved in the structure or not
 * (the denominator of the node-ratio filters and the query length of the e-value fit, result.rs:145) */
int64_t fdh_queries_residue_count(const fdh_queries *qs, int64_t q);
void fdh_queries_get_indices(const fdh_queries *qs, int64_t q, int64_t *indices);
void fdh_queries_free(fdh_queries *qs);

typedef struct {
    fd_prefilter_params prefilter;
    float ca_dist_cutoff;             /* --ca-distance, default 1.0 */
    int skip_match;                   /* --skip-match */
    uint64_t max_matching_node_count; /* --max-node */
    float max_matching_node_ratio;    /* --max-node-ratio */
    float rmsd_cutoff;                /* --rmsd, 0 = none */
    uint64_t connected_node_count;    /* --connected-node */
    float connected_node_ratio;       /* --connected-node-ratio */
    int skip_ca_match;                /* --skip-ca-match */
    int host_threads;                 /* threads for the graph / residue-mapping step, 0 = all cores */
    int verify_mode;                  /* 0: fused kernel, general path only for flagged candidates; 1: general path only */
    /* != 0: every match row also gets its StructureSimilarityMetrics (tm_score, gdt_ts, gdt_ha, chamfer_distance,
     * hausdorff_distance; src/controller/retrieve.rs:776-831, src/structure/metrics.rs) from fd_metrics_store_batch,
     * and the residue indices behind the matched residues are kept (fdh_results_metrics / _residue_index) */
    int want_metrics;
    /* --partial-fit: matches of more than three residues are superposed by LMS-QCP (fd_lmsqcp_store_batch) and report
     * the RMSD of the inlier core (src/controller/retrieve.rs:773-814); the search then verifies through the general path */
    int partial_fit;
} fdh_search_params;

/* query_pdb.rs:348-452 for the whole batch: count_query -> filter/sort/top -> retrieval -> Kabsch ->
 * filter_after_matching -> MatchFilter -> default sorts.  Needs fd_index_attach and, unless skip_match,
 * fd_store_attach on ctx. */
/* labels (may be NULL): store whose per-residue (chain, residue number) label the matched target residues;
 * without it fdh_residue_match.serial is the residue index inside the target and chain is 0. */
fdh_results *fdh_search(fd_ctx *ctx, const fdh_queries *qs, const fdh_search_params *p, const fdh_store *labels);
/* The same over a database whose structures are split into contiguous id ranges, one per rank / GPU (fd_comm_init on
 * every rank's ctx first).  qs holds THIS RANK'S OWN queries; ctx has the rank's LOCAL index attached (ids 0 ..,
 * database ids first_id ..) and a structure store with ALL structures of the database (the owner of a query verifies
 * its candidates wherever they live; 38 B per residue).  fdh_queries_finalize_sharded all-gathers the ranks' query
 * descriptors and all-reduces the local posting counts into global list lengths (the reference's idf weights);
 * fdh_search_sharded = fd_count_query_sharded (local scan of the whole batch, one all-to-all of per-query top-n
 * blocks, owner-side merge) + verification + rows of the rank's own queries.  Collective calls: every rank calls
 * both, with the same search parameters.  Rows are identical to fdh_search on the unsharded database. */
int fdh_queries_finalize_sharded(fdh_queries *qs, fd_ctx *ctx, uint64_t first_id, uint64_t total_structs);
fdh_results *fdh_search_sharded(fd_ctx *ctx, const fdh_queries *qs, const fdh_search_params *p, const fdh_store *labels);

/* ---- hash-range sharded index (one rank per GPU): the same search in three steps around the caller's collective.
 *   0. every rank: fdh_queries_set_shards (gives every (query edge, owning rank) pair its own vote bit)
 *   1. every rank: fdh_queries_pair_counts -> all-reduce(sum) of the counts -> fdh_queries_finalize_with_counts
 *      (a posting list lives on exactly one shard, so the sum is the global list length; total_structures = lookup.len())
 *   2. every rank: fdh_votes_scan (partial votes of its shard for the whole batch, fd_votes_scan) -> all-reduce
 *   3. rank r: fdh_search_from_votes for its slice [q_begin, q_end) of the batch (fd_votes_select + verification
 *      against the replicated store); the results object has q_end - q_begin queries. */
/* bounds[world + 1]: rank r owns the hashes in [bounds[r], bounds[r+1]); call once per batch before step 2 */
int fdh_queries_set_shards(fdh_queries *qs, const uint64_t *bounds, int world);
/* the fd_query arrays of query q as the scan will see them (hashes / bit_of_hash: fdh_queries_num_hashes entries in
 * scan order; bit_node / bit_group: fdh_queries_num_vote_bits entries) */
int64_t fdh_queries_num_vote_bits(const fdh_queries *qs, int64_t q);
void fdh_queries_get_vote_bits(const fdh_queries *qs, int64_t q, uint32_t *hashes, uint16_t *bit_of_hash,
                               uint16_t *bit_node, uint16_t *bit_group);
/* flat scan inputs of the batch, for the all-gather between ranks that each built the query maps of their own slice:
 * per_query = {hashes, vote bits, pairs} per query; hashes / bit_of_hash in scan order; the observed hash per pair */
void fdh_queries_scan_sizes(const fdh_queries *qs, uint64_t *n_hashes, uint64_t *n_pairs);
void fdh_queries_scan_arrays(const fdh_queries *qs, uint32_t *per_query, uint32_t *hashes, uint32_t *bit_of_hash,
                             uint32_t *pair_hashes);
/* fd_votes_scan_sparse for a whole batch given as gathered flat arrays */
int fdh_votes_scan_sparse_flat(fd_ctx *ctx, uint32_t nq, const uint32_t *per_query, const uint32_t *hashes,
                               const uint32_t *bit_of_hash, const fd_prefilter_params *prefilter,
                               const uint32_t *slice_begin, uint32_t world, fd_votes_layout *layout,
                               uint32_t **d_records, uint64_t *region_offset, uint64_t *region_count);
int64_t fdh_queries_num_pairs(const fdh_queries *qs);
int fdh_queries_pair_counts(const fdh_queries *qs, fd_ctx *ctx, uint32_t *out_counts /* fdh_queries_num_pairs */);
int fdh_queries_finalize_with_counts(fdh_queries *qs, const uint32_t *counts, uint64_t total_structures);
int fdh_votes_scan(fd_ctx *ctx, const fdh_queries *qs, const fd_prefilter_params *prefilter, fd_votes_layout *layout,
                   uint32_t **d_votes);
fdh_results *fdh_search_from_votes(fd_ctx *ctx, const fdh_queries *qs, const fdh_search_params *p,
                                   const fdh_store *labels, const fd_votes_layout *layout, const uint32_t *d_votes,
                                   uint32_t q_begin, uint32_t q_end);

/* per-structure rows: query q owns [struct_offsets[q], struct_offsets[q+1]) ordered idf desc, min_rmsd asc;
 * per-match rows: query q owns [match_offsets[q], match_offsets[q+1]) in emission order, fdh_results_match_order gives
 * the order idf desc, rmsd asc; matched residues: n_query_residues entries per match from res_begin.
 * (layouts declared in folddisco_b200.h: the device writes them, fd_verify_rows) */
typedef fd_struct_row fdh_struct_row;
typedef fd_match_row fdh_match_row;
typedef fd_residue_row fdh_residue_match;

uint64_t fdh_results_num_queries(const fdh_results *r);
const uint64_t *fdh_results_struct_offsets(const fdh_results *r);
const fdh_struct_row *fdh_results_struct_rows(const fdh_results *r);
const uint64_t *fdh_results_match_offsets(const fdh_results *r);
const fdh_match_row *fdh_results_match_rows(const fdh_results *r);
const uint64_t *fdh_results_match_order(const fdh_results *r); /* sorted position -> emission index */
const fdh_residue_match *fdh_results_residues(const fdh_results *r);
uint64_t fdh_results_num_residues(const fdh_results *r);
/* want_metrics searches only (NULL otherwise): five floats per match row, aligned with fdh_results_match_rows --
 * {tm_score, gdt_ts, gdt_ha, chamfer_distance, hausdorff_distance} (the MatchResult.metrics columns of
 * src/controller/result.rs:280-284) -- and per residue entry the target residue index + 1 (0 = unmatched) */
const float *fdh_results_metrics(const fdh_results *r);
const uint32_t *fdh_results_residue_index(const fdh_results *r);
/* C-alpha of residue `residue` of structure `id` (the matching_coordinates column, retrieve.rs:769-771) */
int fdh_store_get_ca(const fdh_store *s, uint64_t id, uint64_t residue, float *xyz3);
/* wall-clock milliseconds of the host-only part of the last search (graph + mapping + assembly) */
double fdh_results_host_ms(const fdh_results *r);
/* bytes the search copied host->device (query descriptors, candidate lists, alignment indices) and
 * device->host (hits, candidate edges/pairs, RMSD/U/t) */
uint64_t fdh_results_h2d_bytes(const fdh_results *r);
/* wall-clock ms of the last search: 0 = fd_count_query_batch call, 1 = verification, 2 = row assembly, 3 = total */
double fdh_results_wall_ms(const fdh_results *r, int which);
uint64_t fdh_results_d2h_bytes(const fdh_results *r);
void fdh_results_free(fdh_results *r);

#ifdef __cplusplus
}
#endif
#endif
