/* fd_oracle.h -- TEST INFRASTRUCTURE.  C API of the CPU oracle.
 *
 * The oracle is a plain CPU restatement of the reference algorithm
 * (steineggerlab/folddisco @ 9375a2d) for the hot path named in BASELINE.json.
 * It is the checker for tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference arm.  Nothing in folddisco_b200/ may include,
 * link or call it.
 *
 * The reference is a Rust crate and cannot be compiled in this environment
 * (no cargo/rustc), so there is no oracle/_ref; the restatement is pinned by
 * the reference's own known-answer vectors instead (see tests/test_oracle_golden.py):
 *   src/controller/graph.rs:71-79   six u32 hashes of the 4CHA triad edges
 *   README.md:218-224, 237-241      per-match / per-structure rows of config 1
 *   src/structure/io/pdb.rs:142     49 residues in data/homeobox/1akha-.pdb
 *   src/controller/query.rs:425-465 query-string parser cases
 *   src/structure/kabsch.rs:592-615 rmsd bounds
 * Parity that the reference itself does not pin (last-ulp libm behaviour, FxHashMap
 * iteration order inside f32 sums) is "unpinned"; see DESIGN.md.
 */
#ifndef FD_ORACLE_H
#define FD_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdo_structure fdo_structure; /* atom SoA  (src/structure/core.rs:9-16)  */
typedef struct fdo_compact fdo_compact;     /* CompactStructure (core.rs:55-67)        */
typedef struct fdo_index fdo_index;         /* FolddiscoIndex + lookup                  */
typedef struct fdo_qmap fdo_qmap;           /* make_query_map output (query.rs:208-329) */
typedef struct fdo_hits fdo_hits;           /* count_query output                       */
typedef struct fdo_matches fdo_matches;     /* retrieval_wrapper output                 */

/* 0 = exact binary64-evaluated math (default), 1 = glibc sinf/cosf/acosf/atan2f */
void fdo_set_math_mode(int use_libm);
float fdo_math_sinf(float x);
float fdo_math_cosf(float x);
float fdo_math_acosf(float x);
float fdo_math_atan2f(float y, float x);

/* ---- structures -------------------------------------------------------- */
fdo_structure *fdo_structure_read_pdb(const char *path);
fdo_structure *fdo_structure_from_atoms(int64_t n_atoms, const float *x, const float *y, const float *z,
                                        const uint8_t *atom_name4, const uint8_t *chain,
                                        const uint8_t *res_name3, const uint64_t *res_serial,
                                        const float *b_factor);
int64_t fdo_structure_num_atoms(const fdo_structure *s);
int64_t fdo_structure_num_residues(const fdo_structure *s);
int fdo_structure_first_chain(const fdo_structure *s); /* chains[0] or -1 */
void fdo_structure_get_atoms(const fdo_structure *s, float *x, float *y, float *z, uint8_t *atom_name4,
                             uint8_t *chain, uint8_t *res_name3, uint64_t *res_serial, float *b_factor);
void fdo_structure_free(fdo_structure *s);

fdo_compact *fdo_compact_build(const fdo_structure *s);
/* aa: 0..19 or 255; residue names become the canonical three-letter codes / "UNK". */
fdo_compact *fdo_compact_from_soa(int64_t n, const float *n_xyz, const float *ca_xyz, const float *cb_xyz,
                                  const uint8_t *cb_valid, const uint8_t *aa, const uint8_t *chain,
                                  const uint64_t *serial, const float *b_factor);
int64_t fdo_compact_nres(const fdo_compact *c);
void fdo_compact_get(const fdo_compact *c, float *n_xyz, float *ca_xyz, float *cb_xyz, uint8_t *cb_valid,
                     uint8_t *aa, uint8_t *res_name3, uint8_t *chain, uint64_t *serial, float *b_factor);
float fdo_compact_avg_plddt(const fdo_compact *c);
int64_t fdo_compact_get_index(const fdo_compact *c, uint8_t chain, uint64_t serial); /* -1 if none */
void fdo_compact_free(fdo_compact *c);

uint8_t fdo_map_aa_to_u8(const uint8_t *res_name3);

/* ---- geometric hash (pdb_tr / PDBTrRosetta) ----------------------------- */
/* returns 1 and fills out[7] = res1,res2,ca_dist,cb_dist,angle,theta1,theta2 if the pair has a feature */
int fdo_pair_feature(const fdo_compact *c, int64_t i, int64_t j, float dist_cutoff, float *out7);
uint32_t fdo_perfect_hash(const float *feature7, uint32_t nbin_dist, uint32_t nbin_angle);
int fdo_hash_is_symmetric(uint32_t hash);
/* Other encodings.  Process-wide selection, like fdo_set_math_mode: t = the reference's HashType index
 * (src/geometry/core.rs:25-38): 0 PDBMotif, 1 PDBMotifSinCos, 2 TrRosetta, 3 PDBTrRosetta (default), 4 PointPairFeature,
 * 7 FolddiscoAngle, 8 FolddiscoDist (5 TertiaryInteraction and 6 Hybrid are not restated: returns -1).  Every function
 * below that takes nbin_dist / nbin_angle then hashes with that encoding (features get 9 slots, unused ones 0), and
 * fdo_pair_feature / fdo_perfect_hash / fdo_hash_is_symmetric follow it too.
 * fdo_set_multiple_bins: the --multiple-bins list as n (dist, angle) pairs; n = 0 turns it off
 * (feature.rs:210-214, query.rs:59-71, retrieve.rs:124-131). */
int fdo_set_hash_type(int t);
/* the 9-slot feature of the selected encoding, and HashValue::perfect_hash without the zero -> default substitution */
int fdo_pair_feature9(const fdo_compact *c, int64_t i, int64_t j, float dist_cutoff, float *out9);
uint32_t fdo_perfect_hash_raw(const float *feature9, uint32_t nbin_dist, uint32_t nbin_angle);
int fdo_get_hash_type(void);
void fdo_set_multiple_bins(int n, const uint32_t *dist_angle_pairs);
/* all ordered pairs, row-major (feature.rs:198-231).  sorted_unique != 0 applies sort+dedup
 * (controller/mod.rs:343-344).  Returns the number of hashes; writes at most cap. */
int64_t fdo_hash_compact(const fdo_compact *c, uint32_t nbin_dist, uint32_t nbin_angle, float dist_cutoff,
                         int sorted_unique, uint32_t *out, int64_t cap);

/* ---- inverted index ------------------------------------------------------ */
/* from per-structure sorted-unique hashes in CSR form; ids = row numbers */
fdo_index *fdo_index_from_csr(const uint32_t *hashes, const uint64_t *row_offsets, uint64_t n_structs);
/* the reference's two-pass builder over compact structures (controller/mod.rs:274-441) */
fdo_index *fdo_index_build(const fdo_compact *const *structs, uint64_t n_structs, uint32_t nbin_dist,
                           uint32_t nbin_angle, float dist_cutoff, int threads);
fdo_index *fdo_index_from_buffers(const uint32_t *hashes, const uint64_t *offsets, uint64_t count,
                                  const uint8_t *values, uint64_t value_bytes);
fdo_index *fdo_index_load(const char *prefix);  /* PREFIX.offset + PREFIX (or PREFIX.value) */
int fdo_index_save(const fdo_index *ix, const char *prefix); /* PREFIX, PREFIX.offset */
uint64_t fdo_index_count(const fdo_index *ix);
uint64_t fdo_index_value_bytes(const fdo_index *ix);
const uint32_t *fdo_index_hashes(const fdo_index *ix);
const uint64_t *fdo_index_offsets(const fdo_index *ix);
const uint8_t *fdo_index_values(const fdo_index *ix);
/* decode one posting list (indextable.rs:83-86); returns length, writes at most cap ids */
int64_t fdo_index_get_entries(const fdo_index *ix, uint32_t hash, uint64_t *out, int64_t cap);
void fdo_index_free(fdo_index *ix);

/* .lookup / .type writers (index/lookup.rs:17-58, cli/config.rs:64-97) */
int fdo_lookup_save(const char *path, uint64_t n, const char *const *names, const uint64_t *nres,
                    const float *plddt);
int fdo_type_save(const char *path, uint32_t nbin_dist, uint32_t nbin_angle, float grid_width,
                  uint64_t chunk_size, uint64_t max_residue, const char *foldcomp_db /* NULL omits the key */);

/* ---- query ---------------------------------------------------------------- */
/* query.rs:331-384.  Returns number of residues; subs_off has n+1 entries, -1 marks "no substitution". */
int64_t fdo_parse_query_string(const char *q, uint8_t default_chain, uint8_t *chains, uint64_t *serials,
                               int64_t *subs_off, uint8_t *subs, int64_t cap_res, int64_t cap_subs);
/* sub_off[i] < 0: residue i has no substitution, else subs[sub_off[i] .. sub_end[i]) */
fdo_qmap *fdo_qmap_make(const fdo_compact *query, const uint8_t *chains, const uint64_t *serials,
                        int64_t n_res, const int64_t *sub_off, const int64_t *sub_end, const uint8_t *subs,
                        uint32_t nbin_dist, uint32_t nbin_angle, const float *dist_thr, int n_dist_thr,
                        const float *angle_thr, int n_angle_thr, float dist_cutoff, int serial_query,
                        const fdo_index *index /* may be NULL */, float total_structures);
int64_t fdo_qmap_size(const fdo_qmap *m);
/* insertion order */
void fdo_qmap_get(const fdo_qmap *m, uint32_t *hash, int64_t *qi, int64_t *qj, uint8_t *primary, float *idf);
int64_t fdo_qmap_num_indices(const fdo_qmap *m);
/* residue_count of src/cli/workflows/query_pdb.rs:355-359 (parsed query residues, resolved or not) */
int64_t fdo_qmap_residue_count(const fdo_qmap *m);
void fdo_qmap_get_indices(const fdo_qmap *m, int64_t *indices);
void fdo_qmap_free(fdo_qmap *m);

/* count_query.rs:82-220 + filter.rs:76-100 + query_pdb.rs:404-411 (stable sort idf desc, top_n).
 * sampling_ratio < 0 / sampling_count < 0 / freq_filter < 0 mean "None". */
typedef struct {
    float sampling_ratio;
    int64_t sampling_count;
    float freq_filter;
    float length_penalty;
    /* StructureFilter, 0 disables each (filter.rs:76-100) */
    uint64_t total_match_count;
    uint64_t covered_node_count;
    float covered_node_ratio;
    float idf_score_cutoff;
    uint64_t num_res_cutoff;
    float plddt_cutoff;
    uint64_t expected_node_count;
    uint64_t top_n; /* UINT64_MAX = all */
    int apply_filter_and_sort; /* 0: raw count_query output in ascending nid */
} fdo_count_params;
fdo_hits *fdo_count_query(const fdo_qmap *m, const fdo_index *ix, uint64_t n_structs, const uint64_t *nres,
                          const float *plddt, const fdo_count_params *p);
int64_t fdo_hits_size(const fdo_hits *h);
void fdo_hits_get(const fdo_hits *h, uint64_t *nid, uint32_t *match_count, uint32_t *node_count,
                  uint32_t *edge_count, float *idf);
void fdo_hits_free(fdo_hits *h);

/* retrieve.rs:364-552 against one target */
fdo_matches *fdo_retrieve(const fdo_qmap *m, const fdo_compact *query, const fdo_compact *target,
                          uint32_t nbin_dist, uint32_t nbin_angle, float dist_cutoff, float ca_dist_cutoff);
int64_t fdo_matches_size(const fdo_matches *r);
int64_t fdo_matches_num_query(const fdo_matches *r);
/* which = 0: result (rescued, default output); 1: result_from_hash (--skip-ca-match).
 * res_some/res_chain/res_serial are [n_matches * n_query]; U is 9 per match, t is 3. */
void fdo_matches_get(const fdo_matches *r, int which, uint8_t *res_some, uint8_t *res_chain,
                     uint64_t *res_serial, float *rmsd, float *idf, float *U, float *t);
/* `--partial-fit`: retrieve() superposes more than three matched residues with LMS-QCP (src/structure/lms_qcp.rs,
 * default parameters; retrieve.rs:773-814) and reports the RMSD of the inlier core.  Process-wide, like the hash type. */
void fdo_set_partial_fit(int on);
/* LmsQcpSuperimposer::run over explicit point lists (n >= 3): mov is rotated onto ref; returns the size of the inlier
 * core and writes at most cap of its indices in the order they joined */
int64_t fdo_lms_qcp(int64_t n, const float *ref3, const float *mov3, float *U9, float *t3, float *rms_inliers,
                    int64_t *inliers, int64_t cap);
/* StructureSimilarityMetrics of every match (src/structure/metrics.rs:44-345, computed in
 * rmsd_with_calpha_and_rottran, retrieve.rs:776-831): out5[5 * k] = tm_score, gdt_ts, gdt_ha, chamfer, hausdorff */
void fdo_matches_get_metrics(const fdo_matches *r, int which, float *out5);
/* the same over explicit point lists: n reference (query) points, n moving (target) points, the superposition U9 / t3 */
void fdo_similarity_metrics(int64_t n, const float *ref3, const float *mov3, const float *U9, const float *t3, float *out5);
int64_t fdo_matches_max_node_count(const fdo_matches *r);
float fdo_matches_min_rmsd(const fdo_matches *r);
/* retrieve_with_prefilter output: (i, j, hash) triples in emission order */
int64_t fdo_matches_num_edges(const fdo_matches *r);
void fdo_matches_get_edges(const fdo_matches *r, int64_t *ei, int64_t *ej, uint32_t *eh);
void fdo_matches_free(fdo_matches *r);

/* kabsch.rs:157-554, mode 2.  x is rotated onto y.  Returns 0 on success. */
int fdo_kabsch(int64_t n, const float *x3, const float *y3, float *U9, float *t3, float *rmsd);

/* ---- batched pipeline for the CPU baseline (bench.py) ---------------------- */
/* Runs count_query -> filter/sort/top -> retrieval for n_q query maps against an in-memory
 * store of compact structures, node-group/candidate parallel like the reference.  Returns the
 * total number of match rows and fills per-query summary arrays (may be NULL). */
int64_t fdo_query_batch(const fdo_qmap *const *maps, const fdo_compact *const *queries, int64_t n_q,
                        const fdo_index *ix, const fdo_compact *const *store, uint64_t n_structs,
                        const uint64_t *nres, const float *plddt, const fdo_count_params *p,
                        uint32_t nbin_dist, uint32_t nbin_angle, float dist_cutoff, float ca_dist_cutoff,
                        int skip_match, int threads, uint64_t *hits_per_query, uint64_t *matches_per_query,
                        uint64_t *posting_bytes_per_query);

#ifdef __cplusplus
}
#endif
#endif
