/* folddisco_b200.h -- C ABI of libfolddisco_b200.so
 *
 * B200 (sm_100a) drop-in for the data-parallel hot path of steineggerlab/folddisco @ 9375a2d:
 *   (i)  index build : all-residue-pair geometric hashing  -> delta+LEB128 posting lists
 *   (ii) query       : posting-list scan -> per-structure vote -> candidate re-hash -> Kabsch RMSD
 * The reference has no plugin/operator registry; this seam is modelled on its one existing FFI
 * (lib/foldcomp/foldcompffi.h:18-21 create/process/free/destroy, called from src/structure/io/fcz.rs:80-94).
 * Each entry point names the reference call site it replaces.  INTEGRATION.md shows the Rust binding.
 *
 * Conventions: plain C structs and pointers only.  The caller owns every input buffer; the library owns
 * device memory.  Outputs documented as "library-allocated" are released with fd_free().  Every call
 * returns FD_OK (0) or a negative error code and records a message readable with fd_last_error(); no C++
 * exception crosses the ABI.  An fd_ctx is bound to one CUDA device and is not thread-safe (one per host
 * thread / one per rank).  Calls are synchronous on return.  There is no CPU fallback: without a usable
 * CUDA device fd_create fails.
 */
#ifndef FOLDDISCO_B200_H
#define FOLDDISCO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FD_OK 0
#define FD_ERR_CUDA (-1)     /* CUDA runtime error (message has the cudaError string) */
#define FD_ERR_ARG (-2)      /* invalid argument */
#define FD_ERR_STATE (-3)    /* call order: e.g. query before fd_index_attach */
#define FD_ERR_NOMEM (-4)    /* host or device allocation failed */
#define FD_ERR_LIMIT (-5)    /* an implementation limit was exceeded (message says which) */

typedef struct fd_ctx fd_ctx;

/* ---- lifecycle --------------------------------------------------------------------------------- */
int fd_create(fd_ctx **out, int device);
void fd_destroy(fd_ctx *ctx);
/* A second context on the same device that SHARES the parent's attached index and structure store (read-only) but
 * has its own stream, events and staging buffers: a host can run two synchronous call sequences (one per host
 * thread) whose kernels, copies and host-side steps overlap.  The child must not attach anything itself and must be
 * destroyed before the parent; after the parent attaches a new index / store call fd_fork_refresh. */
int fd_fork(fd_ctx *parent, fd_ctx **out);
int fd_fork_refresh(fd_ctx *parent, fd_ctx *child);
/* the i-th (0..7) cached fork of ctx, created on first use, refreshed to ctx's current attachments, owned and
 * destroyed by ctx; fd_lanes_fold_stats adds the lanes' stage times / launch counts to ctx's and clears them */
int fd_lane(fd_ctx *ctx, int i, fd_ctx **out);
void fd_lanes_fold_stats(fd_ctx *ctx);
int fd_device(const fd_ctx *ctx);             /* CUDA device index of the context */
const char *fd_last_error(const fd_ctx *ctx); /* ctx may be NULL: last error of a failed fd_create */
void fd_free(void *p);                        /* for library-allocated host outputs */
const char *fd_version(void);
/* number of this library's kernels launched through ctx since creation (bench.py's gpu_launches) */
uint64_t fd_kernel_launches(const fd_ctx *ctx);
/* host threads the library uses for its query-parallel host steps when the caller passes 0: the environment variable
 * FD_HOST_THREADS if set (several ranks share one box: cores / ranks), else the number of hardware threads */
int fd_default_host_threads(void);
/* debug probe of the host worker pool: distinct host threads that ran a region of nt workers spinning spin_us each
 * (1..nt), or -1 if some worker index did not run exactly once */
int fd_parallel_probe(int nt, int spin_us);
/* cumulative device time (ms, CUDA events on the library's stream) of the named stage since creation:
 * "hash", "postings", "attach", "lookup", "scan", "select", "verify" (= "verify_edges" + "verify_components" +
 * "verify_kabsch"), "edges", "kabsch"; 0 if the stage never ran */
double fd_stage_ms(const fd_ctx *ctx, const char *stage);
/* bookkeeping of the in-library host: n candidates of the last search took the general verification path
 * (reported as the launch count of the pseudo-stage "general_candidates") */
void fd_note_general_path(fd_ctx *ctx, uint64_t n);
/* adds host wall time to the named pseudo-stage (read back with fd_stage_ms) */
void fd_note_host_ms(fd_ctx *ctx, const char *stage, double ms);
uint64_t fd_stage_launches(const fd_ctx *ctx, const char *stage);

/* ---- structures ---------------------------------------------------------------------------------- */
/* A batch of CompactStructures (reference src/structure/core.rs:55-67) as one SoA; structure s owns the
 * residues [row_offsets[s], row_offsets[s+1]).  xyz arrays are interleaved x,y,z per residue.
 * aa = map_aa_to_u8(residue name) (src/utils/convert.rs:53-81): 0..19, 255 = unknown.  A residue whose name is
 * not the canonical three-letter code of its amino acid (MSE -> MET, SEP -> SER, ...) carries 128 + code: it
 * hashes like the amino acid but is skipped by prefilter_amino_acid, which matches names exactly
 * (src/controller/retrieve.rs:580-596).  cb_valid may be NULL (all valid): 0 marks a residue without CB and
 * without a backbone C to place a virtual one (src/structure/core.rs:129-155); it yields no features. */
typedef struct {
    uint64_t n_structs;
    const uint64_t *row_offsets; /* n_structs + 1 */
    const float *n_xyz;
    const float *ca_xyz;
    const float *cb_xyz;
    const uint8_t *aa;
    const uint8_t *cb_valid;
} fd_struct_batch;

/* Encodings of the reference's `--type` (HashType, src/geometry/core.rs:10-60): the values of fd_hash_params.hash_type
 * are the reference's HashType index + 1, 0 selects the default. */
#define FD_HASH_DEFAULT 0          /* = PDBTrRosetta */
#define FD_HASH_PDBMOTIF 1         /* src/geometry/pdb_motif.rs */
#define FD_HASH_PDBMOTIFSINCOS 2   /* src/geometry/pdb_motif_sincos.rs */
#define FD_HASH_TRROSETTA 3        /* src/geometry/trrosetta.rs */
#define FD_HASH_PDBTRROSETTA 4     /* src/geometry/pdb_tr.rs */
#define FD_HASH_POINTPAIRFEATURE 5 /* src/geometry/ppf.rs */
#define FD_HASH_TERTIARYINTERACTION 6 /* src/geometry/tertiary_interaction.rs (CA trace around i and j) */
#define FD_HASH_HYBRID 7           /* src/geometry/hybrid.rs (amino-acid groups + PDBTrRosetta geometry + CA-trace torsions) */
#define FD_HASH_FOLDDISCOANGLE 8   /* src/geometry/folddisco_angle.rs */
#define FD_HASH_FOLDDISCODIST 9    /* src/geometry/folddisco_dist.rs */
#define FD_MAX_MULTIPLE_BINS 8

/* The hashing parameters of one index (IndexConfig, src/cli/config.rs:9-19).  nbin_dist / nbin_angle: if either is
 * 0 the encoding's defaults are used for both (feature.rs:215-221; PDBTrRosetta 16 / 4); larger values clamp the way
 * each encoding's perfect_hash does.  n_multiple_bins > 0 = `--multiple-bins`: every residue pair is hashed once per
 * (multiple_bins[2k], multiple_bins[2k+1]) = (dist, angle) pair instead (feature.rs:210-214).  A zero-initialised tail
 * (hash_type, n_multiple_bins) is the default PDBTrRosetta single-bin index. */
typedef struct {
    uint32_t nbin_dist;
    uint32_t nbin_angle;
    float dist_cutoff; /* IndexConfig.grid_width, default 20.0 (src/cli/main.rs:41) */
    uint32_t hash_type; /* FD_HASH_* */
    uint32_t n_multiple_bins;
    uint32_t multiple_bins[16]; /* 2 * FD_MAX_MULTIPLE_BINS: dist, angle, dist, angle ... */
} fd_hash_params;

/* ---- (i) index build ----------------------------------------------------------------------------- */
/* Replaces get_geometric_hash_as_u32_from_structure(...) + sort_unstable + dedup at
 * src/controller/mod.rs:334-344 (and :414-421) for a whole batch: per structure the sorted unique hashes.
 * Outputs are library-allocated: (*out_hashes)[(*out_row_offsets)[s] .. (*out_row_offsets)[s+1]). */
int fd_hash_structures(fd_ctx *ctx, const fd_struct_batch *batch, const fd_hash_params *params,
                       uint32_t **out_hashes, uint64_t **out_row_offsets);

/* The three arrays of the reference's on-disk index, byte-identical to what
 * FolddiscoIndex::save_offset_to_file / wrapup_offset_and_save_entries write
 * (src/index/indextable.rs:239-264, 297-326):
 *   PREFIX.offset = u64 count | u32 hashes[count] | u64 offsets[count+1];   PREFIX = values[value_bytes] */
typedef struct {
    uint64_t count;
    uint32_t *hashes;
    uint64_t *offsets;
    uint64_t value_bytes;
    uint8_t *values;
} fd_index_buffers;

/* Replaces count_single_entry / allocate_entries / add_single_entry / prune_to_sparse
 * (src/index/indextable.rs:88-105, 171-237, 267-295) given per-structure sorted unique hashes;
 * structure ids are first_id + row number. */
int fd_build_postings(fd_ctx *ctx, const uint32_t *hashes, const uint64_t *row_offsets, uint64_t n_structs,
                      uint64_t first_id, fd_index_buffers *out);
/* Fused: Folddisco::collect_and_count + allocate_entries + add_entries (src/controller/mod.rs:274-441) for a
 * batch resident in host memory; hashes never leave the device.  Only hashes in [hash_lo, hash_hi) are kept
 * (hash-range shard of a multi-GPU build; pass 0, 0xFFFFFFFF+1 = 1<<32 for everything). */
int fd_build_index(fd_ctx *ctx, const fd_struct_batch *batch, const fd_hash_params *params, uint64_t first_id,
                   uint64_t hash_lo, uint64_t hash_hi, fd_index_buffers *out);
void fd_free_index_buffers(fd_index_buffers *b);

/* ---- (ii) query ---------------------------------------------------------------------------------- */
/* Replaces load_folddisco_index (src/index/indextable.rs:331-394) + load_lookup_from_file: copies the index
 * (host buffers may be the mmaps of PREFIX.offset / PREFIX) to the device and builds the device-side
 * directory, skip table and per-list posting counts.  nres / plddt are the .lookup columns
 * (src/index/lookup.rs:77-91); ids in the postings are positions in these arrays. */
int fd_index_attach(fd_ctx *ctx, const uint32_t *hashes, const uint64_t *offsets, uint64_t count,
                    const uint8_t *values, uint64_t value_bytes, uint64_t n_structs, const uint32_t *nres,
                    const float *plddt);
/* FolddiscoIndex::get_entries(hash).len() (src/index/indextable.rs:83-86) for n hashes; 0 if absent. */
int fd_posting_counts(fd_ctx *ctx, const uint32_t *hashes, uint64_t n, uint32_t *out_counts);
/* Decode one posting list (debug / parity helper).  Library-allocated *out_ids. */
int fd_get_entries(fd_ctx *ctx, uint32_t hash, uint64_t **out_ids, uint64_t *out_n);

/* One query = the key set of make_query_map's hash map (src/controller/query.rs:208-329) with, per hash, the
 * query edge it came from.  edge_of_hash[h] indexes the query's distinct (i, j) edges; edge_node[e] is a dense
 * id of the edge's source residue i (the node group of src/controller/count_query.rs:256-273). */
typedef struct {
    uint32_t n_hashes;
    const uint32_t *hashes;
    const uint16_t *edge_of_hash;
    uint32_t n_edges;
    const uint16_t *edge_node;
    uint32_t n_nodes;
    uint32_t expected_node_count; /* residue_count of src/cli/workflows/query_pdb.rs:355-359 */
    /* Optional (NULL = every entry of edge_node is its own query edge).  edge_group[e] names the query edge that
     * vote bit e belongs to; equal values must be consecutive and must not cross a multiple of 32.  Bits of one
     * group count as ONE edge in edge_count.  Used with hash-range shards: an edge whose hashes live on several
     * ranks gets one bit per rank, so that the ranks' edge masks are disjoint and a sum merges them exactly. */
    const uint16_t *edge_group;
} fd_query;

/* count_query arguments (src/controller/count_query.rs:82-88) + StructureFilter::filter_before_matching
 * (src/controller/filter.rs:76-100) + --top (query_pdb.rs:404-411).  Negative sampling/freq = None. */
typedef struct {
    float sampling_ratio;
    int64_t sampling_count;
    float freq_filter;
    float length_penalty; /* default 0.5 */
    uint64_t total_match_count;
    uint64_t covered_node_count;
    float covered_node_ratio;
    float idf_score_cutoff;
    uint64_t num_res_cutoff; /* default 50000 */
    float plddt_cutoff;
    uint64_t top_n; /* UINT64_MAX = all */
} fd_prefilter_params;

typedef struct {
    uint32_t nid;
    uint32_t match_count;
    uint32_t node_count;
    uint32_t edge_count;
    float idf;
} fd_struct_hit;

/* Replaces count_query(...) (src/cli/workflows/query_pdb.rs:391-394) followed by filter_before_matching,
 * the stable sort by idf descending and truncate(top_n) (:395-411) for a batch of queries.
 * Library-allocated: hits of query q are (*out_hits)[(*out_offsets)[q] .. (*out_offsets)[q+1]), ordered by
 * idf descending, ties by ascending nid. */
int fd_count_query_batch(fd_ctx *ctx, const fd_query *queries, uint32_t n_queries,
                         const fd_prefilter_params *params, fd_struct_hit **out_hits, uint64_t **out_offsets);
/* fd_count_query_batch for a batch the caller may search again: batch_id != 0 is the caller's promise that queries and
 * params are identical to the previous call on ctx that carried the same id (and the attached index is the same; a new
 * fd_index_attach invalidates the cache).  The library then reuses the uploaded batch and its lookup results instead
 * of flattening, uploading and looking the hashes up again.  batch_id == 0: no caching (= fd_count_query_batch). */
int fd_count_query_batch_id(fd_ctx *ctx, const fd_query *queries, uint32_t n_queries,
                            const fd_prefilter_params *params, uint64_t batch_id, fd_struct_hit **out_hits,
                            uint64_t **out_offsets);
/* The same for one ID-RANGE SHARD of a larger database (multi-GPU, SURVEY 8e ablation / fd_search_sharded): the
 * attached index holds the postings of a contiguous range of structure ids (local ids 0 .. n_structs) and
 * global_counts[k] is the length of the k-th query hash's list in the WHOLE database (hashes of all queries
 * concatenated in batch order; the sum over shards of fd_posting_counts), global_n_structs its structure count.  Every
 * shard then weighs a hash by log2(N / len) exactly as the unsharded index does (count_query.rs:130) and uses the same
 * fixed-point scale, so a structure's row is bit-identical to the unsharded one; the caller merges the shards'
 * per-query top-n lists (idf descending, id ascending) after adding each shard's first id.  global_counts == NULL
 * is fd_count_query_batch. */
int fd_count_query_batch_ex(fd_ctx *ctx, const fd_query *queries, uint32_t n_queries,
                            const fd_prefilter_params *params, const uint32_t *global_counts,
                            uint64_t global_n_structs, fd_struct_hit **out_hits, uint64_t **out_offsets);
/* ---- multi-GPU: NCCL inside the library + id-range shards ---------------------------------------------
 * One context per rank / GPU.  Rank 0 obtains a unique id (fd_comm_unique_id) and hands its FD_COMM_ID_BYTES bytes to
 * the other ranks by any means (a file, MPI, torch.distributed ...); every rank then calls fd_comm_init.  The
 * collectives below run on the context's stream; the *_host helpers stage host buffers through device memory. */
#define FD_COMM_ID_BYTES 128
int fd_comm_unique_id(uint8_t *out_id /* FD_COMM_ID_BYTES */);
int fd_comm_init(fd_ctx *ctx, const uint8_t *id /* FD_COMM_ID_BYTES */, int rank, int world);
void fd_comm_destroy(fd_ctx *ctx); /* also done by fd_destroy */
int fd_comm_rank(const fd_ctx *ctx);
int fd_comm_world(const fd_ctx *ctx);
int fd_comm_allgather(fd_ctx *ctx, const void *send, uint64_t bytes, void *recv /* world * bytes */);
int fd_comm_allreduce_u32(fd_ctx *ctx, uint32_t *inout, uint64_t n); /* sum */
int fd_comm_barrier(fd_ctx *ctx);

/* count_query over a database whose structures are split into contiguous ID RANGES, one per rank (the
 * reference's merge is per structure id, src/controller/count_query.rs:172-217, and its top-n a sort + truncate,
 * src/cli/workflows/query_pdb.rs:404-411, so id-range shards need only a top-n exchange).  Every rank passes the SAME
 * batch (all ranks' queries in the same order) and the global list lengths (all-reduced fd_posting_counts); rank r owns
 * the queries [slice_begin[r], slice_begin[r + 1]).  The rank scans its local index (ids 0 .. n_structs, reported as
 * first_id + id) for the whole batch, keeps its top_n per query, ONE all-to-all of fixed-size blocks (ncclSend /
 * ncclRecv group) sends every query's block to its owner, the owner merges (idf descending, id ascending) and returns
 * the global top_n of ITS OWN queries: *out_offsets has slice_begin[r + 1] - slice_begin[r] + 1 entries.  Rows are
 * bit-identical to fd_count_query_batch on the unsharded index.  top_n must be in [1, 4096 / world]; every rank must
 * hold at least one structure. */
int fd_count_query_sharded(fd_ctx *ctx, const fd_query *queries, uint32_t n_queries, const fd_prefilter_params *params,
                           const uint32_t *global_counts, uint64_t global_n_structs, uint64_t first_id,
                           const uint32_t *slice_begin /* world + 1 */, fd_struct_hit **out_hits,
                           uint64_t **out_offsets);
/* the same with a batch id (see fd_count_query_batch_id): every rank scans the whole batch of all ranks, so the
 * flatten + upload + lookup saved per repeated batch grows with the number of ranks */
int fd_count_query_sharded_id(fd_ctx *ctx, const fd_query *queries, uint32_t n_queries, const fd_prefilter_params *params,
                              const uint32_t *global_counts, uint64_t global_n_structs, uint64_t first_id,
                              const uint32_t *slice_begin /* world + 1 */, uint64_t batch_id, fd_struct_hit **out_hits,
                              uint64_t **out_offsets);
/* bytes this rank sent to other ranks in the last fd_count_query_sharded */
uint64_t fd_last_exchange_bytes(const fd_ctx *ctx);

/* posting bytes the last fd_count_query_batch had to read (sum over found query hashes of their list
 * length in bytes) -- the algorithmic bytes of SURVEY 8d */
uint64_t fd_last_posting_bytes(const fd_ctx *ctx);

/* ---- multi-GPU: hash-range shards ------------------------------------------------------------------
 * With the index split by hash range across ranks (fd_build_index(hash_lo, hash_hi) or a slice of the on-disk
 * arrays), count_query's per-structure accumulators (src/controller/count_query.rs:60-67 CompactEntry, :133-159)
 * are additive over shards: every rank produces PARTIAL votes for the whole batch in device memory, the caller
 * merges them with ONE sum all-reduce over the whole buffer (the edge planes are bit masks: the caller gives every
 * (query edge, owning rank) pair its own bit through fd_query.edge_group, so the ranks' masks are disjoint and the
 * sum is their OR), and fd_votes_select finishes count_query (node/edge counts, length penalty,
 * filter_before_matching, idf sort, --top) for a slice of the batch.
 *
 * Device layout: u32 votes[planes][n_queries][n_structs];
 *   narrow = 1: plane 0 = match_count << 24 | idf fixed point (24 bit)         (queries of at most 255 hashes)
 *   narrow = 0: plane 0 = idf fixed point (32 bit), plane 1 = match_count
 *   then edge_words planes of edge bitmask (bit e of word e/32 = query edge e has a hit in the structure).
 * The idf fixed-point scale of query q is 2^floor(log2(range / (n_hashes_q * log2(n_structs) + 1))): identical on
 * every rank. */
typedef struct {
    uint32_t n_queries;
    uint32_t n_structs;
    uint32_t narrow;
    uint32_t edge_words;
    uint32_t planes;
    uint32_t first_query; /* the buffer holds queries first_query .. first_query + n_queries of the batch */
    uint64_t words;       /* planes * n_queries * n_structs (dense); record words (sparse) */
} fd_votes_layout;

/* Scan this rank's shard for the whole batch.  *d_votes is a DEVICE pointer owned by ctx (valid until the next
 * fd_votes_scan / fd_destroy); the call returns after the scan has completed on the device. */
int fd_votes_scan(fd_ctx *ctx, const fd_query *queries, uint32_t n_queries, const fd_prefilter_params *params,
                  fd_votes_layout *layout, uint32_t **d_votes);
/* Sparse form of the same exchange (the default of folddisco_b200/sharded.py): only NON-EMPTY cells travel.
 * slice_begin[world + 1] assigns the queries [slice_begin[r], slice_begin[r+1]) to rank r, which will finish them.
 * The scan packs every non-empty (query, structure) cell of this rank's shard as a record of 1 + planes words
 * {key = (query - slice_begin[dest]) * n_structs + nid, plane words...} into the region of its destination:
 * records of rank r are (*d_records)[(region_offset[r] + k) * (1 + planes)], k < region_count[r]  (device memory
 * owned by ctx; region_offset has world + 1 entries, region_count world).  The caller moves the regions with one
 * all-to-all; the receiver clears a dense buffer for its slice (fd_votes_merge_begin), adds its own region and
 * every received one (fd_votes_apply: plane 0[/1] add, edge planes OR) and finishes with fd_votes_select. */
int fd_votes_scan_sparse(fd_ctx *ctx, const fd_query *queries, uint32_t n_queries, const fd_prefilter_params *params,
                         const uint32_t *slice_begin, uint32_t world, fd_votes_layout *layout, uint32_t **d_records,
                         uint64_t *region_offset, uint64_t *region_count);
int fd_votes_merge_begin(fd_ctx *ctx, const fd_votes_layout *layout, uint32_t first_query, uint32_t n_queries,
                         fd_votes_layout *slice_layout, uint32_t **d_dense);
int fd_votes_apply(fd_ctx *ctx, const fd_votes_layout *slice_layout, uint32_t *d_dense, const uint32_t *d_records,
                   uint64_t n_records);

/* Finish count_query from merged votes (any device pointer with the layout above) for queries
 * [q_begin, q_end) of the batch; outputs as fd_count_query_batch, offsets has q_end - q_begin + 1 entries. */
int fd_votes_select(fd_ctx *ctx, const fd_query *queries, uint32_t n_queries, const fd_prefilter_params *params,
                    const fd_votes_layout *layout, const uint32_t *d_votes, uint32_t q_begin, uint32_t q_end,
                    fd_struct_hit **out_hits, uint64_t **out_offsets);

/* HBM-resident compact-structure store replacing the per-candidate file re-read of retrieval_wrapper
 * (src/controller/retrieve.rs:375-376); ids = positions in the batch = posting ids. */
int fd_store_attach(fd_ctx *ctx, const fd_struct_batch *batch);

/* Optional: the PAIR TABLE of the attached store -- for every structure the residue pairs (i, j) that get a geometric
 * hash, sorted by hash (8 bytes per pair, about 115 pairs per residue), i.e. exactly what
 * get_geometric_hash_as_u32_from_structure (src/controller/feature.rs:198-231) enumerates, kept instead of thrown
 * away.  With it fd_verify_candidates_* looks every query hash up in the candidate (directory + short binary search)
 * instead of re-screening and re-hashing the candidate's residue pairs (retrieve.rs:85-143 does that per candidate);
 * results are identical.  params must be the hash parameters of the searches that follow (otherwise the table is
 * ignored).  max_bytes > 0: fail with FD_ERR_LIMIT (and keep the re-hash path) when the table would be larger;
 * *out_bytes = its size (may be NULL). */
int fd_store_build_pair_table(fd_ctx *ctx, const fd_hash_params *params, uint64_t max_bytes, uint64_t *out_bytes);

/* Per-query inputs of retrieve_with_prefilter (src/controller/retrieve.rs:52-156). */
typedef struct {
    uint32_t n_hashes;
    const uint32_t *hashes_sorted; /* the query hash set, ascending */
    uint32_t n_aa_dist;            /* observed_distance_map flattened (query.rs:271-280), map-insertion order */
    const uint8_t *aa1;
    const uint8_t *aa2;
    const float *ca_dist;
    const uint32_t *q_index; /* query residue index i of the entry */
} fd_retrieval_query;

typedef struct {
    uint32_t cand; /* index into the candidate arrays */
    uint32_t i, j; /* target residue indices */
    uint32_t hash;
} fd_cand_edge;
typedef struct {
    uint32_t cand;
    uint32_t q_index;
    uint32_t i, j;
    uint32_t k; /* position of the aa_dist entry inside its (aa1,aa2) list: candidate_pairs order */
} fd_cand_pair;

/* Replaces prefilter_amino_acid + retrieve_with_prefilter (src/controller/retrieve.rs:563-602, 52-156) for
 * n_cand (query, target) candidates against the attached store.  Library-allocated outputs, sorted by
 * (cand, i, j[, k]) which is the reference's emission order. */
int fd_candidate_edges_batch(fd_ctx *ctx, const fd_retrieval_query *queries, uint32_t n_queries,
                             const uint32_t *cand_query, const uint32_t *cand_nid, uint64_t n_cand,
                             const fd_hash_params *params, float ca_dist_cutoff, fd_cand_edge **out_edges,
                             uint64_t *out_n_edges, fd_cand_pair **out_pairs, uint64_t *out_n_pairs);

/* Replaces KabschSuperimposer::run -> kabsch(x, y, mode 2) (src/structure/kabsch.rs:73-95, 157-554) for a
 * batch: alignment a rotates mov[pt_offsets[a]..pt_offsets[a+1]) onto ref[...]; f64 inside, f32 out.
 * rmsd[n], U9[9n] row-major, t3[3n] are caller-allocated. */
int fd_kabsch_batch(fd_ctx *ctx, const float *mov_xyz, const float *ref_xyz, const uint32_t *pt_offsets,
                    uint32_t n_align, float *rmsd, float *U9, float *t3);

/* rmsd_with_calpha_and_rottran (src/controller/retrieve.rs:756-834) for a batch, without moving target
 * coordinates through the host: alignment a superposes, for k in [pair_offsets[a], pair_offsets[a+1]), CA and CB
 * of residue pair_tres[k] of stored structure align_nid[a] onto CA and CB of query residue pair_qres[k]
 * (an index into q_ca_xyz / q_cb_xyz, the concatenated query structures, n_q_res residues). */
int fd_kabsch_store_batch(fd_ctx *ctx, const float *q_ca_xyz, const float *q_cb_xyz, uint64_t n_q_res,
                          const uint32_t *align_nid, const uint32_t *pair_offsets, uint32_t n_align,
                          const uint32_t *pair_qres, const uint32_t *pair_tres, float *rmsd, float *U9, float *t3);

/* `--partial-fit`: rmsd_with_calpha_and_rottran with lms = true (src/controller/retrieve.rs:773-814) for a
 * batch.  Same arguments as fd_kabsch_store_batch.  An alignment of more than three residues is superposed by LMS-QCP
 * (LmsQcpSuperimposer::run with its default parameters, src/structure/lms_qcp.rs:28-40, 91-236: 500 deterministic seed
 * triples scored by their median residual, forward growth of the inlier core up to 2 A) and rmsd[a] is the RMSD over the
 * core; three residues or fewer use Kabsch.  At most 64 residues per alignment (FD_ERR_LIMIT). */
int fd_lmsqcp_store_batch(fd_ctx *ctx, const float *q_ca_xyz, const float *q_cb_xyz, uint64_t n_q_res,
                          const uint32_t *align_nid, const uint32_t *pair_offsets, uint32_t n_align,
                          const uint32_t *pair_qres, const uint32_t *pair_tres, float *rmsd, float *U9, float *t3);
/* LMS-QCP by the host build of csrc/fd_lmsqcp.cuh over explicit point lists (parity probe, no device): ref = query
 * points, mov = target points; 0, or -1 when n_points is outside 3 .. 128 */
int fd_lmsqcp_host(const float *ref_xyz, const float *mov_xyz, uint32_t n_points, float *U9, float *t3,
                   float *rms_inliers);

/* Similarity metrics of verified matches: StructureSimilarityMetrics::calculate_all as
 * rmsd_with_calpha_and_rottran evaluates it for every match (src/controller/retrieve.rs:776-831,
 * src/structure/metrics.rs:44-345; the columns tm_score, gdt_ts, gdt_ha, chamfer_distance, hausdorff_distance of
 * src/controller/result.rs:280-284 and the MatchFilter cutoffs of src/controller/filter.rs:217-236).  Same alignment
 * description as fd_kabsch_store_batch plus the superposition (U9 row-major, t3) it returned; out_metrics[5 * a] =
 * {tm_score, gdt_ts, gdt_ha, chamfer_distance, hausdorff_distance} of alignment a.  One thread per match on the device;
 * the reference's distance-vs-squared-distance quirk of TM / GDT is reproduced. */
int fd_metrics_store_batch(fd_ctx *ctx, const float *q_ca_xyz, const float *q_cb_xyz, uint64_t n_q_res,
                           const uint32_t *align_nid, const uint32_t *pair_offsets, uint32_t n_align,
                           const uint32_t *pair_qres, const uint32_t *pair_tres, const float *U9, const float *t3,
                           float *out_metrics);
/* the same arithmetic by the host build of csrc/fd_metrics.cuh over explicit point lists (parity probe, no device):
 * ref = query points, mov = target points, n_points each */
void fd_metrics_host(const float *ref_xyz, const float *mov_xyz, uint32_t n_points, const float *U9, const float *t3,
                     float *out5);
/* One verified match = one connected component of one candidate (a row of retrieval_wrapper's result vector,
 * src/controller/retrieve.rs:364-552).  res[k] = 1 + index of the target residue matched to the k-th query
 * residue, 0 = none ("_"). */
typedef struct {
    uint32_t cand;
    uint32_t node_count;
    float idf;  /* calculate_subgraph_idf (retrieve.rs:705-719) */
    float rmsd; /* rmsd_with_calpha_and_rottran (retrieve.rs:756-834) */
    float U[9];
    float t[3];
    uint32_t res[16];
} fd_match_record;

/* Per-query inputs of retrieval_wrapper: the query map (sorted by hash) with, per hash, its edge (i, j) as
 * residue indices of the query structure, its idf and is_symmetric() (src/geometry/pdb_tr.rs:158-162); the
 * observed distance map; the query residue list (query_indices, make_query_map's second output); and the query
 * structure's CA / CB coordinates. */
typedef struct {
    uint32_t n_hashes;
    const uint32_t *hashes_sorted;
    const uint32_t *hash_qi;
    const uint32_t *hash_qj;
    const float *hash_idf;
    const uint8_t *hash_symmetric;
    uint32_t n_aa_dist;
    const uint8_t *aa1;
    const uint8_t *aa2;
    const float *ca_dist;
    const uint32_t *q_index;
    uint32_t n_indices;
    const uint32_t *indices;
    uint32_t n_residues;
    const float *ca_xyz;
    const float *cb_xyz;
} fd_verify_query;

/* Fused replacement of retrieval_wrapper(...) (src/cli/workflows/query_pdb.rs:425-447, src/controller/
 * retrieve.rs:364-552) for n_cand (query, candidate) pairs against the attached store: candidate re-hash,
 * graph components, residue mapping + rescue and Kabsch RMSD on the device (three kernels, nothing but match records leaves HBM).  Library-allocated outputs:
 * records grouped by candidate in component order; flags[c] != 0 marks a candidate that exceeds the kernel's
 * shared-memory limits (more than 256 matching edges, 64 graph nodes or 16 query residues) and
 * must be verified with fd_candidate_edges_batch + fd_kabsch_store_batch instead (no records are emitted
 * for it). */
int fd_verify_candidates_batch(fd_ctx *ctx, const fd_verify_query *queries, uint32_t n_queries,
                               const uint32_t *cand_query, const uint32_t *cand_nid, uint64_t n_cand,
                               const fd_hash_params *params, float ca_dist_cutoff, int skip_ca_match,
                               fd_match_record **out_records, uint64_t *out_n, uint8_t **out_flags);

/* The same computation without the output copies: *out_records (all records, grouped by candidate in component
 * order), *out_first (n_cand + 1 offsets: candidate c owns records [first[c], first[c+1])) and *out_flags are views of
 * pinned staging buffers owned by ctx, valid until the next verification call on ctx.  For hosts that consume the
 * records immediately (the in-library host does). */
int fd_verify_candidates_view(fd_ctx *ctx, const fd_verify_query *queries, uint32_t n_queries,
                              const uint32_t *cand_query, const uint32_t *cand_nid, uint64_t n_cand,
                              const fd_hash_params *params, float ca_dist_cutoff, int skip_ca_match,
                              const fd_match_record **out_records, uint64_t *out_n, const uint32_t **out_first,
                              const uint8_t **out_flags);

/* The query side of the verification is independent of the candidates (the reference rebuilds it per query inside
 * retrieval_wrapper, retrieve.rs:364-452: hash set, observed_distance_map, query residues).  fd_verify_prepare
 * flattens it once per query batch and keeps it resident on ctx's device; fd_verify_candidates_prepared is
 * fd_verify_candidates_view on those tables (cand_query indexes the prepared batch).  Release with
 * fd_verify_prepared_free before fd_destroy(ctx).  fd_verify_prepared_bytes: bytes uploaded by the prepare call. */
typedef struct fd_verify_prepared fd_verify_prepared;
int fd_verify_prepare(fd_ctx *ctx, const fd_verify_query *queries, uint32_t n_queries, fd_verify_prepared **out);
void fd_verify_prepared_free(fd_verify_prepared *p);
uint64_t fd_verify_prepared_bytes(const fd_verify_prepared *p);
int fd_verify_candidates_prepared(fd_ctx *ctx, const fd_verify_prepared *prepared, const uint32_t *cand_query,
                                  const uint32_t *cand_nid, uint64_t n_cand, const fd_hash_params *params,
                                  float ca_dist_cutoff, int skip_ca_match, const fd_match_record **out_records,
                                  uint64_t *out_n, const uint32_t **out_first, const uint8_t **out_flags);

/* ---- result rows on the device ------------------------------------------------------------------------
 * The rows a search hands out (folddisco_b200_host.h names them fdh_struct_row / fdh_match_row /
 * fdh_residue_match): per-candidate summary (src/controller/retrieve.rs:539-551), per-match rows, matched residues with
 * their (chain, residue number) labels, and the default sort orders (src/controller/sort.rs:218-222, 454-458). */
typedef struct {
    uint32_t nid, total_match_count, node_count, edge_count;
    float idf;
    uint32_t max_matching_node_count;
    float min_rmsd_with_max_match;
    uint64_t match_begin, match_end; /* this structure's matches in match-emission order (unsorted view) */
} fd_struct_row;
typedef struct {
    uint32_t nid;
    uint32_t node_count;
    float idf;
    float rmsd;
    float U[9];
    float t[3];
    uint64_t res_begin; /* n_query_residues entries in the residue arrays */
} fd_match_row;
typedef struct {
    uint8_t some;
    uint8_t chain;
    uint64_t serial;
} fd_residue_row;
/* (chain, residue number) of every residue of the attached structure store, for fd_verify_rows (optional: without
 * them a matched residue is reported as chain 0, number = residue index). */
int fd_store_attach_labels(fd_ctx *ctx, const uint8_t *chain, const uint64_t *serial, uint64_t n_residues);
int fd_store_has_labels(const fd_ctx *ctx); /* 1 after fd_store_attach_labels on the current store */
/* fd_verify_candidates_prepared that leaves the match records ON THE DEVICE (no record copy): *out_n records,
 * *out_first / *out_flags as above.  The records stay until the next verification call on ctx, fd_verify_rows or
 * fd_verify_records_fetch. */
int fd_verify_candidates_device(fd_ctx *ctx, const fd_verify_prepared *prepared, const uint32_t *cand_query,
                                const uint32_t *cand_nid, uint64_t n_cand, const fd_hash_params *params,
                                float ca_dist_cutoff, int skip_ca_match, uint64_t *out_n, const uint32_t **out_first,
                                const uint8_t **out_flags);
/* the kept records copied to pinned staging owned by ctx (the path of hosts that assemble rows themselves) */
int fd_verify_records_fetch(fd_ctx *ctx, const fd_match_record **out_records);
/* Row assembly on the device for the kept records: query q owns candidates [cand_offsets[q], cand_offsets[q+1]) in
 * count_query order, hits[c] is the count_query row of candidate c, res_offsets[q] the first residue row of query q
 * (res_offsets[q+1] - res_offsets[q] = matches of q x query residues of q).  Outputs are caller-allocated host
 * memory (pinned memory makes the copies asynchronous): structs[n_cand] sorted per query (idf desc, min_rmsd asc,
 * stable), matches[n_records] in emission order, match_order[n_records] (sorted position -> emission index: idf desc,
 * rmsd asc, stable), residues[res_offsets[n_queries]].  needs_host_sort[q] = 1: the query's orders were left as
 * emitted (more than 2048 matches or candidates, NaN / -0.0 keys) and the caller sorts them. */
typedef struct {
    uint32_t n_queries;
    const uint64_t *cand_offsets;
    const fd_struct_hit *hits;
    const uint64_t *res_offsets;
    fd_struct_row *structs;
    fd_match_row *matches;
    uint64_t *match_order;
    fd_residue_row *residues;
    uint8_t *needs_host_sort;
} fd_rows_request;
int fd_verify_rows(fd_ctx *ctx, const fd_rows_request *request);
/* The two steps as ONE pipelined call: the candidates are verified in a few chunks cut at query boundaries, and the rows
 * of a finished chunk are assembled and copied to the host while the next chunks are still being verified (the copy of
 * the finished rows is the longest piece of a search once everything else runs on the device).  The caller provides the
 * output arrays with capacities (fd_verify_match_capacity rows for matches / match_order, that many times the largest
 * n_res for residues) and n_res[q], the number of query residues of query q.  On return `done` tells whether the rows
 * were delivered; 0 (a candidate needs the general verification path, or a capacity was too small): the records are
 * kept on the device as after fd_verify_candidates_device, and the caller uses fd_verify_rows / fd_verify_records_fetch.
 * res_offsets[n_queries + 1] is filled like fd_rows_request.res_offsets. */
typedef struct {
    uint32_t n_queries;
    const uint64_t *cand_offsets;
    const fd_struct_hit *hits;
    const uint32_t *n_res;
    uint64_t match_capacity, residue_capacity;
    fd_struct_row *structs;
    fd_match_row *matches;
    uint64_t *match_order;
    fd_residue_row *residues;
    uint8_t *needs_host_sort;
    uint64_t *res_offsets;
    int done;
} fd_rows_plan;
uint64_t fd_verify_match_capacity(const fd_ctx *ctx, uint64_t n_cand);
int fd_verify_candidates_rows(fd_ctx *ctx, const fd_verify_prepared *prepared, const uint32_t *cand_query,
                              const uint32_t *cand_nid, uint64_t n_cand, const fd_hash_params *params,
                              float ca_dist_cutoff, int skip_ca_match, fd_rows_plan *plan, uint64_t *out_n,
                              const uint32_t **out_first, const uint8_t **out_flags);

/* number of structures of the attached index (lookup.len()); 0 if none */
uint64_t fd_index_num_structs(const fd_ctx *ctx);

/* ---- parity / debug probes ------------------------------------------------------------------------- */
/* op: 0 sin, 1 cos, 2 acos, 3 atan2(a, b); evaluates the device math used by the hash kernels */
int fd_math_probe(fd_ctx *ctx, int op, const float *a, const float *b, uint64_t n, float *out);
/* the geometric hash of residue pairs (i, j) of one SoA structure by the host build of csrc/fd_geom.cuh: exact route
 * (binary64 trigonometry), fast route (closed forms + margin test; 0 where it declined) and the declined flags.  The
 * kernels use fast-then-exact (pair_hash_auto); this probe is how the tests prove fast == exact wherever it answers. */
void fd_pair_hash_host(const float *n_xyz, const float *ca_xyz, const float *cb_xyz, const uint8_t *aa,
                       const uint32_t *pair_i, const uint32_t *pair_j, uint64_t n_pairs, const fd_hash_params *params,
                       uint32_t *out_exact, uint32_t *out_fast, uint8_t *out_declined);
/* Other encodings (fd_hash_params.hash_type / multiple_bins): every ordered residue pair (i, j), i != j, of one SoA
 * structure hashed on the HOST by the header the kernels compile (csrc/fd_hashtypes.cuh), pairs in row-major order and
 * the bin pairs of one residue pair in list order -- get_geometric_hash_as_u32_from_structure before its sort + dedup
 * (src/controller/feature.rs:198-231).  Returns the number of hashes (writes at most cap), -1 for refused parameters.
 * fd_typed_is_symmetric_host: HashValue::is_symmetric of the encoding (1 / 0; -1 = encoding not built). */
int64_t fd_typed_hash_host(const float *n_xyz, const float *ca_xyz, const float *cb_xyz, const uint8_t *aa,
                           const uint8_t *cb_valid, uint64_t n_res, const fd_hash_params *params, uint32_t *out,
                           int64_t cap);
int fd_typed_is_symmetric_host(uint32_t hash_type, uint32_t hash);
/* same functions evaluated by the host build of the same header (no device needed) */
void fd_math_host(int op, const float *a, const float *b, uint64_t n, float *out);

#ifdef __cplusplus
}
#endif
#endif /* FOLDDISCO_B200_H */
