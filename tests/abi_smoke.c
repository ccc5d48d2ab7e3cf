/* abi_smoke.c -- the C ABI exercised from plain C (TEST INFRASTRUCTURE).
 *
 * Compiled with gcc -std=c11 against include/folddisco_b200.h and include/folddisco_b200_host.h and linked with
 * libfolddisco_b200.so, so that
 *   - the headers are valid C (no C++ leaks into the boundary),
 *   - every struct a foreign-language binding mirrors has the size / field offsets asserted below (the Rust block of
 *     INTEGRATION.md and folddisco_b200/capi.py are hand-written mirrors: a header change that they miss fails here),
 *   - every entry point named below exists in the library (the link step),
 *   - and, on a GPU box (argv[1] = a file written by tests/test_abi.py), the call sequence a host makes runs end to
 *     end: fd_create -> store -> fd_build_index (K1 + K2) -> fd_index_attach -> fd_store_attach (+ pair table) ->
 *     fdh_queries_add / finalize -> fdh_search (K3 + K6) -> rows -> free.
 * Without a GPU fd_create must fail loudly (no CPU fallback): exit code 3.
 *
 * Input file (little endian): u64 n_structs | u64 row_offsets[n+1] | f32 n_xyz[3R] | f32 ca_xyz[3R] | f32 cb_xyz[3R] |
 *   u8 aa[R] | u8 cb_valid[R] | u8 chain[R] | u64 serial[R] | u64 query_struct | u64 qlen | char query[qlen].
 * Output: one line per structure row "S nid match node edge idf" and per match row "M nid node idf rmsd".
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "folddisco_b200.h"
#include "folddisco_b200_host.h"

_Static_assert(sizeof(fd_struct_batch) == 56, "fd_struct_batch");
_Static_assert(sizeof(fd_hash_params) == 84, "fd_hash_params");
_Static_assert(offsetof(fd_hash_params, hash_type) == 12 && offsetof(fd_hash_params, multiple_bins) == 20, "fd_hash_params tail");
_Static_assert(sizeof(fd_index_buffers) == 40, "fd_index_buffers");
_Static_assert(sizeof(fd_query) == 56, "fd_query");
_Static_assert(offsetof(fd_query, expected_node_count) == 44, "fd_query.expected_node_count");
_Static_assert(offsetof(fd_query, edge_group) == 48, "fd_query.edge_group");
_Static_assert(sizeof(fd_prefilter_params) == 72, "fd_prefilter_params");
_Static_assert(offsetof(fd_prefilter_params, top_n) == 64, "fd_prefilter_params.top_n");
_Static_assert(sizeof(fd_struct_hit) == 20, "fd_struct_hit");
_Static_assert(sizeof(fd_votes_layout) == 32, "fd_votes_layout");
_Static_assert(sizeof(fd_retrieval_query) == 56, "fd_retrieval_query");
_Static_assert(sizeof(fd_cand_edge) == 16, "fd_cand_edge");
_Static_assert(sizeof(fd_cand_pair) == 20, "fd_cand_pair");
_Static_assert(sizeof(fd_match_record) == 128, "fd_match_record");
_Static_assert(offsetof(fd_match_record, res) == 64, "fd_match_record.res");
_Static_assert(sizeof(fd_verify_query) == 128, "fd_verify_query");
_Static_assert(offsetof(fd_verify_query, cb_xyz) == 120, "fd_verify_query.cb_xyz");
_Static_assert(sizeof(fdh_query_params) == 120, "fdh_query_params");
_Static_assert(sizeof(fdh_search_params) == 128, "fdh_search_params");
_Static_assert(sizeof(fdh_struct_row) == 48, "fdh_struct_row");
_Static_assert(sizeof(fdh_match_row) == 72, "fdh_match_row");
_Static_assert(sizeof(fdh_residue_match) == 16, "fdh_residue_match");
_Static_assert(sizeof(fd_struct_row) == 48 && sizeof(fd_match_row) == 72 && sizeof(fd_residue_row) == 16, "device row layouts");
_Static_assert(sizeof(fd_rows_request) == 72, "fd_rows_request");
_Static_assert(sizeof(fd_rows_plan) == 104, "fd_rows_plan");
_Static_assert(FD_COMM_ID_BYTES == 128, "NCCL unique id");

/* entry points that the run below does not call: taking their address makes the link step check them */
static const void *const abi_symbols[] = {
    (const void *)fd_fork, (const void *)fd_fork_refresh, (const void *)fd_lane, (const void *)fd_hash_structures,
    (const void *)fd_build_postings, (const void *)fd_posting_counts, (const void *)fd_get_entries,
    (const void *)fd_count_query_batch, (const void *)fd_count_query_batch_ex, (const void *)fd_count_query_sharded,
    (const void *)fd_comm_unique_id, (const void *)fd_comm_init, (const void *)fd_comm_destroy,
    (const void *)fd_comm_allgather, (const void *)fd_comm_allreduce_u32, (const void *)fd_comm_barrier,
    (const void *)fd_votes_scan, (const void *)fd_votes_scan_sparse, (const void *)fd_votes_merge_begin,
    (const void *)fd_votes_apply, (const void *)fd_votes_select, (const void *)fd_candidate_edges_batch,
    (const void *)fd_kabsch_batch, (const void *)fd_kabsch_store_batch, (const void *)fd_verify_candidates_batch,
    (const void *)fd_verify_candidates_view, (const void *)fd_verify_prepare, (const void *)fd_verify_prepared_free,
    (const void *)fd_verify_candidates_prepared, (const void *)fd_store_build_pair_table,
    (const void *)fdh_index_save, (const void *)fdh_index_load, (const void *)fdh_store_save, (const void *)fdh_store_load,
    (const void *)fdh_queries_finalize_sharded, (const void *)fdh_search_sharded, (const void *)fdh_search_from_votes,
};

static void *slurp(FILE *f, size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (!p || fread(p, 1, bytes, f) != bytes) {
        fprintf(stderr, "abi_smoke: short input file\n");
        exit(2);
    }
    return p;
}

int main(int argc, char **argv) {
    (void)abi_symbols;
    printf("version %s\n", fd_version());
    fd_ctx *ctx = NULL;
    if (fd_create(&ctx, 0) != FD_OK) {
        fprintf(stderr, "fd_create: %s\n", fd_last_error(NULL));
        return 3; /* no CUDA device: the library refuses, it has no CPU path */
    }
    if (argc < 2) {
        fd_destroy(ctx);
        return 0;
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint64_t S = 0;
    if (fread(&S, 8, 1, f) != 1) return 2;
    uint64_t *ro = (uint64_t *)slurp(f, (S + 1) * 8);
    const uint64_t R = ro[S];
    float *nx = (float *)slurp(f, R * 12), *ca = (float *)slurp(f, R * 12), *cb = (float *)slurp(f, R * 12);
    uint8_t *aa = (uint8_t *)slurp(f, R), *cbv = (uint8_t *)slurp(f, R), *chain = (uint8_t *)slurp(f, R);
    uint64_t *serial = (uint64_t *)slurp(f, R * 8);
    uint64_t qs_id = 0, qlen = 0;
    if (fread(&qs_id, 8, 1, f) != 1 || fread(&qlen, 8, 1, f) != 1) return 2;
    char *qraw = (char *)slurp(f, qlen);
    char *qstr = (char *)calloc(qlen + 1, 1);
    memcpy(qstr, qraw, qlen);
    free(qraw);
    fclose(f);

    /* the database: a store of CompactStructures; the index from it (K1 + K2), attached with its lookup */
    fdh_store *store = fdh_store_new();
    fdh_compact *query_struct = NULL;
    for (uint64_t s = 0; s < S; s++) {
        const uint64_t b = ro[s], n = ro[s + 1] - ro[s];
        fdh_compact *c = fdh_compact_from_soa((int64_t)n, nx + 3 * b, ca + 3 * b, cb + 3 * b, cbv + b, aa + b, chain + b,
                                              serial + b, NULL);
        char name[32];
        snprintf(name, sizeof(name), "s%llu", (unsigned long long)s);
        if (s + 1 < S) fdh_store_add(store, c, name); /* the last structure of the file is the query structure */
        if (s == qs_id) query_struct = c;
        else fdh_compact_free(c);
    }
    fd_hash_params hp = {0, 0, 20.0f, FD_HASH_DEFAULT, 0, {0}};
    fd_struct_batch batch;
    if (fdh_store_batch(store, &batch) != FD_OK) return 4;
    fd_index_buffers ib;
    if (fd_build_index(ctx, &batch, &hp, 0, 0, 1ull << 32, &ib) != FD_OK) {
        fprintf(stderr, "fd_build_index: %s\n", fd_last_error(ctx));
        return 4;
    }
    printf("index %llu hashes %llu posting bytes\n", (unsigned long long)ib.count, (unsigned long long)ib.value_bytes);
    fdh_index *ix = fdh_index_from_buffers(&ib, store, &hp);
    fd_free_index_buffers(&ib);
    if (!ix || fdh_index_attach(ctx, ix) != FD_OK) return 5;
    if (fd_store_attach(ctx, &batch) != FD_OK) return 5;
    uint64_t table_bytes = 0;
    if (fd_store_build_pair_table(ctx, &hp, 0, &table_bytes) != FD_OK) return 5;
    printf("structures %llu pair table %llu bytes\n", (unsigned long long)fd_index_num_structs(ctx),
           (unsigned long long)table_bytes);

    /* one query through the host interface: make_query_map -> count_query -> verification -> rows */
    const float dthr[1] = {0.5f}, athr[1] = {5.0f};
    fdh_query_params qp = {hp, dthr, 1, athr, 1, 0};
    fdh_queries *qs = fdh_queries_new(&qp);
    if (fdh_queries_add(qs, query_struct, qstr) < 0 || fdh_queries_finalize(qs, ctx) != FD_OK) {
        fprintf(stderr, "query: %s\n", fdh_last_error());
        return 6;
    }
    fdh_search_params sp;
    memset(&sp, 0, sizeof(sp));
    sp.prefilter.sampling_ratio = -1.0f;
    sp.prefilter.sampling_count = -1;
    sp.prefilter.freq_filter = -1.0f;
    sp.prefilter.length_penalty = 0.5f;
    sp.prefilter.num_res_cutoff = 50000;
    sp.prefilter.top_n = UINT64_MAX;
    sp.ca_dist_cutoff = 1.0f;
    fdh_results *res = fdh_search(ctx, qs, &sp, store);
    if (!res) {
        fprintf(stderr, "fdh_search: %s\n", fdh_last_error());
        return 7;
    }
    const uint64_t *so = fdh_results_struct_offsets(res), *mo = fdh_results_match_offsets(res);
    const fdh_struct_row *sr = fdh_results_struct_rows(res);
    const fdh_match_row *mr = fdh_results_match_rows(res);
    const uint64_t *order = fdh_results_match_order(res);
    for (uint64_t k = so[0]; k < so[1]; k++)
        printf("S %u %u %u %u %.4f\n", sr[k].nid, sr[k].total_match_count, sr[k].node_count, sr[k].edge_count, sr[k].idf);
    for (uint64_t k = mo[0]; k < mo[1]; k++) {
        const fdh_match_row *m = &mr[order[k]];
        printf("M %u %u %.4f %.4f\n", m->nid, m->node_count, m->idf, m->rmsd);
    }
    printf("launches %llu\n", (unsigned long long)fd_kernel_launches(ctx));
    fdh_results_free(res);
    fdh_queries_free(qs);
    fdh_index_free(ix);
    fdh_compact_free(query_struct);
    fdh_store_free(store);
    fd_destroy(ctx);
    free(ro); free(nx); free(ca); free(cb); free(aa); free(cbv); free(chain); free(serial); free(qstr);
    return 0;
}
