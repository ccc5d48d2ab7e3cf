"""Python mirror of the reference's host interface for the hot path, over include/folddisco_b200_host.h.

Names follow the reference (src/lib.rs prelude): read_structure_from_path, parse_query_string, Folddisco
(index builder), load_folddisco_index, query (make_query_map + count_query + retrieval_wrapper for a batch).
All compute runs in the CUDA kernels behind the C ABI.
"""
import ctypes as C
import os

import numpy as np

from . import capi
from .capi import FdError, HashParams, PrefilterParams, VP, _StructBatch, _IndexBuffers, _ptr

UINT64_MAX = capi.UINT64_MAX


class _QueryParams(C.Structure):
    _fields_ = [("hash", HashParams), ("dist_thr", VP), ("n_dist_thr", C.c_int), ("angle_thr", VP),
                ("n_angle_thr", C.c_int), ("serial_query", C.c_int)]


class SearchParams(C.Structure):
    """Reference CLI defaults (src/cli/main.rs:49-97)."""
    _fields_ = [("prefilter", PrefilterParams), ("ca_dist_cutoff", C.c_float), ("skip_match", C.c_int),
                ("max_matching_node_count", C.c_uint64), ("max_matching_node_ratio", C.c_float),
                ("rmsd_cutoff", C.c_float), ("connected_node_count", C.c_uint64), ("connected_node_ratio", C.c_float),
                ("skip_ca_match", C.c_int), ("host_threads", C.c_int), ("verify_mode", C.c_int),
                ("want_metrics", C.c_int), ("partial_fit", C.c_int)]

    def __init__(self, top_n=UINT64_MAX, ca_dist_cutoff=1.0, skip_match=False, host_threads=0, verify_mode=0,
                 want_metrics=False, partial_fit=False, **prefilter):
        super().__init__(PrefilterParams(top_n=top_n, **prefilter), ca_dist_cutoff, int(skip_match), 0, 0.0, 0.0, 0,
                         0.0, 0, host_threads, verify_mode, int(want_metrics), int(partial_fit))


STRUCT_ROW = np.dtype([("nid", np.uint32), ("total_match_count", np.uint32), ("node_count", np.uint32),
                       ("edge_count", np.uint32), ("idf", np.float32), ("max_matching_node_count", np.uint32),
                       ("min_rmsd_with_max_match", np.float32), ("_pad", np.uint32), ("match_begin", np.uint64),
                       ("match_end", np.uint64)])
MATCH_ROW = np.dtype([("nid", np.uint32), ("node_count", np.uint32), ("idf", np.float32), ("rmsd", np.float32),
                      ("U", np.float32, 9), ("t", np.float32, 3), ("res_begin", np.uint64)])
RES_MATCH = np.dtype([("some", np.uint8), ("chain", np.uint8), ("_pad", np.uint8, 6), ("serial", np.uint64)])

_sigs_done = False


def _lib():
    global _sigs_done
    L = capi.lib()
    if _sigs_done:
        return L

    def sig(name, res, args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args

    PP = C.POINTER
    sig("fdh_last_error", C.c_char_p, [])
    sig("fdh_compact_read_pdb", VP, [C.c_char_p])
    sig("fdh_compact_read_structure", VP, [C.c_char_p])
    sig("fdh_compact_from_atoms", VP, [C.c_int64, VP, VP, VP, VP, VP, VP, VP, VP])
    sig("fdh_compact_from_soa", VP, [C.c_int64, VP, VP, VP, VP, VP, VP, VP, VP])
    sig("fdh_compact_nres", C.c_int64, [VP])
    sig("fdh_compact_num_residues_raw", C.c_int64, [VP])
    sig("fdh_compact_first_chain", C.c_int, [VP])
    sig("fdh_compact_avg_plddt", C.c_float, [VP])
    sig("fdh_compact_get", None, [VP] + [VP] * 8)
    sig("fdh_compact_free", None, [VP])
    sig("fdh_parse_path_by_id_type", C.c_int64, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64])
    sig("fdh_fcz_db_open", VP, [C.c_char_p])
    sig("fdh_fcz_db_close", None, [VP])
    sig("fdh_fcz_db_size", C.c_int64, [VP])
    sig("fdh_fcz_db_name", C.c_char_p, [VP, C.c_int64])
    sig("fdh_fcz_db_key", C.c_uint64, [VP, C.c_int64])
    sig("fdh_fcz_db_find", C.c_int64, [VP, C.c_char_p])
    sig("fdh_fcz_db_read", VP, [VP, C.c_int64])
    sig("fdh_compact_from_fcz", VP, [C.c_char_p, C.c_uint64])
    sig("fdh_index_db_key", C.c_uint64, [VP, C.c_uint64])
    sig("fdh_index_set_db_keys", C.c_int, [VP, VP, C.c_uint64])
    sig("fdh_index_foldcomp_db", C.c_char_p, [VP])
    sig("fdh_store_new", VP, [])
    sig("fdh_store_add", C.c_int64, [VP, VP, C.c_char_p])
    sig("fdh_store_add_soa", C.c_int64, [VP, C.c_uint64, VP, VP, VP, VP, VP, C.c_char_p])
    sig("fdh_store_size", C.c_uint64, [VP])
    sig("fdh_store_num_residues", C.c_uint64, [VP])
    sig("fdh_store_get_lookup", None, [VP, VP, VP])
    sig("fdh_store_name", C.c_char_p, [VP, C.c_uint64])
    sig("fdh_store_batch", C.c_int, [VP, PP(_StructBatch)])
    sig("fdh_store_attach", C.c_int, [VP, VP])
    sig("fdh_store_free", None, [VP])
    sig("fdh_store_save", C.c_int, [VP, C.c_char_p])
    sig("fdh_store_load", VP, [C.c_char_p])
    sig("fdh_index_build", VP, [VP, VP, PP(HashParams)])
    sig("fdh_index_from_buffers", VP, [PP(_IndexBuffers), VP, PP(HashParams)])
    sig("fdh_index_save", C.c_int, [VP, VP, C.c_char_p, C.c_uint64, C.c_char_p])
    sig("fdh_index_load", VP, [C.c_char_p])
    sig("fdh_index_get", C.c_int, [VP, PP(_IndexBuffers)])
    sig("fdh_index_num_structs", C.c_uint64, [VP])
    sig("fdh_index_get_lookup", None, [VP, VP, VP])
    sig("fdh_index_name", C.c_char_p, [VP, C.c_uint64])
    sig("fdh_index_get_params", None, [VP, PP(HashParams)])
    sig("fdh_index_attach", C.c_int, [VP, VP])
    sig("fdh_index_free", None, [VP])
    sig("fdh_parse_query_string", C.c_int64, [C.c_char_p, C.c_uint8, VP, VP, VP, VP, VP, C.c_int64, C.c_int64])
    sig("fdh_queries_new", VP, [PP(_QueryParams)])
    sig("fdh_queries_add", C.c_int64, [VP, VP, C.c_char_p])
    sig("fdh_queries_add_many", C.c_int64, [VP, PP(VP), PP(C.c_char_p), C.c_int64, C.c_int])
    sig("fdh_queries_add_many_indexed", C.c_int64, [VP, PP(VP), C.c_int64, PP(C.c_char_p), C.c_int64, VP, VP, C.c_int64, C.c_int])
    sig("fdh_queries_size", C.c_int64, [VP])
    sig("fdh_queries_finalize", C.c_int, [VP, VP])
    sig("fdh_queries_num_hashes", C.c_int64, [VP, C.c_int64])
    sig("fdh_queries_get_map", None, [VP, C.c_int64, VP, VP, VP, VP, VP])
    sig("fdh_queries_num_indices", C.c_int64, [VP, C.c_int64])
    sig("fdh_queries_get_indices", None, [VP, C.c_int64, VP])
    sig("fdh_queries_free", None, [VP])
    sig("fdh_search", VP, [VP, VP, PP(SearchParams), VP])
    sig("fdh_queries_finalize_sharded", C.c_int, [VP, VP, C.c_uint64, C.c_uint64])
    sig("fdh_search_sharded", VP, [VP, VP, PP(SearchParams), VP])
    sig("fdh_queries_set_shards", C.c_int, [VP, VP, C.c_int])
    sig("fdh_queries_num_vote_bits", C.c_int64, [VP, C.c_int64])
    sig("fdh_queries_get_vote_bits", None, [VP, C.c_int64, VP, VP, VP, VP])
    sig("fdh_queries_scan_sizes", None, [VP, PP(C.c_uint64), PP(C.c_uint64)])
    sig("fdh_queries_scan_arrays", None, [VP, VP, VP, VP, VP])
    sig("fdh_votes_scan_sparse_flat", C.c_int, [VP, C.c_uint32, VP, VP, VP, PP(PrefilterParams), VP, C.c_uint32,
                                                PP(capi.VotesLayout), PP(VP), VP, VP])
    sig("fdh_queries_num_pairs", C.c_int64, [VP])
    sig("fdh_queries_pair_counts", C.c_int, [VP, VP, VP])
    sig("fdh_queries_finalize_with_counts", C.c_int, [VP, VP, C.c_uint64])
    sig("fdh_votes_scan", C.c_int, [VP, VP, PP(PrefilterParams), PP(capi.VotesLayout), PP(VP)])
    sig("fdh_search_from_votes", VP, [VP, VP, PP(SearchParams), VP, PP(capi.VotesLayout), VP, C.c_uint32, C.c_uint32])
    sig("fdh_results_num_queries", C.c_uint64, [VP])
    for n in ("struct_offsets", "struct_rows", "match_offsets", "match_rows", "match_order", "residues"):
        sig("fdh_results_" + n, VP, [VP])
    sig("fdh_results_num_residues", C.c_uint64, [VP])
    sig("fdh_results_metrics", VP, [VP])
    sig("fdh_results_residue_index", VP, [VP])
    sig("fdh_store_get_ca", C.c_int, [VP, C.c_uint64, C.c_uint64, VP])
    sig("fdh_results_host_ms", C.c_double, [VP])
    sig("fdh_results_h2d_bytes", C.c_uint64, [VP])
    sig("fdh_results_wall_ms", C.c_double, [VP, C.c_int])
    sig("fdh_results_d2h_bytes", C.c_uint64, [VP])
    sig("fdh_results_free", None, [VP])
    _sigs_done = True
    return L


def _err():
    return _lib().fdh_last_error().decode()


class CompactStructure:
    """src/structure/core.rs:55-67"""

    def __init__(self, handle):
        if not handle:
            raise FdError(_err())
        self.h = handle

    @classmethod
    def from_atoms(cls, a):
        arrs = [np.ascontiguousarray(a["x"], np.float32), np.ascontiguousarray(a["y"], np.float32),
                np.ascontiguousarray(a["z"], np.float32), np.ascontiguousarray(a["atom_name"], np.uint8).reshape(-1),
                np.ascontiguousarray(a["chain"], np.uint8), np.ascontiguousarray(a["res_name"], np.uint8).reshape(-1),
                np.ascontiguousarray(a["res_serial"], np.uint64), np.ascontiguousarray(a["b_factor"], np.float32)]
        return cls(_lib().fdh_compact_from_atoms(len(arrs[0]), *[_ptr(x) for x in arrs]))

    @classmethod
    def from_soa(cls, n_xyz, ca_xyz, cb_xyz, aa, cb_valid=None, chain=None, serial=None, b_factor=None):
        req = [np.ascontiguousarray(n_xyz, np.float32), np.ascontiguousarray(ca_xyz, np.float32),
               np.ascontiguousarray(cb_xyz, np.float32)]
        aa = np.ascontiguousarray(aa, np.uint8)
        opt = [None if v is None else np.ascontiguousarray(v, dt) for v, dt in
               ((cb_valid, np.uint8), (chain, np.uint8), (serial, np.uint64), (b_factor, np.float32))]
        return cls(_lib().fdh_compact_from_soa(len(aa), _ptr(req[0]), _ptr(req[1]), _ptr(req[2]), _ptr(opt[0]),
                                               _ptr(aa), _ptr(opt[1]), _ptr(opt[2]), _ptr(opt[3])))

    @property
    def num_residues(self):
        return _lib().fdh_compact_nres(self.h)

    @property
    def first_chain(self):
        return _lib().fdh_compact_first_chain(self.h)

    @property
    def avg_plddt(self):
        return _lib().fdh_compact_avg_plddt(self.h)

    def soa(self):
        n = self.num_residues
        d = dict(n_xyz=np.zeros((n, 3), np.float32), ca_xyz=np.zeros((n, 3), np.float32),
                 cb_xyz=np.zeros((n, 3), np.float32), cb_valid=np.zeros(n, np.uint8), aa=np.zeros(n, np.uint8),
                 chain=np.zeros(n, np.uint8), serial=np.zeros(n, np.uint64), b_factor=np.zeros(n, np.float32))
        _lib().fdh_compact_get(self.h, *[_ptr(d[k]) for k in ("n_xyz", "ca_xyz", "cb_xyz", "cb_valid", "aa", "chain",
                                                               "serial", "b_factor")])
        return d

    def __del__(self):
        if getattr(self, "h", None):
            _lib().fdh_compact_free(self.h)
            self.h = None


def read_structure_from_path(path):
    """read_structure_from_path(path).to_compact()  (src/controller/io.rs:337-379): .pdb / .ent / .cif, optionally .gz"""
    return CompactStructure(_lib().fdh_compact_read_structure(os.fsencode(path)))


def parse_path_by_id_type(path, id_type):
    """src/controller/mode.rs:19-31, 70-125: the lookup id of an input path under `--id TYPE`"""
    n = _lib().fdh_parse_path_by_id_type(os.fsencode(path), id_type.encode(), None, 0)
    buf = C.create_string_buffer(n + 1)
    _lib().fdh_parse_path_by_id_type(os.fsencode(path), id_type.encode(), buf, n + 1)
    return buf.value.decode()


class FoldcompDb:
    """FoldcompDbReader (src/structure/io/fcz.rs:21-136): PATH + PATH.index + PATH.lookup.  Entries are decoded by the
    Foldcomp codec library, bound at run time through the C ABI the reference binds ($FD_FOLDCOMP_LIB or
    libfoldcomp_ffi.so on the loader path); the named entries are listed in ascending key order (get_paths)."""

    def __init__(self, path):
        self.h = _lib().fdh_fcz_db_open(os.fsencode(path))
        if not self.h:
            raise FdError(_err())

    def __len__(self):
        return _lib().fdh_fcz_db_size(self.h)

    def names(self):
        return [os.fsdecode(_lib().fdh_fcz_db_name(self.h, k)) for k in range(len(self))]  # names are bytes on disk

    def keys(self):
        return [int(_lib().fdh_fcz_db_key(self.h, k)) for k in range(len(self))]

    def find(self, name):
        return _lib().fdh_fcz_db_find(self.h, os.fsencode(name))

    def read(self, k):
        """read_single_structure_by_id(...).to_compact() of the k-th named entry"""
        return CompactStructure(_lib().fdh_fcz_db_read(self.h, k))

    def read_by_name(self, name):
        k = self.find(name)
        if k < 0:
            raise FdError("Entry with name %s not found." % name)
        return self.read(k)

    def __del__(self):
        if getattr(self, "h", None):
            _lib().fdh_fcz_db_close(self.h)
            self.h = None


def compact_from_fcz(data):
    """one Foldcomp-compressed entry (bytes) -> CompactStructure"""
    return CompactStructure(_lib().fdh_compact_from_fcz(data, len(data)))


def parse_query_string(q, default_chain=ord("A")):
    """src/controller/query.rs:331-384 -> ([(chain, residue)], [None | [aa,...]])"""
    cap = 1 << 16
    chains, serials = np.zeros(cap, np.uint8), np.zeros(cap, np.uint64)
    off, end, subs = np.zeros(cap, np.int64), np.zeros(cap, np.int64), np.zeros(cap * 20, np.uint8)
    n = _lib().fdh_parse_query_string(q.encode(), default_chain, _ptr(chains), _ptr(serials), _ptr(off), _ptr(end),
                                      _ptr(subs), cap, cap * 20)
    if n < 0:
        raise ValueError("Invalid residue in query string %r" % q)
    res = [(int(chains[i]), int(serials[i])) for i in range(n)]
    sub = [None if off[i] < 0 else subs[off[i]:end[i]].tolist() for i in range(n)]
    return res, sub


class Store:
    """The database: an ordered set of CompactStructures (ids = positions = `.lookup` ids)."""

    def __init__(self, handle=None):
        self.h = handle if handle is not None else _lib().fdh_store_new()

    def save(self, path):
        """PREFIX.store: the on-disk companion of the index (fdh_store_save)"""
        if _lib().fdh_store_save(self.h, path.encode()) != 0:
            raise FdError(_err())

    @classmethod
    def load(cls, path):
        h = _lib().fdh_store_load(path.encode())
        if not h:
            raise FdError(_err())
        return cls(h)

    def add(self, compact, name):
        return _lib().fdh_store_add(self.h, compact.h, name.encode())

    def add_soa(self, batch, prefix="synth_"):
        """batch: dict from folddisco_b200.synth.generate"""
        ro = np.ascontiguousarray(batch["row_offsets"], np.uint64)
        return _lib().fdh_store_add_soa(self.h, len(ro) - 1, _ptr(ro), _ptr(np.ascontiguousarray(batch["n_xyz"], np.float32)),
                                        _ptr(np.ascontiguousarray(batch["ca_xyz"], np.float32)),
                                        _ptr(np.ascontiguousarray(batch["cb_xyz"], np.float32)),
                                        _ptr(np.ascontiguousarray(batch["aa"], np.uint8)), prefix.encode())

    def __len__(self):
        return _lib().fdh_store_size(self.h)

    @property
    def num_residues(self):
        return _lib().fdh_store_num_residues(self.h)

    def lookup(self):
        n = len(self)
        nres, plddt = np.zeros(n, np.uint32), np.zeros(n, np.float32)
        _lib().fdh_store_get_lookup(self.h, _ptr(nres), _ptr(plddt))
        return nres, plddt

    def name(self, i):
        return _lib().fdh_store_name(self.h, i).decode()

    def batch_view(self):
        b = _StructBatch()
        _lib().fdh_store_batch(self.h, C.byref(b))
        return b

    def attach(self, ctx, pair_table=False, hash_params=None, max_table_bytes=0):
        """fd_store_attach: copy the compact structures to HBM for candidate verification.  pair_table=True also
        builds the store's pair table (fd_store_build_pair_table: 8 B per hashed residue pair) so that verification
        looks query hashes up instead of re-hashing candidates; returns its size in bytes (None: over the budget)"""
        if _lib().fdh_store_attach(self.h, ctx.h) != 0:  # structures + (chain, residue number) labels
            raise FdError(_err())
        if pair_table:
            return ctx.store_build_pair_table(hash_params, max_table_bytes)
        return 0

    def __del__(self):
        if getattr(self, "h", None):
            _lib().fdh_store_free(self.h)
            self.h = None


class FolddiscoIndex:
    """src/index/indextable.rs FolddiscoIndex + lookup + config."""

    def __init__(self, handle):
        if not handle:
            raise FdError(_err())
        self.h = handle

    @classmethod
    def build(cls, ctx, store, params=None):
        """Folddisco::collect_and_count / allocate_entries / add_entries on the GPU"""
        params = params or HashParams()
        return cls(_lib().fdh_index_build(ctx.h, store.h, C.byref(params)))

    def save(self, store, prefix, max_residue=50000, foldcomp_db=None):
        rc = _lib().fdh_index_save(self.h, store.h if store is not None else None, os.fsencode(prefix), max_residue,
                                   None if foldcomp_db is None else foldcomp_db.encode())
        if rc != 0:
            raise FdError(_err())

    def set_db_keys(self, keys):
        """the database keys of an index built from a Foldcomp database (Folddisco::numeric_db_key_vec): 5th lookup
        column; save() then records input_format FCZDB"""
        k = np.ascontiguousarray(keys, np.uint64)
        if _lib().fdh_index_set_db_keys(self.h, _ptr(k), len(k)) != 0:
            raise FdError(_err())

    @property
    def foldcomp_db(self):
        return _lib().fdh_index_foldcomp_db(self.h).decode()

    def buffers(self):
        v = _IndexBuffers()
        _lib().fdh_index_get(self.h, C.byref(v))
        return capi.IndexBuffers(
            np.ctypeslib.as_array(v.hashes, (v.count,)).copy() if v.count else np.zeros(0, np.uint32),
            np.ctypeslib.as_array(v.offsets, (v.count + 1,)).copy(),
            np.ctypeslib.as_array(v.values, (v.value_bytes,)).copy() if v.value_bytes else np.zeros(0, np.uint8))

    @property
    def num_structs(self):
        return _lib().fdh_index_num_structs(self.h)

    def lookup(self):
        n = self.num_structs
        nres, plddt = np.zeros(n, np.uint32), np.zeros(n, np.float32)
        _lib().fdh_index_get_lookup(self.h, _ptr(nres), _ptr(plddt))
        return nres, plddt

    def name(self, i):
        return _lib().fdh_index_name(self.h, i).decode()

    @property
    def params(self):
        p = HashParams()
        _lib().fdh_index_get_params(self.h, C.byref(p))
        return p

    def attach(self, ctx):
        ctx._check(_lib().fdh_index_attach(ctx.h, self.h), "fd_index_attach")
        ctx.n_structs = self.num_structs

    def __del__(self):
        if getattr(self, "h", None):
            _lib().fdh_index_free(self.h)
            self.h = None


def load_folddisco_index(prefix):
    """load_folddisco_index + load_lookup_from_file + read_index_config_from_file"""
    return FolddiscoIndex(_lib().fdh_index_load(os.fsencode(prefix)))


class QueryInputs:
    """The host-side inputs of a batch -- CompactStructure objects and query strings, query k = (compacts[which_compact[k]],
    query_strings[which_string[k]]) -- held as the C arrays fdh_queries_add_many_indexed takes.  Keeps the Python
    objects alive.  Built once per set of inputs; QueryBatch.add_prepared uses it any number of times."""

    def __init__(self, compacts, query_strings, which_compact=None, which_string=None):
        self.compacts = list(compacts)
        self.query_strings = list(query_strings)
        n = len(self.compacts)
        self.which_compact = np.ascontiguousarray(np.arange(n) if which_compact is None else which_compact, np.uint32)
        self.which_string = np.ascontiguousarray(np.arange(len(self.query_strings)) if which_string is None else which_string,
                                                 np.uint32)
        assert len(self.which_compact) == len(self.which_string)
        self.n_compacts, self.n_strings = n, len(self.query_strings)
        self.handles = (VP * n)(*[c.h for c in self.compacts])
        self._encoded = [q.encode() for q in self.query_strings]
        self.strings = (C.c_char_p * len(self._encoded))(*self._encoded)
        self.per_query = [self.query_strings[int(k)] for k in self.which_string]


class QueryBatch:
    """A batch of (structure, query string) with their query maps (make_query_map, query.rs:208-329)."""

    def __init__(self, hash_params=None, dist_thr=(0.5,), angle_thr=(5.0,), serial_query=False):
        self._dt = np.asarray(dist_thr, np.float32)
        self._at = np.asarray(angle_thr, np.float32)
        p = _QueryParams(hash_params or HashParams(), _ptr(self._dt), len(self._dt), _ptr(self._at), len(self._at),
                         int(serial_query))
        self.h = _lib().fdh_queries_new(C.byref(p))
        if not self.h:
            raise FdError(_err())
        self.query_strings = []

    def add(self, compact, query_string):
        q = _lib().fdh_queries_add(self.h, compact.h, query_string.encode())
        if q < 0:
            raise FdError(_err())
        self.query_strings.append(query_string)
        return q

    def add_many(self, compacts, query_strings, threads=0):
        """make_query_map for many (structure, query string) pairs, query-parallel on the host"""
        n = len(compacts)
        hs = (VP * n)(*[c.h for c in compacts])
        qs = (C.c_char_p * n)(*[q.encode() for q in query_strings])
        first = _lib().fdh_queries_add_many(self.h, hs, qs, n, threads)
        if first < 0:
            raise FdError(_err())
        self.query_strings.extend(query_strings)
        return first

    def add_many_indexed(self, compacts, query_strings, which_compact, which_string, threads=0):
        """add_many for batches that reuse a few structures / query strings: query k is
        (compacts[which_compact[k]], query_strings[which_string[k]]).  Every query map is still built on its own; only
        the marshalling of 2 x n Python objects per call goes away (~1 ms per 1 024 queries)."""
        wc = np.ascontiguousarray(which_compact, np.uint32)
        ws = np.ascontiguousarray(which_string, np.uint32)
        assert len(wc) == len(ws)
        hs = (VP * len(compacts))(*[c.h for c in compacts])
        qs = (C.c_char_p * len(query_strings))(*[q.encode() for q in query_strings])
        first = _lib().fdh_queries_add_many_indexed(self.h, hs, len(compacts), qs, len(query_strings), _ptr(wc), _ptr(ws),
                                                    len(wc), threads)
        if first < 0:
            raise FdError(_err())
        self.query_strings.extend(query_strings[int(k)] for k in ws)
        return first

    def add_prepared(self, inputs, threads=0):
        """add_many_indexed for inputs that were marshalled once (QueryInputs): the call hands the library the same C
        arrays every time, so a serving loop that answers batch after batch from the same host structures pays no
        Python list -> C array conversion per batch (1.2 ms per 1 024 queries)"""
        first = _lib().fdh_queries_add_many_indexed(self.h, inputs.handles, inputs.n_compacts, inputs.strings,
                                                    inputs.n_strings, _ptr(inputs.which_compact), _ptr(inputs.which_string),
                                                    len(inputs.which_compact), threads)
        if first < 0:
            raise FdError(_err())
        if self.query_strings:
            self.query_strings.extend(inputs.per_query)
        else:
            self.query_strings = inputs.per_query  # shared, read-only
        return first

    def __len__(self):
        return _lib().fdh_queries_size(self.h)

    def finalize(self, ctx):
        rc = _lib().fdh_queries_finalize(self.h, ctx.h)
        if rc != 0:
            raise FdError(_err())

    # ---- id-range shards over the library's own NCCL communicator (ctx.comm_init) ----
    def finalize_sharded(self, ctx, first_id, total_structs):
        """collective: all-gather of the ranks' query descriptors + all-reduce of the posting counts"""
        if _lib().fdh_queries_finalize_sharded(self.h, ctx.h, int(first_id), int(total_structs)) != 0:
            raise FdError(_err())

    # ---- hash-range shards (see folddisco_b200/sharded.py) ----
    def set_shards(self, bounds):
        b = np.ascontiguousarray(bounds, np.uint64)
        if _lib().fdh_queries_set_shards(self.h, _ptr(b), len(b) - 1) != 0:
            raise FdError(_err())

    def vote_bits(self, q):
        """the fd_query arrays of query q: hashes, bit_of_hash (per hash), bit_node, bit_group (per vote bit)"""
        nh, nb = _lib().fdh_queries_num_hashes(self.h, q), _lib().fdh_queries_num_vote_bits(self.h, q)
        d = dict(hashes=np.zeros(nh, np.uint32), bit_of_hash=np.zeros(nh, np.uint16), bit_node=np.zeros(nb, np.uint16),
                 bit_group=np.zeros(nb, np.uint16))
        _lib().fdh_queries_get_vote_bits(self.h, q, *[_ptr(d[k]) for k in ("hashes", "bit_of_hash", "bit_node", "bit_group")])
        return d

    def scan_arrays(self):
        """flat scan inputs of this batch: per_query u32[nq, 3] = (hashes, vote bits, pairs), hashes, bit_of_hash,
        pair_hashes (all uint32)"""
        nh, npair = C.c_uint64(), C.c_uint64()
        _lib().fdh_queries_scan_sizes(self.h, C.byref(nh), C.byref(npair))
        nq = len(self)
        per_query = np.zeros((nq, 3), np.uint32)
        hashes, bits = np.zeros(nh.value, np.uint32), np.zeros(nh.value, np.uint32)
        pairs = np.zeros(npair.value, np.uint32)
        _lib().fdh_queries_scan_arrays(self.h, _ptr(per_query), _ptr(hashes), _ptr(bits), _ptr(pairs))
        return per_query, hashes, bits, pairs

    def pair_counts(self, ctx):
        """posting counts of every query pair's observed hash in the index attached to ctx (0 if absent)"""
        out = np.zeros(_lib().fdh_queries_num_pairs(self.h), np.uint32)
        if _lib().fdh_queries_pair_counts(self.h, ctx.h, _ptr(out)) != 0:
            raise FdError(_err())
        return out

    def finalize_with_counts(self, counts, total_structures):
        c = np.ascontiguousarray(counts, np.uint32)
        assert len(c) == _lib().fdh_queries_num_pairs(self.h)
        if _lib().fdh_queries_finalize_with_counts(self.h, _ptr(c), int(total_structures)) != 0:
            raise FdError(_err())

    def query_map(self, q):
        n = _lib().fdh_queries_num_hashes(self.h, q)
        d = dict(hash=np.zeros(n, np.uint32), qi=np.zeros(n, np.int64), qj=np.zeros(n, np.int64),
                 primary=np.zeros(n, np.uint8), idf=np.zeros(n, np.float32))
        _lib().fdh_queries_get_map(self.h, q, _ptr(d["hash"]), _ptr(d["qi"]), _ptr(d["qj"]), _ptr(d["primary"]),
                                   _ptr(d["idf"]))
        return d

    def indices(self, q):
        n = _lib().fdh_queries_num_indices(self.h, q)
        out = np.zeros(n, np.int64)
        _lib().fdh_queries_get_indices(self.h, q, _ptr(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            _lib().fdh_queries_free(self.h)
            self.h = None


class Results:
    """Rows of a batch search.  The arrays are zero-copy views of the library-owned result object."""

    def __init__(self, handle):
        if not handle:
            raise FdError(_err())
        L = _lib()
        self.h = handle
        nq = L.fdh_results_num_queries(handle)

        def arr(ptr, n, dt):
            if n == 0 or not ptr:
                return np.zeros(0, dt)
            return np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr), dtype=dt, count=n)

        self.struct_offsets = arr(L.fdh_results_struct_offsets(handle), nq + 1, np.uint64)
        self.match_offsets = arr(L.fdh_results_match_offsets(handle), nq + 1, np.uint64)
        ns, nm = int(self.struct_offsets[-1]), int(self.match_offsets[-1])
        self.structs = arr(L.fdh_results_struct_rows(handle), ns, STRUCT_ROW)
        self.matches = arr(L.fdh_results_match_rows(handle), nm, MATCH_ROW)
        self.match_order = arr(L.fdh_results_match_order(handle), nm, np.uint64)
        self.residues = arr(L.fdh_results_residues(handle), L.fdh_results_num_residues(handle), RES_MATCH)
        # SearchParams(want_metrics=True): [nm, 5] = tm_score, gdt_ts, gdt_ha, chamfer_distance, hausdorff_distance per
        # match row, and the target residue index + 1 behind every residue entry (copies; None otherwise)
        mp, rp = L.fdh_results_metrics(handle), L.fdh_results_residue_index(handle)
        self.metrics = arr(mp, 5 * nm, np.float32).reshape(-1, 5).copy() if mp else None
        self.residue_index = arr(rp, L.fdh_results_num_residues(handle), np.uint32).copy() if rp else None
        self.host_ms = L.fdh_results_host_ms(handle)
        self.h2d_bytes = L.fdh_results_h2d_bytes(handle)
        self.d2h_bytes = L.fdh_results_d2h_bytes(handle)
        self.wall_ms = {k: L.fdh_results_wall_ms(handle, i) for i, k in enumerate(("count_query", "verify", "assemble", "total"))}

    def __del__(self):
        if getattr(self, "h", None):
            for k in ("struct_offsets", "match_offsets", "structs", "matches", "match_order", "residues"):
                setattr(self, k, None)
            _lib().fdh_results_free(self.h)
            self.h = None

    def structures(self, q):
        """per-structure rows of query q (a copy: the library-owned block is recycled when this object is freed)"""
        return self.structs[int(self.struct_offsets[q]):int(self.struct_offsets[q + 1])].copy()

    def sorted_matches(self, q):
        """per-match rows of query q in the default output order (idf desc, rmsd asc)"""
        o = self.match_order[int(self.match_offsets[q]):int(self.match_offsets[q + 1])]
        return self.matches[o.astype(np.int64)]

    def residue_string(self, match_row, n_query_residues):
        r = self.residues[int(match_row["res_begin"]):int(match_row["res_begin"]) + n_query_residues]
        return ",".join("%s%d" % (chr(x["chain"]), x["serial"]) if x["some"] else "_" for x in r)


def votes_scan(ctx, queries, prefilter):
    """partial votes of ctx's index shard for the whole batch -> (VotesLayout, device pointer owned by ctx)"""
    lay, ptr = capi.VotesLayout(), VP()
    if _lib().fdh_votes_scan(ctx.h, queries.h, C.byref(prefilter), C.byref(lay), C.byref(ptr)) != 0:
        raise FdError(_err())
    return lay, ptr.value


def votes_scan_sparse(ctx, per_query, hashes, bit_of_hash, prefilter, slice_begin):
    """sparse partial votes of ctx's shard for the whole (gathered) batch -> (layout, device pointer of the record
    pool, region_offset u64[world + 1], region_count u64[world]); records have 1 + layout.planes words"""
    per_query = np.ascontiguousarray(per_query, np.uint32)
    hashes = np.ascontiguousarray(hashes, np.uint32)
    bits = np.ascontiguousarray(bit_of_hash, np.uint32)
    sb = np.ascontiguousarray(slice_begin, np.uint32)
    world = len(sb) - 1
    lay, ptr = capi.VotesLayout(), VP()
    off, cnt = np.zeros(world + 1, np.uint64), np.zeros(world, np.uint64)
    if _lib().fdh_votes_scan_sparse_flat(ctx.h, len(per_query), _ptr(per_query), _ptr(hashes), _ptr(bits),
                                         C.byref(prefilter), _ptr(sb), world, C.byref(lay), C.byref(ptr), _ptr(off),
                                         _ptr(cnt)) != 0:
        raise FdError(_err())
    return lay, ptr.value, off, cnt


def votes_merge_begin(ctx, layout, n_queries):
    """cleared dense vote planes for this rank's n_queries queries -> (slice layout, device pointer)"""
    sl, ptr = capi.VotesLayout(), VP()
    ctx._check(capi.lib().fd_votes_merge_begin(ctx.h, C.byref(layout), 0, n_queries, C.byref(sl), C.byref(ptr)),
               "fd_votes_merge_begin")
    return sl, ptr.value


def votes_apply(ctx, slice_layout, d_dense, d_records, n_records):
    ctx._check(capi.lib().fd_votes_apply(ctx.h, C.byref(slice_layout), VP(d_dense), VP(d_records), int(n_records)),
               "fd_votes_apply")


def search_from_votes(ctx, queries, params, layout, d_votes, q_begin, q_end, labels=None):
    """the rest of the search for queries [q_begin, q_end) from merged votes"""
    return Results(_lib().fdh_search_from_votes(ctx.h, queries.h, C.byref(params), labels.h if labels is not None else None,
                                                C.byref(layout), VP(d_votes), q_begin, q_end))


def search(ctx, queries, params=None, labels=None):
    """query_pdb.rs:348-452 for the batch: count_query -> filter/sort/top -> retrieval -> Kabsch -> filters -> sort"""
    params = params or SearchParams()
    return Results(_lib().fdh_search(ctx.h, queries.h, C.byref(params), labels.h if labels is not None else None))


def search_sharded(ctx, queries, params=None, labels=None):
    """collective: this rank's own queries against the id-range sharded database (QueryBatch.finalize_sharded first)"""
    params = params or SearchParams()
    return Results(_lib().fdh_search_sharded(ctx.h, queries.h, C.byref(params), labels.h if labels is not None else None))


class QueryMapWorker:
    """make_query_map of a whole batch (QueryBatch.add_prepared) on a second host thread, for a serving loop that
    answers batch after batch: the query maps of batch k+1 are built while batch k is finalized and searched.  A batch
    that has not been finalized touches no context and no device, so the thread does host work only (on the library's
    worker pool, which concurrent regions share); finalize and search stay on the caller's thread -- with id-range
    shards they are collectives and must be issued in the same order on every rank.  The reference's loop is
    query-parallel (query_pdb.rs:348 `into_par_iter`); here the parallelism is between the host half of one batch
    and the device half of the previous one.

        w = QueryMapWorker(index.params); w.start(inputs[0])
        for k in range(n): qb = w.take(); w.start(inputs[k + 1]); qb.finalize(ctx); rows = search(ctx, qb, ...)"""

    def __init__(self, hash_params=None, dist_thr=(0.5,), angle_thr=(5.0,), serial_query=False, threads=0):
        from concurrent.futures import ThreadPoolExecutor
        self._args = (hash_params, tuple(dist_thr), tuple(angle_thr), serial_query)
        if threads <= 0:  # leave a quarter of the host threads to the caller's finalize / search, which run meanwhile
            threads = max(1, capi.lib().fd_default_host_threads() * 3 // 4)
        self._threads = threads
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="fd-query-maps")
        self._fut = None

    def _build(self, inputs):
        hp, dt, at, serial = self._args
        qb = QueryBatch(hp, dist_thr=dt, angle_thr=at, serial_query=serial)
        qb.add_prepared(inputs, self._threads)  # ctypes releases the GIL for the duration of the call
        return qb

    def start(self, inputs):
        """begin the query maps of one batch (a QueryInputs); one batch in flight at a time"""
        if self._fut is not None:
            raise FdError("QueryMapWorker.start: the previous batch has not been taken")
        self._fut = self._pool.submit(self._build, inputs)

    def wait(self):
        """block until the batch in flight is built (it stays with the worker until take)"""
        if self._fut is not None:
            self._fut.exception()

    def take(self):
        """the finished QueryBatch (not finalized); raises what add_prepared raised"""
        if self._fut is None:
            raise FdError("QueryMapWorker.take: nothing was started")
        fut, self._fut = self._fut, None
        return fut.result()

    def close(self):
        self._fut = None
        self._pool.shutdown(wait=True)


def search_batches(ctx, batches, params=None, hash_params=None, labels=None, dist_thr=(0.5,), angle_thr=(5.0,),
                   finalize=None, search_fn=None):
    """Generator: one Results per QueryInputs of `batches`, in order, with the query maps of the next batch built on a
    second host thread (QueryMapWorker) while the current one is finalized and searched.  finalize(qb) / search_fn(qb)
    replace QueryBatch.finalize(ctx) / search(ctx, qb, params, labels) for sharded databases."""
    params = params or SearchParams()
    w = QueryMapWorker(hash_params, dist_thr, angle_thr)
    try:
        it = iter(batches)
        nxt = next(it, None)
        if nxt is not None:
            w.start(nxt)
        while nxt is not None:
            qb = w.take()
            nxt = next(it, None)
            if nxt is not None:
                w.start(nxt)
            if finalize is not None:
                finalize(qb)
            else:
                qb.finalize(ctx)
            yield search_fn(qb) if search_fn is not None else search(ctx, qb, params, labels)
    finally:
        w.close()


class _LaneContext:
    """a fork of a Context that the parent owns (fd_lane): same device, index and store, own stream and staging"""

    def __init__(self, parent, i):
        h = VP()
        parent._check(capi.lib().fd_lane(parent.h, i, C.byref(h)), "fd_lane")
        self.h = h


_stream_pool = None


def search_stream(ctx, structures, query_strings, params=None, hash_params=None, sub_batch=256, labels=None,
                  dist_thr=(0.5,), angle_thr=(5.0,)):
    """Queries given as host structures + query strings -> list of Results, one per sub-batch of `sub_batch` queries, in
    order.  The query side of sub-batch k+1 (make_query_map, the idf lookup, the verification tables) is prepared on a
    second host thread, on a lane of ctx, while sub-batch k is searched: the reference's query loop is query-parallel
    (query_pdb.rs:348 `into_par_iter`), here the parallelism is between the host-heavy and the GPU-heavy half of
    consecutive sub-batches.  Pays off only when a sub-batch is large enough to amortise the per-call fixed costs
    (a few host/device round trips): on the bench workload 1 024 queries in one call (12.3 ms) beat 2 x 512 (13.1 ms)."""
    global _stream_pool
    from concurrent.futures import ThreadPoolExecutor
    params = params or SearchParams()
    n = len(structures)
    if n == 0:
        return []
    if _stream_pool is None:
        _stream_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="fd-prepare")
    lane = _LaneContext(ctx, 7)
    bounds = [(a, min(n, a + sub_batch)) for a in range(0, n, sub_batch)]

    def prepare(k):
        a, b = bounds[k]
        qb = QueryBatch(hash_params, dist_thr=dist_thr, angle_thr=angle_thr)
        qb.add_many(structures[a:b], query_strings[a:b])
        qb.finalize(lane)
        return qb

    out = []
    fut = _stream_pool.submit(prepare, 0)
    for k in range(len(bounds)):
        qb = fut.result()
        if k + 1 < len(bounds):
            fut = _stream_pool.submit(prepare, k + 1)
        out.append(search(ctx, qb, params, labels))
    return out
