"""ctypes binding of include/folddisco_b200.h.  No torch types cross this boundary; no CPU fallback."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfolddisco_b200.so")
COMM_ID_BYTES = 128
UINT64_MAX = (1 << 64) - 1
VP = C.c_void_p


class FdError(RuntimeError):
    pass


def library_path():
    return _SO


_lib = None


class _StructBatch(C.Structure):
    _fields_ = [("n_structs", C.c_uint64), ("row_offsets", VP), ("n_xyz", VP), ("ca_xyz", VP), ("cb_xyz", VP),
                ("aa", VP), ("cb_valid", VP)]


# FD_HASH_* of include/folddisco_b200.h: the reference's HashType index + 1, 0 = default (PDBTrRosetta)
HASH_TYPES = {"default": 0, "PDBMotif": 1, "PDBMotifSinCos": 2, "TrRosetta": 3, "PDBTrRosetta": 4, "PointPairFeature": 5,
              "FolddiscoAngle": 8, "FolddiscoDist": 9}
MAX_MULTIPLE_BINS = 8


class HashParams(C.Structure):
    """fd_hash_params: bins, cutoff, encoding (`--type`) and the `--multiple-bins` list of (dist, angle) pairs"""
    _fields_ = [("nbin_dist", C.c_uint32), ("nbin_angle", C.c_uint32), ("dist_cutoff", C.c_float),
                ("hash_type", C.c_uint32), ("n_multiple_bins", C.c_uint32),
                ("multiple_bins", C.c_uint32 * (2 * MAX_MULTIPLE_BINS))]

    def __init__(self, nbin_dist=0, nbin_angle=0, dist_cutoff=20.0, hash_type=0, multiple_bins=()):
        if isinstance(hash_type, str):
            hash_type = HASH_TYPES[hash_type]
        mb = [int(v) for pair in multiple_bins for v in pair]
        if len(mb) > 2 * MAX_MULTIPLE_BINS:
            raise ValueError("at most %d (dist, angle) pairs" % MAX_MULTIPLE_BINS)
        arr = (C.c_uint32 * (2 * MAX_MULTIPLE_BINS))(*mb)
        super().__init__(nbin_dist, nbin_angle, dist_cutoff, hash_type, len(mb) // 2, arr)


class _IndexBuffers(C.Structure):
    _fields_ = [("count", C.c_uint64), ("hashes", C.POINTER(C.c_uint32)), ("offsets", C.POINTER(C.c_uint64)),
                ("value_bytes", C.c_uint64), ("values", C.POINTER(C.c_uint8))]


class _Query(C.Structure):
    _fields_ = [("n_hashes", C.c_uint32), ("hashes", VP), ("edge_of_hash", VP), ("n_edges", C.c_uint32),
                ("edge_node", VP), ("n_nodes", C.c_uint32), ("expected_node_count", C.c_uint32), ("edge_group", VP)]


class PrefilterParams(C.Structure):
    """count_query arguments + StructureFilter + --top; defaults = reference CLI defaults (src/cli/main.rs:49-97)"""
    _fields_ = [("sampling_ratio", C.c_float), ("sampling_count", C.c_int64), ("freq_filter", C.c_float),
                ("length_penalty", C.c_float), ("total_match_count", C.c_uint64), ("covered_node_count", C.c_uint64),
                ("covered_node_ratio", C.c_float), ("idf_score_cutoff", C.c_float), ("num_res_cutoff", C.c_uint64),
                ("plddt_cutoff", C.c_float), ("top_n", C.c_uint64)]

    def __init__(self, top_n=UINT64_MAX, length_penalty=0.5, num_res_cutoff=50000, **kw):
        super().__init__(-1.0, -1, -1.0, length_penalty, 0, 0, 0.0, 0.0, num_res_cutoff, 0.0, top_n)
        for k, v in kw.items():
            setattr(self, k, v)


class VotesLayout(C.Structure):
    """fd_votes_layout: u32 votes[planes][n_queries][n_structs] in device memory"""
    _fields_ = [("n_queries", C.c_uint32), ("n_structs", C.c_uint32), ("narrow", C.c_uint32), ("edge_words", C.c_uint32),
                ("planes", C.c_uint32), ("first_query", C.c_uint32), ("words", C.c_uint64)]


class DeviceWords:
    """A library-owned device buffer of u32 words exposed through __cuda_array_interface__ (as int32, so that any
    collective library can sum it in place; two's-complement addition is bit-identical to unsigned)."""

    def __init__(self, ptr, words):
        self.__cuda_array_interface__ = {"shape": (int(words),), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class _StructHit(C.Structure):
    _fields_ = [("nid", C.c_uint32), ("match_count", C.c_uint32), ("node_count", C.c_uint32),
                ("edge_count", C.c_uint32), ("idf", C.c_float)]


HIT_DTYPE = np.dtype([("nid", np.uint32), ("match_count", np.uint32), ("node_count", np.uint32),
                      ("edge_count", np.uint32), ("idf", np.float32)])
EDGE_DTYPE = np.dtype([("cand", np.uint32), ("i", np.uint32), ("j", np.uint32), ("hash", np.uint32)])
PAIR_DTYPE = np.dtype([("cand", np.uint32), ("q_index", np.uint32), ("i", np.uint32), ("j", np.uint32),
                       ("k", np.uint32)])


class _RetrievalQuery(C.Structure):
    _fields_ = [("n_hashes", C.c_uint32), ("hashes_sorted", VP), ("n_aa_dist", C.c_uint32), ("aa1", VP), ("aa2", VP),
                ("ca_dist", VP), ("q_index", VP)]


def lib():
    """Loads libfolddisco_b200.so.  Raises FdError if it has not been built (python -m folddisco_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise FdError("%s is missing: build it with `python -m folddisco_b200.build` "
                      "(there is no CPU fallback)" % _SO)
    L = C.CDLL(_SO)

    def sig(name, res, args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args

    PP = C.POINTER
    sig("fd_create", C.c_int, [PP(VP), C.c_int])
    sig("fd_destroy", None, [VP])
    sig("fd_last_error", C.c_char_p, [VP])
    sig("fd_free", None, [VP])
    sig("fd_version", C.c_char_p, [])
    sig("fd_default_host_threads", C.c_int, [])
    sig("fd_fork", C.c_int, [VP, PP(VP)])
    sig("fd_fork_refresh", C.c_int, [VP, VP])
    sig("fd_lane", C.c_int, [VP, C.c_int, PP(VP)])
    sig("fd_lanes_fold_stats", None, [VP])
    sig("fd_kernel_launches", C.c_uint64, [VP])
    sig("fd_stage_ms", C.c_double, [VP, C.c_char_p])
    sig("fd_stage_launches", C.c_uint64, [VP, C.c_char_p])
    sig("fd_typed_hash_host", C.c_int64, [VP, VP, VP, VP, VP, C.c_uint64, PP(HashParams), VP, C.c_int64])
    sig("fd_typed_is_symmetric_host", C.c_int, [C.c_uint32, C.c_uint32])
    sig("fd_hash_structures", C.c_int, [VP, PP(_StructBatch), PP(HashParams), PP(PP(C.c_uint32)), PP(PP(C.c_uint64))])
    sig("fd_build_postings", C.c_int, [VP, VP, VP, C.c_uint64, C.c_uint64, PP(_IndexBuffers)])
    sig("fd_build_index", C.c_int, [VP, PP(_StructBatch), PP(HashParams), C.c_uint64, C.c_uint64, C.c_uint64,
                                    PP(_IndexBuffers)])
    sig("fd_free_index_buffers", None, [PP(_IndexBuffers)])
    sig("fd_index_attach", C.c_int, [VP, VP, VP, C.c_uint64, VP, C.c_uint64, C.c_uint64, VP, VP])
    sig("fd_posting_counts", C.c_int, [VP, VP, C.c_uint64, VP])
    sig("fd_get_entries", C.c_int, [VP, C.c_uint32, PP(PP(C.c_uint64)), PP(C.c_uint64)])
    sig("fd_count_query_batch", C.c_int, [VP, PP(_Query), C.c_uint32, PP(PrefilterParams), PP(PP(_StructHit)),
                                          PP(PP(C.c_uint64))])
    sig("fd_count_query_batch_ex", C.c_int, [VP, PP(_Query), C.c_uint32, PP(PrefilterParams), VP, C.c_uint64,
                                             PP(PP(_StructHit)), PP(PP(C.c_uint64))])
    sig("fd_last_posting_bytes", C.c_uint64, [VP])
    sig("fd_comm_unique_id", C.c_int, [VP])
    sig("fd_comm_init", C.c_int, [VP, VP, C.c_int, C.c_int])
    sig("fd_comm_destroy", None, [VP])
    sig("fd_comm_rank", C.c_int, [VP])
    sig("fd_comm_world", C.c_int, [VP])
    sig("fd_comm_allgather", C.c_int, [VP, VP, C.c_uint64, VP])
    sig("fd_comm_allreduce_u32", C.c_int, [VP, VP, C.c_uint64])
    sig("fd_comm_barrier", C.c_int, [VP])
    sig("fd_count_query_sharded", C.c_int, [VP, PP(_Query), C.c_uint32, PP(PrefilterParams), VP, C.c_uint64, C.c_uint64,
                                            VP, PP(PP(_StructHit)), PP(PP(C.c_uint64))])
    sig("fd_last_exchange_bytes", C.c_uint64, [VP])
    sig("fd_votes_scan", C.c_int, [VP, PP(_Query), C.c_uint32, PP(PrefilterParams), PP(VotesLayout), PP(VP)])
    sig("fd_votes_scan_sparse", C.c_int, [VP, PP(_Query), C.c_uint32, PP(PrefilterParams), VP, C.c_uint32,
                                          PP(VotesLayout), PP(VP), VP, VP])
    sig("fd_votes_merge_begin", C.c_int, [VP, PP(VotesLayout), C.c_uint32, C.c_uint32, PP(VotesLayout), PP(VP)])
    sig("fd_votes_apply", C.c_int, [VP, PP(VotesLayout), VP, VP, C.c_uint64])
    sig("fd_votes_select", C.c_int, [VP, PP(_Query), C.c_uint32, PP(PrefilterParams), PP(VotesLayout), VP, C.c_uint32,
                                     C.c_uint32, PP(PP(_StructHit)), PP(PP(C.c_uint64))])
    sig("fd_store_attach", C.c_int, [VP, PP(_StructBatch)])
    sig("fd_store_build_pair_table", C.c_int, [VP, PP(HashParams), C.c_uint64, PP(C.c_uint64)])
    sig("fd_candidate_edges_batch", C.c_int, [VP, PP(_RetrievalQuery), C.c_uint32, VP, VP, C.c_uint64, PP(HashParams),
                                              C.c_float, PP(VP), PP(C.c_uint64), PP(VP), PP(C.c_uint64)])
    sig("fd_kabsch_batch", C.c_int, [VP, VP, VP, VP, C.c_uint32, VP, VP, VP])
    sig("fd_math_probe", C.c_int, [VP, C.c_int, VP, VP, C.c_uint64, VP])
    sig("fd_math_host", None, [C.c_int, VP, VP, C.c_uint64, VP])
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(VP)


class StructBatch:
    """A batch of CompactStructures as one SoA (fd_struct_batch).  Keeps the numpy arrays alive."""

    def __init__(self, row_offsets, n_xyz, ca_xyz, cb_xyz, aa, cb_valid=None):
        self.row_offsets = np.ascontiguousarray(row_offsets, np.uint64)
        self.n_xyz = np.ascontiguousarray(n_xyz, np.float32).reshape(-1, 3)
        self.ca_xyz = np.ascontiguousarray(ca_xyz, np.float32).reshape(-1, 3)
        self.cb_xyz = np.ascontiguousarray(cb_xyz, np.float32).reshape(-1, 3)
        self.aa = np.ascontiguousarray(aa, np.uint8)
        self.cb_valid = None if cb_valid is None else np.ascontiguousarray(cb_valid, np.uint8)
        R = int(self.row_offsets[-1])
        assert len(self.aa) == R and len(self.ca_xyz) == R and len(self.n_xyz) == R and len(self.cb_xyz) == R
        self.c = _StructBatch(len(self.row_offsets) - 1, _ptr(self.row_offsets), _ptr(self.n_xyz), _ptr(self.ca_xyz),
                              _ptr(self.cb_xyz), _ptr(self.aa), _ptr(self.cb_valid))

    @classmethod
    def from_list(cls, structs):
        """structs: iterable of dicts with n_xyz, ca_xyz, cb_xyz, aa[, cb_valid]"""
        structs = list(structs)
        ro = np.zeros(len(structs) + 1, np.uint64)
        ro[1:] = np.cumsum([len(s["aa"]) for s in structs])
        cat = lambda k, shape, dt: (np.concatenate([np.asarray(s[k], dt).reshape(shape) for s in structs])
                                    if structs else np.zeros(shape, dt).reshape(shape)[:0])
        cbv = None
        if any(s.get("cb_valid") is not None for s in structs):
            cbv = np.concatenate([np.asarray(s["cb_valid"], np.uint8) if s.get("cb_valid") is not None
                                  else np.ones(len(s["aa"]), np.uint8) for s in structs])
        return cls(ro, cat("n_xyz", (-1, 3), np.float32), cat("ca_xyz", (-1, 3), np.float32),
                   cat("cb_xyz", (-1, 3), np.float32), cat("aa", (-1,), np.uint8), cbv)

    @property
    def n_structs(self):
        return len(self.row_offsets) - 1

    @property
    def nres(self):
        return np.diff(self.row_offsets).astype(np.uint32)


class IndexBuffers:
    """hashes / offsets / values arrays of the reference on-disk index (copied out of the library)."""

    def __init__(self, hashes, offsets, values):
        self.hashes, self.offsets, self.values = hashes, offsets, values

    @property
    def count(self):
        return len(self.hashes)

    @property
    def value_bytes(self):
        return len(self.values)

    def offset_file_bytes(self):
        """PREFIX.offset payload (src/index/indextable.rs:297-326)"""
        return (np.uint64(self.count).tobytes() + self.hashes.astype("<u4").tobytes()
                + self.offsets.astype("<u8").tobytes())


def _take(ptr, n, dtype):
    """copy n items out of a library-allocated buffer and fd_free it"""
    if n:
        arr = np.frombuffer((C.c_char * (n * np.dtype(dtype).itemsize)).from_address(
            C.cast(ptr, VP).value), dtype=dtype, count=n).copy()
    else:
        arr = np.zeros(0, dtype)
    lib().fd_free(C.cast(ptr, VP))
    return arr


class Context:
    """fd_ctx: one CUDA device, one host thread."""

    def __init__(self, device=0):
        self.h = VP()
        rc = lib().fd_create(C.byref(self.h), device)
        if rc != 0:
            msg = lib().fd_last_error(None).decode()
            self.h = None
            raise FdError("fd_create failed (%d): %s" % (rc, msg))
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            lib().fd_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise FdError("%s failed (%d): %s" % (what, rc, lib().fd_last_error(self.h).decode()))

    # ---- stats ----
    @property
    def kernel_launches(self):
        return lib().fd_kernel_launches(self.h)

    def stage_ms(self, stage):
        return lib().fd_stage_ms(self.h, stage.encode())

    def stage_launches(self, stage):
        return lib().fd_stage_launches(self.h, stage.encode())

    @property
    def last_posting_bytes(self):
        return lib().fd_last_posting_bytes(self.h)

    # ---- (i) index build ----
    def hash_structures(self, batch, params=None):
        """-> (hashes u32[], row_offsets u64[S+1]): per structure the sorted unique hashes"""
        params = params or HashParams()
        ph, pr = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint64)()
        self._check(lib().fd_hash_structures(self.h, C.byref(batch.c), C.byref(params), C.byref(ph), C.byref(pr)),
                    "fd_hash_structures")
        ro = _take(pr, batch.n_structs + 1, np.uint64)
        return _take(ph, int(ro[-1]), np.uint32), ro

    def _take_index(self, b):
        out = IndexBuffers(
            np.ctypeslib.as_array(b.hashes, (b.count,)).copy() if b.count else np.zeros(0, np.uint32),
            np.ctypeslib.as_array(b.offsets, (b.count + 1,)).copy(),
            np.ctypeslib.as_array(b.values, (b.value_bytes,)).copy() if b.value_bytes else np.zeros(0, np.uint8))
        lib().fd_free_index_buffers(C.byref(b))
        return out

    def build_postings(self, hashes, row_offsets, first_id=0):
        hashes = np.ascontiguousarray(hashes, np.uint32)
        row_offsets = np.ascontiguousarray(row_offsets, np.uint64)
        b = _IndexBuffers()
        self._check(lib().fd_build_postings(self.h, _ptr(hashes), _ptr(row_offsets), len(row_offsets) - 1, first_id,
                                            C.byref(b)), "fd_build_postings")
        return self._take_index(b)

    def build_index(self, batch, params=None, first_id=0, hash_lo=0, hash_hi=1 << 32):
        params = params or HashParams()
        b = _IndexBuffers()
        self._check(lib().fd_build_index(self.h, C.byref(batch.c), C.byref(params), first_id, hash_lo, hash_hi,
                                         C.byref(b)), "fd_build_index")
        return self._take_index(b)

    # ---- (ii) query ----
    def index_attach(self, index, nres, plddt=None):
        hashes = np.ascontiguousarray(index.hashes, np.uint32)
        offsets = np.ascontiguousarray(index.offsets, np.uint64)
        values = np.ascontiguousarray(index.values, np.uint8)
        nres = np.ascontiguousarray(nres, np.uint32)
        plddt = np.zeros(len(nres), np.float32) if plddt is None else np.ascontiguousarray(plddt, np.float32)
        self._check(lib().fd_index_attach(self.h, _ptr(hashes), _ptr(offsets), len(hashes), _ptr(values), len(values),
                                          len(nres), _ptr(nres), _ptr(plddt)), "fd_index_attach")
        self.n_structs = len(nres)

    def posting_counts(self, hashes):
        hashes = np.ascontiguousarray(hashes, np.uint32)
        out = np.zeros(len(hashes), np.uint32)
        self._check(lib().fd_posting_counts(self.h, _ptr(hashes), len(hashes), _ptr(out)), "fd_posting_counts")
        return out

    def get_entries(self, h):
        p, n = C.POINTER(C.c_uint64)(), C.c_uint64()
        self._check(lib().fd_get_entries(self.h, int(h), C.byref(p), C.byref(n)), "fd_get_entries")
        return _take(p, n.value, np.uint64)

    def _query_array(self, queries):
        nq = len(queries)
        arr = (_Query * max(nq, 1))()
        keep = []
        for k, q in enumerate(queries):
            h = np.ascontiguousarray(q["hashes"], np.uint32)
            e = np.ascontiguousarray(q["edge_of_hash"], np.uint16)
            en = np.ascontiguousarray(q["edge_node"], np.uint16)
            keep += [h, e, en]
            eg = None
            if q.get("edge_group") is not None:
                eg = np.ascontiguousarray(q["edge_group"], np.uint16)
                keep.append(eg)
            arr[k] = _Query(len(h), _ptr(h), _ptr(e), len(en), _ptr(en), int(q["n_nodes"]),
                            int(q.get("expected_node_count", q["n_nodes"])), _ptr(eg))
        return arr, keep

    def count_query_batch(self, queries, params=None, global_counts=None, global_n_structs=0):
        """queries: list of dicts {hashes u32[], edge_of_hash u16[], edge_node u16[], n_nodes, expected_node_count
        [, edge_group u16[]]} -> list of structured arrays (HIT_DTYPE), one per query, idf descending / nid ascending.
        global_counts (u32, one per query hash in batch order) + global_n_structs: the attached index is one id-range
        shard of a larger database (fd_count_query_batch_ex)"""
        params = params or PrefilterParams()
        nq = len(queries)
        arr, keep = self._query_array(queries)
        ph, po = C.POINTER(_StructHit)(), C.POINTER(C.c_uint64)()
        if global_counts is not None:
            gc = np.ascontiguousarray(global_counts, np.uint32)
            assert len(gc) == sum(len(q["hashes"]) for q in queries)
            self._check(lib().fd_count_query_batch_ex(self.h, arr, nq, C.byref(params), _ptr(gc), int(global_n_structs),
                                                      C.byref(ph), C.byref(po)), "fd_count_query_batch_ex")
        else:
            self._check(lib().fd_count_query_batch(self.h, arr, nq, C.byref(params), C.byref(ph), C.byref(po)),
                        "fd_count_query_batch")
        off = _take(po, nq + 1, np.uint64)
        hits = _take(ph, int(off[-1]), HIT_DTYPE)
        return [hits[int(off[k]):int(off[k + 1])] for k in range(nq)]

    # ---- multi-GPU: NCCL inside the library, id-range shards ----
    @staticmethod
    def comm_unique_id():
        """128 bytes from ncclGetUniqueId: rank 0 creates them, every rank passes them to comm_init"""
        buf = np.zeros(COMM_ID_BYTES, np.uint8)
        if lib().fd_comm_unique_id(_ptr(buf)) != 0:
            raise FdError("fd_comm_unique_id failed")
        return buf

    def comm_init(self, unique_id, rank, world):
        uid = np.ascontiguousarray(unique_id, np.uint8)
        assert len(uid) == COMM_ID_BYTES
        self._check(lib().fd_comm_init(self.h, _ptr(uid), int(rank), int(world)), "fd_comm_init")

    @property
    def comm_rank(self):
        return lib().fd_comm_rank(self.h)

    @property
    def comm_world(self):
        return lib().fd_comm_world(self.h)

    def comm_allreduce_u32(self, a):
        a = np.ascontiguousarray(a, np.uint32).copy()
        self._check(lib().fd_comm_allreduce_u32(self.h, _ptr(a), len(a)), "fd_comm_allreduce_u32")
        return a

    def comm_barrier(self):
        self._check(lib().fd_comm_barrier(self.h), "fd_comm_barrier")

    def count_query_sharded(self, queries, params, global_counts, global_n_structs, first_id, slice_begin):
        """fd_count_query_sharded: queries = the WHOLE batch; -> hits of this rank's own queries (global ids)"""
        nq = len(queries)
        arr, keep = self._query_array(queries)
        gc = np.ascontiguousarray(global_counts, np.uint32)
        sb = np.ascontiguousarray(slice_begin, np.uint32)
        ph, po = C.POINTER(_StructHit)(), C.POINTER(C.c_uint64)()
        self._check(lib().fd_count_query_sharded(self.h, arr, nq, C.byref(params), _ptr(gc), int(global_n_structs),
                                                 int(first_id), _ptr(sb), C.byref(ph), C.byref(po)),
                    "fd_count_query_sharded")
        r = self.comm_rank
        n = int(sb[r + 1] - sb[r])
        off = _take(po, n + 1, np.uint64)
        hits = _take(ph, int(off[-1]), HIT_DTYPE)
        return [hits[int(off[k]):int(off[k + 1])] for k in range(n)]

    @property
    def last_exchange_bytes(self):
        return lib().fd_last_exchange_bytes(self.h)

    # ---- multi-GPU: partial votes of a hash-range shard ----
    def votes_scan(self, queries, params=None):
        """-> (VotesLayout, device pointer) : partial votes of the attached shard for the whole batch"""
        params = params or PrefilterParams()
        arr, keep = self._query_array(queries)
        lay, ptr = VotesLayout(), VP()
        self._check(lib().fd_votes_scan(self.h, arr, len(queries), C.byref(params), C.byref(lay), C.byref(ptr)),
                    "fd_votes_scan")
        return lay, ptr.value

    def votes_select(self, queries, layout, d_votes, q_begin, q_end, params=None):
        params = params or PrefilterParams()
        arr, keep = self._query_array(queries)
        ph, po = C.POINTER(_StructHit)(), C.POINTER(C.c_uint64)()
        self._check(lib().fd_votes_select(self.h, arr, len(queries), C.byref(params), C.byref(layout), VP(d_votes),
                                          q_begin, q_end, C.byref(ph), C.byref(po)), "fd_votes_select")
        n = q_end - q_begin
        off = _take(po, n + 1, np.uint64)
        hits = _take(ph, int(off[-1]), HIT_DTYPE)
        return [hits[int(off[k]):int(off[k + 1])] for k in range(n)]

    def store_attach(self, batch):
        self._check(lib().fd_store_attach(self.h, C.byref(batch.c)), "fd_store_attach")

    def store_build_pair_table(self, params=None, max_bytes=0):
        """fd_store_build_pair_table -> bytes of the table, or None when it exceeds max_bytes (re-hash path stays)"""
        params = params or HashParams()
        nbytes = C.c_uint64()
        rc = lib().fd_store_build_pair_table(self.h, C.byref(params), int(max_bytes), C.byref(nbytes))
        if rc == -5 and max_bytes:
            return None
        self._check(rc, "fd_store_build_pair_table")
        return int(nbytes.value)

    def candidate_edges_batch(self, rqueries, cand_query, cand_nid, params=None, ca_dist_cutoff=1.0):
        """rqueries: list of dicts {hashes_sorted, aa1, aa2, ca_dist, q_index}; -> (edges EDGE_DTYPE[], pairs PAIR_DTYPE[])"""
        params = params or HashParams()
        nq = len(rqueries)
        arr = (_RetrievalQuery * max(nq, 1))()
        keep = []
        for k, q in enumerate(rqueries):
            h = np.ascontiguousarray(q["hashes_sorted"], np.uint32)
            a1 = np.ascontiguousarray(q["aa1"], np.uint8)
            a2 = np.ascontiguousarray(q["aa2"], np.uint8)
            d = np.ascontiguousarray(q["ca_dist"], np.float32)
            qi = np.ascontiguousarray(q["q_index"], np.uint32)
            keep += [h, a1, a2, d, qi]
            arr[k] = _RetrievalQuery(len(h), _ptr(h), len(a1), _ptr(a1), _ptr(a2), _ptr(d), _ptr(qi))
        cq = np.ascontiguousarray(cand_query, np.uint32)
        cn = np.ascontiguousarray(cand_nid, np.uint32)
        pe, pp, ne, np_ = VP(), VP(), C.c_uint64(), C.c_uint64()
        self._check(lib().fd_candidate_edges_batch(self.h, arr, nq, _ptr(cq), _ptr(cn), len(cq), C.byref(params),
                                                   ca_dist_cutoff, C.byref(pe), C.byref(ne), C.byref(pp),
                                                   C.byref(np_)), "fd_candidate_edges_batch")
        return _take(pe, ne.value, EDGE_DTYPE), _take(pp, np_.value, PAIR_DTYPE)

    def kabsch_batch(self, mov, ref, pt_offsets):
        """rotate mov[a] onto ref[a]; -> (rmsd[n], U[n,3,3], t[n,3])"""
        mov = np.ascontiguousarray(mov, np.float32).reshape(-1, 3)
        ref = np.ascontiguousarray(ref, np.float32).reshape(-1, 3)
        po = np.ascontiguousarray(pt_offsets, np.uint32)
        n = len(po) - 1
        rmsd, U, t = np.zeros(n, np.float32), np.zeros((n, 3, 3), np.float32), np.zeros((n, 3), np.float32)
        self._check(lib().fd_kabsch_batch(self.h, _ptr(mov), _ptr(ref), _ptr(po), n, _ptr(rmsd), _ptr(U), _ptr(t)),
                    "fd_kabsch_batch")
        return rmsd, U, t

    # ---- probes ----
    def math_probe(self, op, a, b=None):
        a = np.ascontiguousarray(a, np.float32)
        b = None if b is None else np.ascontiguousarray(b, np.float32)
        out = np.zeros(len(a), np.float32)
        self._check(lib().fd_math_probe(self.h, op, _ptr(a), _ptr(b), len(a), _ptr(out)), "fd_math_probe")
        return out


def math_host(op, a, b=None):
    a = np.ascontiguousarray(a, np.float32)
    b = None if b is None else np.ascontiguousarray(b, np.float32)
    out = np.zeros(len(a), np.float32)
    lib().fd_math_host(op, _ptr(a), _ptr(b), len(a), _ptr(out))
    return out


def typed_hash_host(n_xyz, ca_xyz, cb_xyz, aa, cb_valid=None, params=None):
    """every ordered residue pair of one structure hashed on the host by csrc/fd_hashtypes.cuh (parity probe of the
    other encodings / `--multiple-bins`): pairs in row-major order, the bin pairs of one residue pair in list order"""
    n_xyz, ca_xyz, cb_xyz = (np.ascontiguousarray(a, np.float32) for a in (n_xyz, ca_xyz, cb_xyz))
    aa = np.ascontiguousarray(aa, np.uint8)
    cbv = None if cb_valid is None else np.ascontiguousarray(cb_valid, np.uint8)
    params = params or HashParams()
    n = lib().fd_typed_hash_host(_ptr(n_xyz), _ptr(ca_xyz), _ptr(cb_xyz), _ptr(aa), _ptr(cbv), len(aa), C.byref(params),
                                 None, 0)
    if n < 0:
        raise FdError("fd_typed_hash_host: parameters refused")
    out = np.zeros(max(n, 1), np.uint32)
    lib().fd_typed_hash_host(_ptr(n_xyz), _ptr(ca_xyz), _ptr(cb_xyz), _ptr(aa), _ptr(cbv), len(aa), C.byref(params),
                             _ptr(out), n)
    return out[:n]


def typed_is_symmetric_host(hash_type, h):
    return lib().fd_typed_is_symmetric_host(hash_type, int(h))
