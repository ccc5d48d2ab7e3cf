"""ctypes binding of the CPU oracle (oracle/fd_oracle.h).  TEST INFRASTRUCTURE ONLY.

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm as the
checker.  Nothing under folddisco_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "libfd_oracle.so")

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
VP = C.c_void_p
UINT64_MAX = (1 << 64) - 1


class CountParams(C.Structure):
    _fields_ = [
        ("sampling_ratio", C.c_float),
        ("sampling_count", C.c_int64),
        ("freq_filter", C.c_float),
        ("length_penalty", C.c_float),
        ("total_match_count", C.c_uint64),
        ("covered_node_count", C.c_uint64),
        ("covered_node_ratio", C.c_float),
        ("idf_score_cutoff", C.c_float),
        ("num_res_cutoff", C.c_uint64),
        ("plddt_cutoff", C.c_float),
        ("expected_node_count", C.c_uint64),
        ("top_n", C.c_uint64),
        ("apply_filter_and_sort", C.c_int),
    ]

    @classmethod
    def defaults(cls, expected_node_count=0, top_n=UINT64_MAX, apply=1):
        """Reference CLI defaults (src/cli/main.rs:49-97)."""
        return cls(-1.0, -1, -1.0, 0.5, 0, 0, 0.0, 0.0, 50000, 0.0, expected_node_count, top_n, apply)


def build(force=False):
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_ROOT, "oracle", f)) > os.path.getmtime(_SO)
        for f in ("fd_oracle.cpp", "fd_oracle.h", "fd_oracle_math.h")
    ):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


_REF_SO = os.path.join(_ROOT, "oracle", "_ref", "libfoldcomp_ffi.so")


def build_ref():
    """oracle/_ref/libfoldcomp_ffi.so: the Foldcomp codec the reference vendors, compiled from the reference tree where it
    lies (oracle/Makefile `ref`).  -> path, or None when neither the file nor /root/reference exists"""
    if not os.path.exists(_REF_SO):
        if not os.path.isdir("/root/reference/lib/foldcomp"):
            return None
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    return _REF_SO


class _FczAtom(C.Structure):  # atom_t (lib/foldcomp/foldcompffi.h:8-16)
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("atom", C.c_uint8 * 4), ("atom_idx", C.c_uint64),
                ("chain", C.c_uint8), ("aa", C.c_uint8 * 3), ("res_idx", C.c_uint64), ("bfactor", C.c_float)]


def foldcomp_atoms(data):
    """one Foldcomp entry through the codec itself (create / process / destroy / free, the calls of fcz.rs:80-94) -> the
    atoms dict Structure.from_atoms takes (Atom::from_c: the same fields under the reference's names)"""
    L = C.CDLL(build_ref())
    L.foldcomp_create.restype = C.c_void_p
    L.foldcomp_process.restype = C.POINTER(_FczAtom)
    L.foldcomp_process.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.foldcomp_free.argtypes = [C.POINTER(_FczAtom)]
    L.foldcomp_destroy.argtypes = [C.c_void_p]
    inst = L.foldcomp_create()
    n = C.c_size_t()
    p = L.foldcomp_process(inst, data, len(data), C.byref(n))
    n = n.value
    a = dict(x=np.zeros(n, np.float32), y=np.zeros(n, np.float32), z=np.zeros(n, np.float32),
             atom_name=np.zeros((n, 4), np.uint8), chain=np.zeros(n, np.uint8), res_name=np.zeros((n, 3), np.uint8),
             res_serial=np.zeros(n, np.uint64), b_factor=np.zeros(n, np.float32))
    for i in range(n):
        t = p[i]
        a["x"][i], a["y"][i], a["z"][i], a["b_factor"][i] = t.x, t.y, t.z, t.bfactor
        a["atom_name"][i] = list(t.atom)
        a["res_name"][i] = list(t.aa)
        a["chain"][i], a["res_serial"][i] = t.chain, t.res_idx
    L.foldcomp_destroy(inst)
    L.foldcomp_free(p)
    return a


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())

    def sig(name, res, args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args

    sig("fdo_set_math_mode", None, [C.c_int])
    for n in ("sinf", "cosf", "acosf"):
        sig("fdo_math_" + n, C.c_float, [C.c_float])
    sig("fdo_math_atan2f", C.c_float, [C.c_float, C.c_float])
    sig("fdo_structure_read_pdb", VP, [C.c_char_p])
    sig("fdo_structure_from_atoms", VP, [C.c_int64, f32p, f32p, f32p, u8p, u8p, u8p, u64p, f32p])
    sig("fdo_structure_num_atoms", C.c_int64, [VP])
    sig("fdo_structure_num_residues", C.c_int64, [VP])
    sig("fdo_structure_first_chain", C.c_int, [VP])
    sig("fdo_structure_get_atoms", None, [VP, f32p, f32p, f32p, u8p, u8p, u8p, u64p, f32p])
    sig("fdo_structure_free", None, [VP])
    sig("fdo_compact_build", VP, [VP])
    sig("fdo_compact_from_soa", VP, [C.c_int64, f32p, f32p, f32p, VP, u8p, VP, VP, VP])
    sig("fdo_compact_nres", C.c_int64, [VP])
    sig("fdo_compact_get", None, [VP, f32p, f32p, f32p, u8p, u8p, u8p, u8p, u64p, f32p])
    sig("fdo_compact_avg_plddt", C.c_float, [VP])
    sig("fdo_compact_get_index", C.c_int64, [VP, C.c_uint8, C.c_uint64])
    sig("fdo_compact_free", None, [VP])
    sig("fdo_map_aa_to_u8", C.c_uint8, [C.c_char_p])
    sig("fdo_pair_feature", C.c_int, [VP, C.c_int64, C.c_int64, C.c_float, f32p])
    sig("fdo_perfect_hash", C.c_uint32, [f32p, C.c_uint32, C.c_uint32])
    sig("fdo_hash_is_symmetric", C.c_int, [C.c_uint32])
    sig("fdo_set_hash_type", C.c_int, [C.c_int])
    sig("fdo_get_hash_type", C.c_int, [])
    sig("fdo_set_multiple_bins", None, [C.c_int, VP])
    sig("fdo_pair_feature9", C.c_int, [VP, C.c_int64, C.c_int64, C.c_float, f32p])
    sig("fdo_perfect_hash_raw", C.c_uint32, [f32p, C.c_uint32, C.c_uint32])
    sig("fdo_hash_compact", C.c_int64, [VP, C.c_uint32, C.c_uint32, C.c_float, C.c_int, VP, C.c_int64])
    sig("fdo_index_from_csr", VP, [u32p, u64p, C.c_uint64])
    sig("fdo_index_build", VP, [C.POINTER(VP), C.c_uint64, C.c_uint32, C.c_uint32, C.c_float, C.c_int])
    sig("fdo_index_from_buffers", VP, [u32p, u64p, C.c_uint64, u8p, C.c_uint64])
    sig("fdo_index_load", VP, [C.c_char_p])
    sig("fdo_index_save", C.c_int, [VP, C.c_char_p])
    sig("fdo_index_count", C.c_uint64, [VP])
    sig("fdo_index_value_bytes", C.c_uint64, [VP])
    sig("fdo_index_hashes", C.POINTER(C.c_uint32), [VP])
    sig("fdo_index_offsets", C.POINTER(C.c_uint64), [VP])
    sig("fdo_index_values", C.POINTER(C.c_uint8), [VP])
    sig("fdo_index_get_entries", C.c_int64, [VP, C.c_uint32, VP, C.c_int64])
    sig("fdo_index_free", None, [VP])
    sig("fdo_lookup_save", C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_char_p), u64p, f32p])
    sig("fdo_type_save", C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, C.c_float, C.c_uint64, C.c_uint64, C.c_char_p])
    sig("fdo_parse_query_string", C.c_int64, [C.c_char_p, C.c_uint8, u8p, u64p, i64p, u8p, C.c_int64, C.c_int64])
    sig("fdo_qmap_make", VP, [VP, u8p, u64p, C.c_int64, i64p, i64p, u8p, C.c_uint32, C.c_uint32, f32p, C.c_int,
                              f32p, C.c_int, C.c_float, C.c_int, VP, C.c_float])
    sig("fdo_qmap_size", C.c_int64, [VP])
    sig("fdo_qmap_get", None, [VP, u32p, i64p, i64p, u8p, f32p])
    sig("fdo_qmap_num_indices", C.c_int64, [VP])
    sig("fdo_qmap_residue_count", C.c_int64, [VP])
    sig("fdo_qmap_get_indices", None, [VP, i64p])
    sig("fdo_qmap_free", None, [VP])
    sig("fdo_count_query", VP, [VP, VP, C.c_uint64, u64p, f32p, C.POINTER(CountParams)])
    sig("fdo_hits_size", C.c_int64, [VP])
    sig("fdo_hits_get", None, [VP, u64p, u32p, u32p, u32p, f32p])
    sig("fdo_hits_free", None, [VP])
    sig("fdo_retrieve", VP, [VP, VP, VP, C.c_uint32, C.c_uint32, C.c_float, C.c_float])
    sig("fdo_matches_size", C.c_int64, [VP])
    sig("fdo_matches_num_query", C.c_int64, [VP])
    sig("fdo_matches_get", None, [VP, C.c_int, u8p, u8p, u64p, f32p, f32p, f32p, f32p])
    sig("fdo_matches_get_metrics", None, [VP, C.c_int, f32p])
    sig("fdo_set_partial_fit", None, [C.c_int])
    sig("fdo_lms_qcp", C.c_int64, [C.c_int64, f32p, f32p, f32p, f32p, f32p, i64p, C.c_int64])
    sig("fdo_similarity_metrics", None, [C.c_int64, f32p, f32p, f32p, f32p, f32p])
    sig("fdo_matches_max_node_count", C.c_int64, [VP])
    sig("fdo_matches_min_rmsd", C.c_float, [VP])
    sig("fdo_matches_num_edges", C.c_int64, [VP])
    sig("fdo_matches_get_edges", None, [VP, i64p, i64p, u32p])
    sig("fdo_matches_free", None, [VP])
    sig("fdo_kabsch", C.c_int, [C.c_int64, f32p, f32p, f32p, f32p, C.POINTER(C.c_float)])
    sig("fdo_query_batch", C.c_int64, [C.POINTER(VP), C.POINTER(VP), C.c_int64, VP, C.POINTER(VP), C.c_uint64,
                                       u64p, f32p, C.POINTER(CountParams), C.c_uint32, C.c_uint32, C.c_float,
                                       C.c_float, C.c_int, C.c_int, VP, VP, VP])
    _lib = L
    return L


# ---------------------------------------------------------------------------------------------
# thin Python objects
# ---------------------------------------------------------------------------------------------


class Structure:
    def __init__(self, handle):
        if not handle:
            raise IOError("oracle: could not read structure")
        self.h = handle

    @classmethod
    def read_pdb(cls, path):
        return cls(lib().fdo_structure_read_pdb(os.fsencode(path)))

    @classmethod
    def from_atoms(cls, atoms):
        a = atoms
        return cls(lib().fdo_structure_from_atoms(
            len(a["x"]), a["x"], a["y"], a["z"], np.ascontiguousarray(a["atom_name"]).reshape(-1),
            a["chain"], np.ascontiguousarray(a["res_name"]).reshape(-1), a["res_serial"], a["b_factor"]))

    def atoms(self):
        n = lib().fdo_structure_num_atoms(self.h)
        a = dict(x=np.zeros(n, np.float32), y=np.zeros(n, np.float32), z=np.zeros(n, np.float32),
                 atom_name=np.zeros(n * 4, np.uint8), chain=np.zeros(n, np.uint8),
                 res_name=np.zeros(n * 3, np.uint8), res_serial=np.zeros(n, np.uint64),
                 b_factor=np.zeros(n, np.float32))
        lib().fdo_structure_get_atoms(self.h, a["x"], a["y"], a["z"], a["atom_name"], a["chain"], a["res_name"],
                                      a["res_serial"], a["b_factor"])
        a["atom_name"] = a["atom_name"].reshape(n, 4)
        a["res_name"] = a["res_name"].reshape(n, 3)
        return a

    @property
    def num_residues(self):
        return lib().fdo_structure_num_residues(self.h)

    @property
    def first_chain(self):
        return lib().fdo_structure_first_chain(self.h)

    def compact(self):
        return Compact(lib().fdo_compact_build(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().fdo_structure_free(self.h)
            self.h = None


class Compact:
    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_soa(cls, n_xyz, ca_xyz, cb_xyz, aa, cb_valid=None, chain=None, serial=None, b_factor=None):
        n = len(aa)
        keep = [np.ascontiguousarray(n_xyz, np.float32), np.ascontiguousarray(ca_xyz, np.float32),
                np.ascontiguousarray(cb_xyz, np.float32), np.ascontiguousarray(aa, np.uint8)]
        opt = []
        for arr, dt in ((cb_valid, np.uint8), (chain, np.uint8), (serial, np.uint64), (b_factor, np.float32)):
            opt.append(None if arr is None else np.ascontiguousarray(arr, dt))
        p = [None if a is None else a.ctypes.data_as(VP) for a in opt]
        return cls(lib().fdo_compact_from_soa(n, keep[0], keep[1], keep[2], p[0], keep[3], p[1], p[2], p[3]))

    @property
    def nres(self):
        return lib().fdo_compact_nres(self.h)

    def soa(self):
        n = self.nres
        d = dict(n_xyz=np.zeros((n, 3), np.float32), ca_xyz=np.zeros((n, 3), np.float32),
                 cb_xyz=np.zeros((n, 3), np.float32), cb_valid=np.zeros(n, np.uint8), aa=np.zeros(n, np.uint8),
                 res_name=np.zeros((n, 3), np.uint8), chain=np.zeros(n, np.uint8), serial=np.zeros(n, np.uint64),
                 b_factor=np.zeros(n, np.float32))
        lib().fdo_compact_get(self.h, d["n_xyz"].reshape(-1), d["ca_xyz"].reshape(-1), d["cb_xyz"].reshape(-1),
                              d["cb_valid"], d["aa"], d["res_name"].reshape(-1), d["chain"], d["serial"],
                              d["b_factor"])
        return d

    @property
    def avg_plddt(self):
        return lib().fdo_compact_avg_plddt(self.h)

    def get_index(self, chain, serial):
        return lib().fdo_compact_get_index(self.h, chain, serial)

    def pair_feature(self, i, j, cutoff=20.0):
        out = np.zeros(7, np.float32)
        ok = lib().fdo_pair_feature(self.h, i, j, cutoff, out)
        return out if ok else None

    def hashes(self, nbin_dist=0, nbin_angle=0, cutoff=20.0, sorted_unique=False):
        n = lib().fdo_hash_compact(self.h, nbin_dist, nbin_angle, cutoff, int(sorted_unique), None, 0)
        out = np.zeros(max(n, 1), np.uint32)
        lib().fdo_hash_compact(self.h, nbin_dist, nbin_angle, cutoff, int(sorted_unique), out.ctypes.data_as(VP), n)
        return out[:n]

    def __del__(self):
        if getattr(self, "h", None):
            lib().fdo_compact_free(self.h)
            self.h = None


class hash_mode:
    """with oracle_lib.hash_mode(hash_type, multiple_bins): ... -- selects the oracle's encoding (the reference's
    HashType index: 0 PDBMotif, 1 PDBMotifSinCos, 2 TrRosetta, 3 PDBTrRosetta, 4 PointPairFeature, 7 FolddiscoAngle,
    8 FolddiscoDist) and the --multiple-bins list for the duration of the block; the default is restored on exit."""

    def __init__(self, hash_type=3, multiple_bins=()):
        self.t, self.mb = hash_type, list(multiple_bins)

    def __enter__(self):
        if lib().fdo_set_hash_type(self.t) != 0:
            raise ValueError("hash type %r is not restated by the oracle" % (self.t,))
        a = np.ascontiguousarray(np.array(self.mb, np.uint32).reshape(-1))
        lib().fdo_set_multiple_bins(len(self.mb), a.ctypes.data if len(self.mb) else None)
        return self

    def __exit__(self, *exc):
        lib().fdo_set_hash_type(3)
        lib().fdo_set_multiple_bins(0, None)
        return False


def pair_feature9(compact, i, j, cutoff=20.0):
    f = np.zeros(9, np.float32)
    return f if lib().fdo_pair_feature9(compact.h, i, j, cutoff, f) else None


def perfect_hash_raw(feature9, nbin_dist, nbin_angle):
    return int(lib().fdo_perfect_hash_raw(np.ascontiguousarray(feature9, np.float32), nbin_dist, nbin_angle))


def perfect_hash(feature7, nbin_dist=0, nbin_angle=0):
    return lib().fdo_perfect_hash(np.ascontiguousarray(feature7, np.float32), nbin_dist, nbin_angle)


class Index:
    def __init__(self, handle):
        if not handle:
            raise IOError("oracle: could not create index")
        self.h = handle

    @classmethod
    def from_csr(cls, hashes, row_offsets):
        return cls(lib().fdo_index_from_csr(np.ascontiguousarray(hashes, np.uint32),
                                           np.ascontiguousarray(row_offsets, np.uint64), len(row_offsets) - 1))

    @classmethod
    def build(cls, compacts, nbin_dist=0, nbin_angle=0, cutoff=20.0, threads=1):
        arr = (VP * len(compacts))(*[c.h for c in compacts])
        return cls(lib().fdo_index_build(arr, len(compacts), nbin_dist, nbin_angle, cutoff, threads))

    @classmethod
    def from_buffers(cls, hashes, offsets, values):
        return cls(lib().fdo_index_from_buffers(np.ascontiguousarray(hashes, np.uint32),
                                               np.ascontiguousarray(offsets, np.uint64), len(hashes),
                                               np.ascontiguousarray(values, np.uint8), len(values)))

    @classmethod
    def load(cls, prefix):
        return cls(lib().fdo_index_load(os.fsencode(prefix)))

    def save(self, prefix):
        if lib().fdo_index_save(self.h, os.fsencode(prefix)) != 0:
            raise IOError("oracle: index save failed")

    @property
    def count(self):
        return lib().fdo_index_count(self.h)

    @property
    def value_bytes(self):
        return lib().fdo_index_value_bytes(self.h)

    @property
    def hashes(self):
        n = self.count
        return np.ctypeslib.as_array(lib().fdo_index_hashes(self.h), (n,)).copy() if n else np.zeros(0, np.uint32)

    @property
    def offsets(self):
        return np.ctypeslib.as_array(lib().fdo_index_offsets(self.h), (self.count + 1,)).copy()

    @property
    def values(self):
        n = self.value_bytes
        return np.ctypeslib.as_array(lib().fdo_index_values(self.h), (n,)).copy() if n else np.zeros(0, np.uint8)

    def entries(self, h):
        n = lib().fdo_index_get_entries(self.h, h, None, 0)
        out = np.zeros(max(n, 1), np.uint64)
        lib().fdo_index_get_entries(self.h, h, out.ctypes.data_as(VP), n)
        return out[:n]

    def __del__(self):
        if getattr(self, "h", None):
            lib().fdo_index_free(self.h)
            self.h = None


def parse_query_string(q, default_chain=ord("A")):
    """-> (chains u8[n], serials u64[n], subs list[None | list[int]])"""
    cap = 65536
    chains = np.zeros(cap, np.uint8)
    serials = np.zeros(cap, np.uint64)
    off = np.zeros(cap + 1, np.int64)
    subs = np.zeros(cap * 20, np.uint8)
    n = lib().fdo_parse_query_string(q.encode(), default_chain, chains, serials, off, subs, cap, cap * 20)
    if n < 0:
        raise ValueError("invalid query string %r" % q)
    total = off[n]
    out = []
    for i in range(n):
        if off[i] < 0:
            out.append(None)
        else:
            end = total
            for j in range(i + 1, n):
                if off[j] >= 0:
                    end = off[j]
                    break
            out.append([int(v) for v in subs[off[i]:end]])
    return chains[:n].copy(), serials[:n].copy(), out


class QueryMap:
    def __init__(self, query, chains, serials, subs=None, nbin_dist=0, nbin_angle=0, dist_thr=(0.5,),
                 angle_thr=(5.0,), cutoff=20.0, serial_query=False, index=None, total_structures=0.0):
        n = len(chains)
        subs = subs if subs is not None else [None] * n
        flat, off, end = [], np.full(n + 1, -1, np.int64), np.zeros(n + 1, np.int64)
        for i, s in enumerate(subs):
            if s is not None:
                off[i] = len(flat)
                flat.extend(s)
                end[i] = len(flat)
        flat = np.asarray(flat if flat else [0], np.uint8)
        dt = np.asarray(dist_thr, np.float32)
        at = np.asarray(angle_thr, np.float32)
        self.query = query
        self.h = lib().fdo_qmap_make(query.h, np.ascontiguousarray(chains, np.uint8),
                                     np.ascontiguousarray(serials, np.uint64), n, off, end, flat, nbin_dist,
                                     nbin_angle, dt if len(dt) else np.zeros(1, np.float32), len(dt),
                                     at if len(at) else np.zeros(1, np.float32), len(at), cutoff,
                                     int(serial_query), index.h if index is not None else None,
                                     float(total_structures))

    def entries(self):
        n = lib().fdo_qmap_size(self.h)
        d = dict(hash=np.zeros(n, np.uint32), qi=np.zeros(n, np.int64), qj=np.zeros(n, np.int64),
                 primary=np.zeros(n, np.uint8), idf=np.zeros(n, np.float32))
        if n:
            lib().fdo_qmap_get(self.h, d["hash"], d["qi"], d["qj"], d["primary"], d["idf"])
        return d

    @property
    def residue_count(self):
        """residue_count of query_pdb.rs:355-359 (the denominator of the node-ratio filters)"""
        return lib().fdo_qmap_residue_count(self.h)

    def indices(self):
        n = lib().fdo_qmap_num_indices(self.h)
        out = np.zeros(max(n, 1), np.int64)
        lib().fdo_qmap_get_indices(self.h, out)
        return out[:n]

    def __del__(self):
        if getattr(self, "h", None):
            lib().fdo_qmap_free(self.h)
            self.h = None


def count_query(qmap, index, nres, plddt=None, params=None):
    S = len(nres)
    nres = np.ascontiguousarray(nres, np.uint64)
    plddt = np.zeros(S, np.float32) if plddt is None else np.ascontiguousarray(plddt, np.float32)
    p = params or CountParams.defaults(expected_node_count=qmap.residue_count)
    h = lib().fdo_count_query(qmap.h, index.h, S, nres, plddt, C.byref(p))
    n = lib().fdo_hits_size(h)
    d = dict(nid=np.zeros(n, np.uint64), match_count=np.zeros(n, np.uint32), node_count=np.zeros(n, np.uint32),
             edge_count=np.zeros(n, np.uint32), idf=np.zeros(n, np.float32))
    if n:
        lib().fdo_hits_get(h, d["nid"], d["match_count"], d["node_count"], d["edge_count"], d["idf"])
    lib().fdo_hits_free(h)
    return d


def retrieve(qmap, target, nbin_dist=0, nbin_angle=0, cutoff=20.0, ca_cutoff=1.0, which=0):
    r = lib().fdo_retrieve(qmap.h, qmap.query.h, target.h, nbin_dist, nbin_angle, cutoff, ca_cutoff)
    n = lib().fdo_matches_size(r)
    nq = lib().fdo_matches_num_query(r)
    d = dict(some=np.zeros((n, nq), np.uint8), chain=np.zeros((n, nq), np.uint8), serial=np.zeros((n, nq), np.uint64),
             rmsd=np.zeros(n, np.float32), idf=np.zeros(n, np.float32), U=np.zeros((n, 9), np.float32),
             t=np.zeros((n, 3), np.float32))
    if n:
        lib().fdo_matches_get(r, which, d["some"].reshape(-1), d["chain"].reshape(-1), d["serial"].reshape(-1),
                              d["rmsd"], d["idf"], d["U"].reshape(-1), d["t"].reshape(-1))
    d["metrics"] = np.zeros((n, 5), np.float32)  # tm_score, gdt_ts, gdt_ha, chamfer_distance, hausdorff_distance
    if n:
        lib().fdo_matches_get_metrics(r, which, d["metrics"].reshape(-1))
    d["max_node_count"] = lib().fdo_matches_max_node_count(r)
    d["min_rmsd"] = lib().fdo_matches_min_rmsd(r)
    ne = lib().fdo_matches_num_edges(r)
    ei, ej, eh = np.zeros(max(ne, 1), np.int64), np.zeros(max(ne, 1), np.int64), np.zeros(max(ne, 1), np.uint32)
    if ne:
        lib().fdo_matches_get_edges(r, ei, ej, eh)
    d["edges"] = (ei[:ne], ej[:ne], eh[:ne])
    lib().fdo_matches_free(r)
    return d


class partial_fit:
    """with oracle_lib.partial_fit(): retrieve() superposes like `--partial-fit` (LMS-QCP above three residues)"""

    def __enter__(self):
        lib().fdo_set_partial_fit(1)
        return self

    def __exit__(self, *exc):
        lib().fdo_set_partial_fit(0)
        return False


def lms_qcp(ref, mov):
    """LmsQcpSuperimposer::run: -> (U[3,3], t[3], rms over the inlier core, inlier indices)"""
    ref = np.ascontiguousarray(ref, np.float32).reshape(-1)
    mov = np.ascontiguousarray(mov, np.float32).reshape(-1)
    n = len(ref) // 3
    U, t, rms = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(1, np.float32)
    idx = np.zeros(n, np.int64)
    k = lib().fdo_lms_qcp(n, ref, mov, U, t, rms, idx, n)
    return U.reshape(3, 3), t, float(rms[0]), idx[:k]


def similarity_metrics(ref, mov, U, t):
    """src/structure/metrics.rs over explicit points: -> [tm_score, gdt_ts, gdt_ha, chamfer_distance, hausdorff_distance]"""
    ref = np.ascontiguousarray(ref, np.float32).reshape(-1)
    mov = np.ascontiguousarray(mov, np.float32).reshape(-1)
    out = np.zeros(5, np.float32)
    lib().fdo_similarity_metrics(len(ref) // 3, ref, mov, np.ascontiguousarray(U, np.float32).reshape(-1),
                                 np.ascontiguousarray(t, np.float32).reshape(-1), out)
    return out


def residues_to_string(some, chain, serial):
    return ",".join("%s%d" % (chr(c), s) if f else "_" for f, c, s in zip(some, chain, serial))


def kabsch(x, y):
    """rotate x onto y; -> (U[3,3], t[3], rmsd)"""
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    y = np.ascontiguousarray(y, np.float32).reshape(-1)
    U, t, r = np.zeros(9, np.float32), np.zeros(3, np.float32), C.c_float()
    lib().fdo_kabsch(len(x) // 3, x, y, U, t, C.byref(r))
    return U.reshape(3, 3), t, r.value
