"""The trigonometry-free route of the geometric hash (csrc/fd_geom.cuh: pair_hash_fast) against the exact route
(binary64 sin/cos/acos/atan2, the one pinned to the reference's KATs) on the host build of the same header: wherever
the fast route answers, it must give the exact hash; it may only decline.  The kernels use fast-then-exact
(pair_hash_auto), so this is what keeps K1 / K4 / K6a bit-exact (reference: src/geometry/pdb_tr.rs:21-75,
src/structure/core.rs:378-403)."""
import ctypes as C
import os

import numpy as np

import fixtures as F


class _HP_unused(C.Structure):
    _fields_ = [("nbin_dist", C.c_uint32), ("nbin_angle", C.c_uint32), ("dist_cutoff", C.c_float)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _compare(db, nbin_dist, nbin_angle, max_structs=None):
    import folddisco_b200 as fd
    L = fd.lib()
    L.fd_pair_hash_host.restype = None
    ro = np.asarray(db["row_offsets"], np.uint64)
    n_structs = len(ro) - 1 if max_structs is None else min(max_structs, len(ro) - 1)
    total = declined = mismatched = 0
    for s in range(n_structs):
        a, b = int(ro[s]), int(ro[s + 1])
        n = np.ascontiguousarray(np.asarray(db["n_xyz"], np.float32).reshape(-1, 3)[a:b])
        ca = np.ascontiguousarray(np.asarray(db["ca_xyz"], np.float32).reshape(-1, 3)[a:b])
        cb = np.ascontiguousarray(np.asarray(db["cb_xyz"], np.float32).reshape(-1, 3)[a:b])
        aa = np.ascontiguousarray(np.asarray(db["aa"], np.uint8)[a:b])
        d = np.linalg.norm(ca[:, None, :] - ca[None, :, :], axis=2)
        i, j = np.nonzero((d <= 20.0) & ~np.eye(len(ca), dtype=bool))
        i, j = i.astype(np.uint32), j.astype(np.uint32)
        exact, fast = np.zeros(len(i), np.uint32), np.zeros(len(i), np.uint32)
        dec = np.zeros(len(i), np.uint8)
        hp = fd.HashParams(nbin_dist, nbin_angle, 20.0)
        L.fd_pair_hash_host(_p(n), _p(ca), _p(cb), _p(aa), _p(i), _p(j), C.c_uint64(len(i)), C.byref(hp), _p(exact),
                            _p(fast), _p(dec))
        ok = dec == 0
        mismatched += int((exact[ok] != fast[ok]).sum())
        declined += int(dec.sum())
        total += len(i)
    return total, declined, mismatched


def _templates():
    z = np.load(os.path.join(F.GOLDEN, "templates.npz"))
    return dict(row_offsets=z["offsets"], n_xyz=z["n_xyz"], ca_xyz=z["ca_xyz"], cb_xyz=z["cb_xyz"], aa=z["aa"])


def test_fast_route_equals_exact_on_shipped_structures():
    total, declined, mismatched = _compare(_templates(), 0, 0)  # every pair within 20 A of the 27 shipped files
    assert total > 2_000_000 and mismatched == 0
    assert declined < total // 200  # the fast route answers > 99.5 % of the pairs


def test_fast_route_equals_exact_other_bins_and_synthetic():
    from folddisco_b200 import synth
    for nbd, nba in ((8, 3), (16, 2), (4, 1)):
        total, declined, mismatched = _compare(_templates(), nbd, nba, max_structs=8)
        assert mismatched == 0 and total > 0
    db = synth.generate(40, 11, mean_len=250.0, max_len=600)
    total, declined, mismatched = _compare(db, 0, 0)
    assert total > 300_000 and mismatched == 0


def test_fast_route_declines_degenerate_geometry():
    """collinear / coincident atoms (NaN angles in the reference): the fast route must decline, never guess"""
    n = np.zeros((4, 3), np.float32)
    ca = np.array([[0, 0, 0], [5, 0, 0], [10, 0, 0], [15, 0, 0]], np.float32)
    cb = ca.copy()  # CB on CA: zero-length CA->CB vectors
    aa = np.zeros(4, np.uint8)
    db = dict(row_offsets=np.array([0, 4], np.uint64), n_xyz=n, ca_xyz=ca, cb_xyz=cb, aa=aa)
    total, declined, mismatched = _compare(db, 0, 0)
    assert total == 12 and declined == total and mismatched == 0
