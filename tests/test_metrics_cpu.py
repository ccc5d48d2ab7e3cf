"""Similarity metrics of a superposed match: the header the device kernel compiles
(csrc/fd_metrics.cuh), built for the host (fd_metrics_host), against the oracle's restatement of
src/structure/metrics.rs:44-345 (the n x n matrix version) -- and the reference's one asserted known answer
(metrics.rs:370-386: identical point sets give TM = GDT-TS = GDT-HA = 1, Chamfer = Hausdorff = 0)."""
import ctypes as C

import numpy as np

import oracle_lib as O


def _host_metrics(ref, mov, U, t):
    import folddisco_b200 as fd
    L = fd.lib()
    L.fd_metrics_host.restype = None
    ref = np.ascontiguousarray(ref, np.float32).reshape(-1)
    mov = np.ascontiguousarray(mov, np.float32).reshape(-1)
    U = np.ascontiguousarray(U, np.float32).reshape(-1)
    t = np.ascontiguousarray(t, np.float32).reshape(-1)
    out = np.zeros(5, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.fd_metrics_host(p(ref), p(mov), C.c_uint32(len(ref) // 3), p(U), p(t), p(out))
    return out


def test_identical_points_reference_known_answer():
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    for m in (_host_metrics(pts, pts, np.eye(3), np.zeros(3)), O.similarity_metrics(pts, pts, np.eye(3), np.zeros(3))):
        assert abs(m[0] - 1) < 1e-6 and abs(m[1] - 1) < 1e-6 and abs(m[2] - 1) < 1e-6
        assert abs(m[3]) < 1e-6 and abs(m[4]) < 1e-6


def test_host_build_equals_oracle_on_random_superpositions():
    rng = np.random.default_rng(7)
    seen_partial_gdt = False
    for n in (4, 6, 8, 16, 22, 24, 32, 64, 130):
        for noise in (0.05, 0.4, 1.5, 4.0):
            ref = rng.normal(0, 6, (n, 3)).astype(np.float32)
            q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            if np.linalg.det(q) < 0:
                q[:, 0] = -q[:, 0]
            mov = ((ref + rng.normal(0, noise, (n, 3))) @ q.T + rng.normal(0, 20, 3)).astype(np.float32)
            U, t, rmsd = O.kabsch(mov, ref)
            a, b = _host_metrics(ref, mov, U, t), O.similarity_metrics(ref, mov, U, t)
            assert np.allclose(a, b, rtol=1e-6, atol=1e-7), (n, noise, a, b)
            assert a[1] == b[1] and a[2] == b[2]            # GDT are counts: exact
            assert 0 < a[0] <= 1 and 0 <= a[2] <= a[1] <= 1 and 0 <= a[3] <= a[4]
            seen_partial_gdt |= 0 < a[2] < 1
    assert seen_partial_gdt


def test_distance_quirk_is_reproduced():
    """metrics.rs:141-147, 160-166 feed a DISTANCE where the formulas expect a squared distance: a point 3 A away counts
    for the 2 A cutoff of GDT-TS (3 <= 2^2) -- kept, because the columns must equal the reference's."""
    ref = np.array([[0, 0, 0], [10, 0, 0]], np.float32)
    mov = np.array([[3, 0, 0], [10, 0, 0]], np.float32)
    m = _host_metrics(ref, mov, np.eye(3), np.zeros(3))
    assert np.array_equal(m, O.similarity_metrics(ref, mov, np.eye(3), np.zeros(3)))
    assert abs(m[1] - (1 + 2 + 2 + 2) / 8) < 1e-7      # cutoffs 1, 2, 4, 8 -> squared 1, 4, 16, 64: d = 3 passes three
    assert abs(m[0] - (1 / (1 + 3 / 0.25) + 1) / 2) < 1e-6  # d0 = 0.5: 1 / (1 + d / d0^2), d not squared
    assert abs(m[3] - 1.5) < 1e-6 and abs(m[4] - 3.0) < 1e-6
