"""LMS-QCP partial superposition (`--partial-fit`): the header the device kernel compiles
(csrc/fd_lmsqcp.cuh), built for the host (fd_lmsqcp_host), against the oracle's restatement of
src/structure/lms_qcp.rs:91-564 -- and the property the reference's own (ignored) test states (:570-586): with one gross
outlier among otherwise exactly superposable pairs the core leaves the outlier out and its RMSD is ~0."""
import ctypes as C

import numpy as np

import oracle_lib as O


def _host_lms(ref, mov):
    import folddisco_b200 as fd
    L = fd.lib()
    L.fd_lmsqcp_host.restype = C.c_int
    ref = np.ascontiguousarray(ref, np.float32).reshape(-1)
    mov = np.ascontiguousarray(mov, np.float32).reshape(-1)
    U, t, rms = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(1, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.fd_lmsqcp_host(p(ref), p(mov), C.c_uint32(len(ref) // 3), p(U), p(t), p(rms))
    return rc, U.reshape(3, 3), t, float(rms[0])


def _pose(rng, pts):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return (pts @ q.T + rng.normal(0, 15, 3)).astype(np.float32)


def test_outlier_is_left_out_of_the_core():
    rng = np.random.default_rng(3)
    ref = rng.normal(0, 5, (12, 3)).astype(np.float32)
    mov = _pose(rng, ref.astype(np.float64))
    mov[7] += np.array([9.0, -7.0, 8.0], np.float32)  # one gross outlier
    U, t, rms, core = O.lms_qcp(ref, mov)
    assert 7 not in core.tolist() and len(core) == 11 and rms < 1e-3
    rc, U2, t2, rms2 = _host_lms(ref, mov)
    assert rc == 0 and rms2 < 1e-3
    assert np.allclose(U, U2, atol=1e-6) and np.allclose(t, t2, atol=1e-4)
    _, _, k_rmsd = O.kabsch(mov, ref)  # the full Kabsch fit pays for the outlier
    assert k_rmsd > 1.0


def test_host_build_equals_oracle():
    rng = np.random.default_rng(11)
    for n in (8, 10, 12, 16, 24, 32, 64, 128):
        for noise, n_out in ((0.05, 0), (0.3, 1), (0.6, 3), (1.5, 0), (3.0, 2)):
            ref = rng.normal(0, 6, (n, 3)).astype(np.float32)
            mov = _pose(rng, ref + rng.normal(0, noise, (n, 3)))
            for k in rng.choice(n, n_out, replace=False):
                mov[k] += rng.normal(0, 8, 3).astype(np.float32)
            U, t, rms, core = O.lms_qcp(ref, mov)
            rc, U2, t2, rms2 = _host_lms(ref, mov)
            assert rc == 0
            assert np.array_equal(U, U2) and np.array_equal(t, t2) and rms == rms2, (n, noise, n_out)
            assert n // 2 <= len(core) <= n and len(set(core.tolist())) == len(core)
    assert _host_lms(np.zeros((2, 3)), np.zeros((2, 3)))[0] == -1 and _host_lms(np.zeros((129, 3)), np.zeros((129, 3)))[0] == -1
