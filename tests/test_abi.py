"""CPU-side checks of the C ABI: the library builds for sm_100a, loads, exports every symbol the header
declares, fails loudly without a GPU, and its host build of the hash math equals the oracle bit for bit."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so():
    from folddisco_b200 import build
    return build.build()


def test_exports_every_declared_symbol(so):
    hdr = open(os.path.join(ROOT, "include", "folddisco_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(fd_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    L = ctypes.CDLL(so)
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    # the host-level header (orchestration mirroring src/lib.rs prelude names, used by the command line)
    hdr = open(os.path.join(ROOT, "include", "folddisco_b200_host.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hnames = set(re.findall(r"\b(fdh_[a-z0-9_]+)\s*\(", hdr))
    assert len(hnames) >= 40
    missing = [n for n in sorted(hnames) if not hasattr(L, n)]
    assert not missing, missing


def test_sass_is_sm100a(so):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:200]


def test_create_fails_loudly_without_gpu(so):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import folddisco_b200 as fd
    with pytest.raises(fd.FdError) as e:
        fd.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_missing_library_raises(monkeypatch, so):
    from folddisco_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "_SO", "/nonexistent/libfolddisco_b200.so")
    with pytest.raises(capi.FdError):
        capi.lib()


def test_host_math_equals_oracle_bit_exact(so):
    """fd_math.cuh compiled for the host == oracle/fd_oracle_math.h (two independent copies of the algorithm)."""
    from folddisco_b200 import capi
    rng = np.random.default_rng(3)
    n = 20000
    a = rng.uniform(-4, 4, n).astype(np.float32)
    u = np.concatenate([rng.uniform(-1, 1, n - 4), [1.0, -1.0, 1.0000001, np.nan]]).astype(np.float32)
    y, x = rng.normal(0, 1, (2, n)).astype(np.float32)
    L = O.lib()
    for op, fn, (p, q) in ((0, L.fdo_math_sinf, (a, None)), (1, L.fdo_math_cosf, (a, None)),
                           (2, L.fdo_math_acosf, (u, None)), (3, L.fdo_math_atan2f, (y, x))):
        got = capi.math_host(op, p, q)
        want = np.array([fn(float(v)) if q is None else fn(float(v), float(w))
                         for v, w in zip(p, p if q is None else q)], np.float32)
        ok = (got == want) | (np.isnan(got) & np.isnan(want))
        assert ok.all(), (op, p[~ok][:5], got[~ok][:5], want[~ok][:5])


def _build_abi_smoke(tmp_path, so):
    """gcc -std=c11 -Werror over both headers, linked against the product library"""
    import subprocess
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(so)
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lfolddisco_b200",
           "-Wl,-rpath," + libdir, "-Wl,-rpath-link,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_headers_compile_as_c_and_link(tmp_path, so):
    """tests/abi_smoke.c: the two headers are valid C11, the struct sizes / offsets that foreign bindings mirror
    (INTEGRATION.md's Rust block, folddisco_b200/capi.py) are what they assert, every entry point it names links; and
    the ctypes mirrors agree with the C sizes.  Without a GPU the program's fd_create fails loudly (exit code 3)."""
    import subprocess
    import ctypes as C
    from folddisco_b200 import capi, host
    exe = _build_abi_smoke(tmp_path, so)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU fallback" in r.stderr
    # the Python mirrors against the sizes asserted in abi_smoke.c
    src = open(os.path.join(ROOT, "tests", "abi_smoke.c")).read()
    sizes = {m.group(1): int(m.group(2)) for m in re.finditer(r"_Static_assert\(sizeof\((\w+)\) == (\d+)", src)}
    mirrors = {"fd_struct_batch": capi._StructBatch, "fd_hash_params": capi.HashParams, "fd_index_buffers": capi._IndexBuffers,
               "fd_query": capi._Query, "fd_prefilter_params": capi.PrefilterParams, "fd_struct_hit": capi._StructHit,
               "fd_votes_layout": capi.VotesLayout, "fd_retrieval_query": capi._RetrievalQuery,
               "fdh_query_params": host._QueryParams, "fdh_search_params": host.SearchParams}
    for name, cls in mirrors.items():
        assert C.sizeof(cls) == sizes[name], name
    assert host.STRUCT_ROW.itemsize == sizes["fdh_struct_row"] and host.MATCH_ROW.itemsize == sizes["fdh_match_row"]
    assert host.RES_MATCH.itemsize == sizes["fdh_residue_match"]
    assert capi.HIT_DTYPE.itemsize == sizes["fd_struct_hit"] and capi.EDGE_DTYPE.itemsize == sizes["fd_cand_edge"]
    assert capi.PAIR_DTYPE.itemsize == sizes["fd_cand_pair"]


@pytest.mark.gpu
def test_abi_smoke_runs_config1(tmp_path, so):
    """the C program end to end on the GPU: config 1 (index of the five serine peptidases, query 4CHA B57,B102,C195)
    through fd_build_index / fd_index_attach / fd_store_attach / fd_store_build_pair_table / fdh_search reproduces the
    README rows (README.md:218-224, 237-241)."""
    import subprocess
    import fixtures as F
    exe = _build_abi_smoke(tmp_path, so)
    atoms = F.config1_atoms()
    names = F.serine_names() + ["query/4CHA.pdb"]
    soas = [O.Structure.from_atoms(atoms[n]).compact().soa() for n in names]
    ro = np.zeros(len(soas) + 1, np.uint64)
    ro[1:] = np.cumsum([len(d["aa"]) for d in soas])
    canon = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE", "PRO", "SER",
             "THR", "TRP", "TYR", "VAL"]
    aas = []
    for d in soas:
        aa = d["aa"].copy()
        for i in range(len(aa)):
            if aa[i] != 255 and bytes(d["res_name"][i]).decode() != canon[aa[i]]:
                aa[i] = 128 + aa[i]
        aas.append(aa)
    q = b"B57,B102,C195"
    path = str(tmp_path / "config1.bin")
    with open(path, "wb") as f:
        f.write(np.uint64(len(soas)).tobytes())
        f.write(ro.tobytes())
        for k in ("n_xyz", "ca_xyz", "cb_xyz"):
            f.write(np.concatenate([d[k].astype(np.float32).reshape(-1) for d in soas]).tobytes())
        f.write(np.concatenate(aas).astype(np.uint8).tobytes())
        for k, dt in (("cb_valid", np.uint8), ("chain", np.uint8), ("serial", np.uint64)):
            f.write(np.concatenate([d[k].astype(dt) for d in soas]).tobytes())
        f.write(np.uint64(len(soas) - 1).tobytes())
        f.write(np.uint64(len(q)).tobytes())
        f.write(q)
    r = subprocess.run([exe, path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    out = r.stdout.splitlines()
    assert "index %d hashes %d posting bytes" % (F.CONFIG1_NUM_HASHES, F.CONFIG1_VALUE_BYTES) in out
    srows = {int(x.split()[1]): x.split()[2:] for x in out if x.startswith("S ")}
    want = {k: F.README_STRUCT_ROWS[os.path.basename(n)] for k, n in enumerate(F.serine_names())}
    assert {k: (int(v[0]), int(v[1]), int(v[2]), v[3]) for k, v in srows.items()} == \
           {k: (w[1], w[2], w[3], "%.4f" % w[0]) for k, w in want.items()}
    mrows = sorted((os.path.basename(F.serine_names()[int(x.split()[1])]), int(x.split()[2]), x.split()[3], x.split()[4])
                   for x in out if x.startswith("M "))
    assert mrows == sorted((t, n, "%.4f" % i, "%.4f" % rm) for t, n, i, rm, _ in F.README_MATCH_ROWS_DEFAULT)


def test_rust_bindings_are_current():
    """bindings/folddisco_b200_sys.rs is generated from the headers (tools/gen_rust_ffi.py): the committed file must
    be what the generator produces now, and must name every entry point of both headers exactly once."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stderr
    hdr = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("folddisco_b200.h", "folddisco_b200_host.h"))
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(fdh?_[a-z0-9_]+)\s*\(", hdr)))
    rs = open(os.path.join(ROOT, "bindings", "folddisco_b200_sys.rs")).read()
    got = re.findall(r"pub fn (\w+)\(", rs)
    assert sorted(got) == names
    assert "pub edge_group: *const u16" in rs  # the field round 1's hand-written block had lost
