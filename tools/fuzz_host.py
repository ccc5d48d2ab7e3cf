"""Mutation fuzzing of the host-side readers and loaders (no GPU needed): structure files (.pdb / .cif), index files
(PREFIX, .offset, .lookup, .type), PREFIX.store, Foldcomp database tables and query strings.  Every mutated input must
be parsed or refused with an error -- a crash ends this script with a signal, a hang is caught by the caller's timeout:

    timeout 900 python tools/fuzz_host.py            # the structure-file part needs /root/reference

Found so far: the mmCIF tokenizer spun forever on a \\v or \\f between two values (fixed; regression test in
tests/test_host_cpu.py)."""
import ctypes as C
import os
import random
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def mutate(data, rng, text=False):
    d = bytearray(data)
    for _ in range(rng.choice([1, 3, 10, 50])):
        if not d:
            d = bytearray(b" ")
        op, pos = rng.random(), rng.randrange(len(d))
        if op < 0.4:
            d[pos] = rng.randrange(256)
        elif op < 0.6:
            del d[pos:pos + rng.randrange(1, 200)]
        elif op < 0.8:
            d[pos:pos] = bytes(rng.randrange(32, 127) if text else rng.randrange(256) for _ in range(rng.randrange(1, 40)))
        elif op < 0.9:
            d = d[:pos]
        else:
            d[pos:pos + 8] = rng.choice([0, 1 << 62, (1 << 64) - 1, len(d), 1 << 32]).to_bytes(8, "little")
    return bytes(d)


def fuzz_structure_files(host, tmp, rng, n=1500):
    ok = err = 0
    for src in (REF + "/data/io_test/cif/2wnb.cif", REF + "/data/serine_peptidases/4cha.pdb"):
        data, ext = open(src, "rb").read(), os.path.splitext(src)[1]
        for _ in range(n):
            p = os.path.join(tmp, "f" + ext)
            open(p, "wb").write(mutate(data, rng, text=True))
            try:
                host.read_structure_from_path(p).soa()
                ok += 1
            except host.FdError:
                err += 1
    print("structure files: parsed", ok, "refused", err)


def fuzz_index_files(host, tmp, rng, n=1200):
    import fixtures as F
    import oracle_lib as O
    from folddisco_b200 import capi
    atoms, names = F.config1_atoms(), F.serine_names()
    store, comps = host.Store(), []
    for nm in names:
        store.add(host.CompactStructure.from_atoms(atoms[nm]), nm)
        comps.append(O.Structure.from_atoms(atoms[nm]).compact())
    oix = O.Index.build(comps)
    hashes, offsets, values = oix.hashes, oix.offsets, oix.values
    b = capi._IndexBuffers(len(hashes), hashes.ctypes.data_as(C.POINTER(C.c_uint32)),
                           offsets.ctypes.data_as(C.POINTER(C.c_uint64)), len(values),
                           values.ctypes.data_as(C.POINTER(C.c_uint8)))
    p = capi.HashParams(0, 0, 20.0)
    ix = host.FolddiscoIndex(host._lib().fdh_index_from_buffers(C.byref(b), store.h, C.byref(p)))
    base = os.path.join(tmp, "base")
    ix.save(store, base, foldcomp_db="x")
    store.save(base + ".store")
    files = {e: open(base + e, "rb").read() for e in ("", ".offset", ".lookup", ".type", ".store")}
    ok = err = 0
    for _ in range(n):
        which = rng.choice(list(files))
        pre = os.path.join(tmp, "m")
        for e, d in files.items():
            open(pre + e, "wb").write(mutate(d, rng) if e == which else d)
        try:
            if which == ".store":
                s = host.Store.load(pre + ".store")
                s.lookup()
                [s.name(i) for i in range(min(len(s), 8))]
            else:
                x = host.load_folddisco_index(pre)
                x.lookup()
                bb = x.buffers()
                int(bb.hashes.sum())
                [x.name(i) for i in range(min(x.num_structs, 8))]
            ok += 1
        except Exception:
            err += 1
    print("index / store files: loaded", ok, "refused", err)


def fuzz_foldcomp_tables_and_query_strings(host, tmp, rng):
    import re
    import fixtures as F
    gold = os.path.join(ROOT, "tests", "golden", "foldcomp", "example_db")
    files = {e: open(gold + e, "rb").read() for e in ("", ".index", ".lookup")}
    ok = err = 0
    for _ in range(600):
        which = rng.choice([".index", ".lookup"])
        for e, d in files.items():
            open(os.path.join(tmp, "db" + e), "wb").write(mutate(d, rng) if e == which else d)
        try:
            db = host.FoldcompDb(os.path.join(tmp, "db"))
            db.names(), db.keys(), db.find("d1asha_")
            ok += 1
        except host.FdError:
            err += 1
    print("Foldcomp tables: opened", ok, "refused", err)
    c = host.CompactStructure.from_atoms(F.config1_atoms()["query/4CHA.pdb"])
    ok = err = 0
    for _ in range(4000):
        q = "".join(rng.choice("ABCabc0123456789,:-: XxZ;\t+") for _ in range(rng.randrange(0, 24)))
        if rng.random() < 0.3:
            q = "B57,B102,C195"[:rng.randrange(14)] + q
        if any(len(x) > 5 for x in re.findall(r"\d+", q)):
            continue  # (a range like 1-999999999 is legal and expands, as in the reference)
        try:
            host.parse_query_string(q)
            host.QueryBatch().add(c, q)
            ok += 1
        except (host.FdError, ValueError):
            err += 1
    print("query strings: accepted", ok, "refused", err)


def main():
    from folddisco_b200 import host
    rng = random.Random(5)
    tmp = tempfile.mkdtemp(prefix="fd_fuzz_")
    try:
        if os.path.isdir(REF):
            fuzz_structure_files(host, tmp, rng)
        fuzz_index_files(host, tmp, rng)
        fuzz_foldcomp_tables_and_query_strings(host, tmp, rng)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
