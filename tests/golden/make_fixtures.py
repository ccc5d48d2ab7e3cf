"""Regenerates the committed fixtures under tests/golden/ from the reference's shipped data.

Run HERE (the build container), where /root/reference exists:
    python tests/golden/make_fixtures.py
The GPU box has no /root/reference, so -m gpu tests, smoke() and bench.py read only these files.

  atoms_config1.npz  parsed ATOM records (x,y,z,atom_name,chain,res_name,res_serial,b_factor) of the five
                     data/serine_peptidases/*.pdb (config 1 index) and the six query/*.pdb files, as read by the
                     oracle's restatement of src/structure/io/{parser,pdb}.rs.
  templates.npz      CompactStructure SoA (N/CA/CB, aa, cb_valid) of the 27 shipped PDB files; data/long/* cropped
                     to their first 2000 residues.  Seeds for the synthetic-structure generator.
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

REF = "/root/reference"


def main():
    atoms = {}
    names = []
    for p in sorted(glob.glob(REF + "/data/serine_peptidases/*.pdb")) + sorted(glob.glob(REF + "/query/*.pdb")):
        key = os.path.relpath(p, REF)
        a = O.Structure.read_pdb(p).atoms()
        names.append(key)
        for k, v in a.items():
            atoms["%s|%s" % (key, k)] = v
    atoms["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "atoms_config1.npz"), **atoms)

    tfiles = (sorted(glob.glob(REF + "/data/serine_peptidases/*.pdb")) + sorted(glob.glob(REF + "/query/*.pdb"))
              + sorted(glob.glob(REF + "/data/homeobox/*.pdb")) + sorted(glob.glob(REF + "/data/zinc/*.pdb"))
              + [REF + "/data/AF-P17538-F1-model_v4.pdb"] + sorted(glob.glob(REF + "/data/long/*.pdb")))
    assert len(tfiles) == 27, len(tfiles)
    cat = {k: [] for k in ("n_xyz", "ca_xyz", "cb_xyz", "aa", "cb_valid")}
    offs = [0]
    for p in tfiles:
        d = O.Structure.read_pdb(p).compact().soa()
        crop = 2000 if "/long/" in p else len(d["aa"])
        ok = (d["aa"][:crop] != 255) & (d["cb_valid"][:crop] == 1)
        for k in cat:
            cat[k].append(d[k][:crop][ok])
        offs.append(offs[-1] + int(ok.sum()))
    out = {k: np.concatenate(v) for k, v in cat.items()}
    out["offsets"] = np.array(offs, np.int64)
    out["names"] = np.array([os.path.relpath(p, REF) for p in tfiles])
    np.savez_compressed(os.path.join(HERE, "templates.npz"), **out)
    print("atoms:", len(names), "files;", "templates:", offs[-1], "residues")


if __name__ == "__main__":
    main()


def long_structure():
    """long_6FF7.npz: the UNCROPPED CompactStructure of data/long/6FF7.pdb (16 769 residues, the largest shipped
    input; SURVEY section 7 minimum slice) with the oracle's answer for it: number of ordered residue pairs that get a
    hash, number of distinct hashes and the sha256 of the sorted unique u32 hash list (src/controller/feature.rs:198-231
    + sort_unstable + dedup of src/controller/mod.rs:343-344)."""
    import hashlib
    c = O.Structure.read_pdb(REF + "/data/long/6FF7.pdb").compact()
    d = c.soa()
    raw = c.hashes()
    uniq = np.unique(raw)
    np.savez_compressed(os.path.join(HERE, "long_6FF7.npz"), n_xyz=d["n_xyz"], ca_xyz=d["ca_xyz"], cb_xyz=d["cb_xyz"],
                        aa=d["aa"], cb_valid=d["cb_valid"], res_name=d["res_name"], chain=d["chain"], serial=d["serial"],
                        n_pairs=np.int64(len(raw)), n_unique=np.int64(len(uniq)),
                        sha256=np.array(hashlib.sha256(uniq.astype("<u4").tobytes()).hexdigest()))
    print("long_6FF7:", len(d["aa"]), "residues,", len(raw), "hashed pairs,", len(uniq), "distinct hashes")


if __name__ == "__main__" and "--long" in sys.argv:
    long_structure()


def foldcomp_database():
    """foldcomp/example_db{,.index,.lookup}: the reference's example Foldcomp database (data/foldcomp/example_db, 24 SCOP
    domains, 52 KB), byte for byte -- input data of the GPU test of `index -p FOLDCOMP_DB`."""
    import shutil
    os.makedirs(os.path.join(HERE, "foldcomp"), exist_ok=True)
    for ext in ("", ".index", ".lookup"):
        shutil.copyfile(REF + "/data/foldcomp/example_db" + ext, os.path.join(HERE, "foldcomp", "example_db" + ext))
