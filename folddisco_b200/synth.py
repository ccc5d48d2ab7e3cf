"""Seeded synthetic structure generator for the benchmark configurations.

Every synthetic structure is a random contiguous crop of one of the 27 PDB files shipped with the reference
(tests/golden/templates.npz, produced by tests/golden/make_fixtures.py), with i.i.d. N(0, 0.35 A) jitter on
every N/CA/CB, 15 % of the residues mutated to a background-frequency draw, and a random rigid transform
(the hash is invariant to it; it exercises the numerics).  Crop length ~ LogNormal(ln 300, 0.6) clipped to
[40, 2000].  This keeps realistic pair-distance / angle statistics (hence realistic posting-length skew) and
plants true near-hits for every shipped motif.  Generator: numpy PCG64 seeded with 0xF01DD15C0 + config number.

Workload generation only -- not on the product path and not part of the oracle.
"""
import os

import numpy as np

_TEMPLATES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                          "templates.npz")
SEED_BASE = 0xF01DD15C0
# UniProt background amino-acid frequencies in the reference's index order (A R N D C Q E G H I L K M F P S T W Y V)
BACKGROUND = np.array([8.25, 5.53, 4.06, 5.45, 1.37, 3.93, 6.75, 7.07, 2.27, 5.96, 9.66, 5.84, 2.42, 3.86, 4.70,
                       6.56, 5.34, 1.08, 2.92, 6.87])
BACKGROUND = BACKGROUND / BACKGROUND.sum()

_tpl = None


def templates():
    global _tpl
    if _tpl is None:
        z = np.load(_TEMPLATES)
        _tpl = {k: z[k] for k in z.files}
    return _tpl


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - z * w)
    R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w)
    R[:, 2, 1] = 2 * (y * z + x * w)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def generate(n_structs, seed, mean_len=300.0, sigma=0.6, min_len=40, max_len=2000, jitter=0.35, mutate=0.15,
             template_ids=None):
    """-> dict(row_offsets u64[S+1], n_xyz, ca_xyz, cb_xyz f32[R,3], aa u8[R]); vectorised, deterministic."""
    t = templates()
    rng = np.random.Generator(np.random.PCG64(seed))
    toff = t["offsets"]
    tlen = np.diff(toff)
    which = rng.integers(0, len(tlen), n_structs)
    if template_ids is not None:  # restrict to a few templates: near-duplicate structures, long posting lists
        which = np.asarray(template_ids, np.int64)[which % len(template_ids)]
    want = np.clip(np.exp(rng.normal(np.log(mean_len), sigma, n_structs)), min_len, max_len).astype(np.int64)
    length = np.minimum(want, tlen[which])
    start = (rng.random(n_structs) * (tlen[which] - length + 1)).astype(np.int64)
    ro = np.zeros(n_structs + 1, np.uint64)
    ro[1:] = np.cumsum(length)
    R = int(ro[-1])
    sid = np.repeat(np.arange(n_structs), length)
    within = np.arange(R) - np.repeat(ro[:-1].astype(np.int64), length)
    src = toff[which][sid] + start[sid] + within
    rot = _random_rotations(rng, n_structs)
    trans = rng.uniform(-50, 50, (n_structs, 3))
    out = {"row_offsets": ro}
    center = t["ca_xyz"][toff[which] + start + length // 2].astype(np.float64)
    for k in ("n_xyz", "ca_xyz", "cb_xyz"):
        p = t[k][src].astype(np.float64) + rng.normal(0.0, jitter, (R, 3))
        p = p - center[sid]
        p = np.einsum("rij,rj->ri", rot[sid], p) + trans[sid]
        out[k] = np.ascontiguousarray(p, np.float32)
    aa = t["aa"][src].copy()
    mut = rng.random(R) < mutate
    aa[mut] = rng.choice(20, int(mut.sum()), p=BACKGROUND).astype(np.uint8)
    out["aa"] = aa
    return out


def split(batch):
    """list of per-structure dicts (views)"""
    ro = batch["row_offsets"].astype(np.int64)
    return [{k: batch[k][ro[s]:ro[s + 1]] for k in ("n_xyz", "ca_xyz", "cb_xyz", "aa")} for s in range(len(ro) - 1)]
