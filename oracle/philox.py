"""Philox4x32-10 and the Monte-Carlo shot sampler, restated in NumPy.  TEST INFRASTRUCTURE ONLY.

Counter-based generator of Salmon et al. (SC'11, "Parallel random numbers: as easy as 1, 2, 3"),
the published algorithm: 10 rounds of  (c0,c1,c2,c3) <- (hi(M1*c2)^c1^k0, lo(M1*c2),
hi(M0*c0)^c3^k1, lo(M0*c0)),  key bumped by the Weyl constants each round.  Known-answer vectors
from the Random123 distribution (kat_vectors) are checked in tests/test_oracle_golden.py.

Shot (point i, shot s) of seed S uses key = (S lo, S hi) and counters (i, s, stream, 0):
stream 0 -> two Doppler normals, stream 1 -> the position normal.  There is no reference code
for shots: this defines them.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k = np.array(key, dtype=np.uint32, copy=True)
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = k[..., 0].copy(), k[..., 1].copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1)


def normal2(seed: int, point, shot, stream: int):
    """Two standard normals per (point, shot): Box-Muller on two 53-bit uniforms in (0,1)."""
    point = np.asarray(point, dtype=np.uint32)
    shot = np.asarray(shot, dtype=np.uint32)
    point, shot = np.broadcast_arrays(point, shot)
    ctr = np.stack([point, shot, np.full_like(point, stream), np.zeros_like(point)], axis=-1)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32),
                          ctr.shape[:-1] + (2,))
    r = philox4x32_10(ctr, key).astype(np.uint64)
    a = (r[..., 1] << np.uint64(32)) | r[..., 0]
    b = (r[..., 3] << np.uint64(32)) | r[..., 2]
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0
    u2 = ((b >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0
    rad = np.sqrt(-2.0 * np.log(u1))
    return rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)


def shot_parameters(seed: int, point: int, shot: int, V: float, sig_det: float, sig_relr: float):
    """-> (off_a, off_b, V_shot) for one shot."""
    da, db = normal2(seed, point, shot, 0)
    er, _ = normal2(seed, point, shot, 1)
    q = 1.0 + sig_relr * float(er)
    return sig_det * float(da), sig_det * float(db), V / q ** 6
