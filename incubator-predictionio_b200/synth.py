"""Deterministic synthetic (user, item, rating) events -- SURVEY.md section 8(d).

Stateless counter hash, so host (NumPy), CUDA (``pio_als_synth_ratings_device``)
and any other language regenerate bit-identical triplets:

    h_j  = splitmix64(seed ^ (4*e + j)),  j = 1..3,  e in [0, nnz)
    user = h_1 mod U                                 (uniform)
    item = floor(I * u^2), u = (h_2 >> 11) * 2^-53   (power-law popularity)
    explicit rating = 1 + (h_3 mod 5)
    implicit count  = 1 + min(ctz(h_3), 9)           (1 + geometric(1/2), truncated at 10)

Stands in for the event scan of the reference's DataSource
(examples/scala-parallel-recommendation/blacklist-items/src/main/scala/DataSource.scala:45-75).
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _ctz64(h: np.ndarray) -> np.ndarray:
    # count trailing zeros, capped at 9 (enough for the truncated geometric)
    out = np.full(h.shape, 9, np.int32)
    for b in range(8, -1, -1):
        out[(h >> np.uint64(b)) & np.uint64(1) == 1] = b
    return out


def synth_ratings(n_users: int, n_items: int, nnz: int, seed: int = 3, implicit: bool = False,
                  start: int = 0, chunk: int = 1 << 24):
    """Return (user int32[nnz], item int32[nnz], rating float32[nnz]) for events start..start+nnz."""
    user = np.empty(nnz, np.int32)
    item = np.empty(nnz, np.int32)
    rating = np.empty(nnz, np.float32)
    s = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
    for lo in range(0, nnz, chunk):
        hi = min(nnz, lo + chunk)
        e = np.arange(start + lo, start + hi, dtype=np.uint64)
        with np.errstate(over="ignore"):
            base = e * np.uint64(4)
        h1 = splitmix64(s ^ (base + np.uint64(1)))
        h2 = splitmix64(s ^ (base + np.uint64(2)))
        h3 = splitmix64(s ^ (base + np.uint64(3)))
        user[lo:hi] = (h1 % np.uint64(n_users)).astype(np.int32)
        u = (h2 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        it = np.floor((float(n_items) * u) * u).astype(np.int64)
        np.minimum(it, n_items - 1, out=it)
        item[lo:hi] = it.astype(np.int32)
        if implicit:
            rating[lo:hi] = (1 + _ctz64(h3)).astype(np.float32)
        else:
            rating[lo:hi] = (1 + (h3 % np.uint64(5)).astype(np.int32)).astype(np.float32)
    return user, item, rating


def synth_init_factors(n_rows: int, rank: int, seed: int, side: int) -> np.ndarray:
    """Unit-L2-norm Gaussian rows from the counter hash keyed by (side,row,col).

    Plays the role of MLlib's `initialize` (N(0,1)^k scaled to unit norm per row,
    SURVEY 8(c)-3); MLlib's own block-seeded XORShift stream is machine-dependent,
    so initial factors are an explicit input of the C ABI instead.
    """
    r = np.arange(n_rows, dtype=np.uint64)[:, None]
    c = np.arange(rank, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        ctr = (r * np.uint64(rank) + c) * np.uint64(2) + (np.uint64(side) << np.uint64(62))
    s = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) ^ np.uint64(0xA5A5A5A55A5A5A5A)
    h1 = splitmix64(s ^ ctr)
    h2 = splitmix64(s ^ (ctr + np.uint64(1)))
    u1 = ((h1 >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)  # (0,1]
    u2 = (h2 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    g = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    g = g.astype(np.float32)
    nrm = np.sqrt((g.astype(np.float64) ** 2).sum(1)).astype(np.float32)
    nrm[nrm == 0] = 1.0
    return (g / nrm[:, None]).astype(np.float32)
