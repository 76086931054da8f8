"""Host statement of how destination rows are dealt to GPUs; the device does the same in
`assign_internal_kernel` (csrc/pio_als.cu).  Used by the multi-process tests and by tooling that needs to know
which rank owns a row; the training path itself never calls this.

Rows are ranked by rating count (descending, ties by id) and dealt in snake order, so every rank owns exactly
R = ceil(n / W) internal rows (equal shards -> one in-place ncclAllGather per half-iteration) and a near-equal
share of the ratings; a rank's rows stay degree-descending (the solve kernel batches neighbours).
Replaces MLlib's hash partitioning of ids into blocks (ALSPartitioner).
"""
from __future__ import annotations

import numpy as np


def rows_per_rank(n_rows: int, world: int) -> int:
    return (n_rows + world - 1) // world


def assign_internal(degree: np.ndarray, world: int):
    """-> (perm[row] = internal id, inv[internal] = row or -1).  internal id = rank * R + local index."""
    n = degree.shape[0]
    R = rows_per_rank(n, world)
    order = np.argsort(-degree.astype(np.int64), kind="stable")          # degree desc, ties by row id
    p = np.arange(n)
    blk, pos = p // world, p % world
    rk = np.where(blk % 2 == 1, world - 1 - pos, pos)
    internal = rk * R + blk
    perm = np.empty(n, np.int64)
    perm[order] = internal
    inv = np.full(world * R, -1, np.int64)
    inv[internal] = order
    return perm, inv


def owner_rank(perm: np.ndarray, n_rows: int, world: int) -> np.ndarray:
    return perm // rows_per_rank(n_rows, world)
