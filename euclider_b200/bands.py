"""Row-band partition of a frame across ranks (one process per GPU).

Pixels are independent (src/universe/mod.rs:316-348), so the only multi-GPU step is a gather of
finished rows to rank 0.  Bands are interleaved (band b -> rank b % world) because the expensive
pixels (glass, mirrors) cluster in the middle rows.  Mirrors eucl_band_rows_for_rank /
frame_row_of_local in the C library.
"""
from __future__ import annotations

from typing import List


def local_rows(height: int, band_rows: int, rank: int, world: int) -> List[int]:
    """Frame rows owned by `rank`, in local-row order."""
    if world <= 1:
        return list(range(height))
    band_rows = band_rows or height
    n_bands = (height + band_rows - 1) // band_rows
    rows: List[int] = []
    for b in range(rank, n_bands, world):
        rows.extend(range(b * band_rows, min((b + 1) * band_rows, height)))
    return rows


def gather_frame(parts, height: int, band_rows: int, world: int):
    """Reassembles the full frame from per-rank compact row blocks (numpy or torch tensors with
    rows on axis 0).  parts[r][k] is local row k of rank r."""
    first = parts[0]
    frame = first.new_empty((height,) + tuple(first.shape[1:])) if hasattr(first, "new_empty") else __import__("numpy").empty(
        (height,) + tuple(first.shape[1:]), dtype=first.dtype)
    for r in range(world):
        rows = local_rows(height, band_rows, r, world)
        for k, y in enumerate(rows):
            frame[y] = parts[r][k]
    return frame
