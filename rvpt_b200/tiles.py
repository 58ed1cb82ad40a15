"""Host-side description of the tile layout the kernels write (include/rvpt_abi.h,
"Device-side tile buffers"): which pixels a rank owns and where they sit in its
tile buffer. Pure index arithmetic (numpy) — used by the multi-GPU gather
plumbing and its CPU tests; it computes no radiance.

Layout: 16x16 tiles (the reference's workgroup footprint, compute_pass.comp:27),
tile_id = ty * tiles_x + tx, rank r owns tile_id % nranks == r as local tile
j = tile_id // nranks. Inside a tile, pixel (px, py) sits at
    q = warp * 32 + lane,  warp = (py // 4) * 2 + (px // 8),  lane = (py % 4) * 8 + (px % 8)
so each warp's 8x4 pixel block is 32 consecutive elements.
"""
from __future__ import annotations

import numpy as np

TILE = 16
TILE_PIXELS = 256


def tile_grid(width: int, height: int) -> tuple[int, int]:
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


def local_tile_counts(width: int, height: int, rank: int, nranks: int) -> tuple[int, int]:
    """(tiles owned by `rank`, padded count equal on all ranks)."""
    tx, ty = tile_grid(width, height)
    n = tx * ty
    owned = (n - rank + nranks - 1) // nranks if rank < n else 0
    return owned, (n + nranks - 1) // nranks


def in_tile_offsets() -> tuple[np.ndarray, np.ndarray]:
    """px, py for q = 0..255."""
    q = np.arange(TILE_PIXELS)
    warp, lane = q >> 5, q & 31
    px = ((warp & 1) << 3) + (lane & 7)
    py = ((warp >> 1) << 2) + (lane >> 3)
    return px, py


def slot_to_raster(width: int, height: int, rank: int, nranks: int) -> np.ndarray:
    """int64 [n_local_padded * 256]: raster index y*W+x of every slot of the
    rank's tile buffer, -1 for padding (tiles past the end, pixels off-image)."""
    tiles_x, tiles_y = tile_grid(width, height)
    n_tiles = tiles_x * tiles_y
    _, padded = local_tile_counts(width, height, rank, nranks)
    j = np.arange(padded)
    g = j * nranks + rank
    tx, ty = g % tiles_x, g // tiles_x
    px, py = in_tile_offsets()
    x = tx[:, None] * TILE + px[None, :]
    y = ty[:, None] * TILE + py[None, :]
    ok = (g[:, None] < n_tiles) & (x < width) & (y < height)
    idx = np.where(ok, y * width + x, -1)
    return idx.reshape(-1).astype(np.int64)


def tiles_from_raster(raster: np.ndarray, rank: int, nranks: int) -> np.ndarray:
    """Extracts a rank's tile buffer ([n_local_padded*256, C]) from an HxWxC
    raster image (padding slots are 0) — what the rank's kernels would write."""
    h, w = raster.shape[:2]
    idx = slot_to_raster(w, h, rank, nranks)
    flat = raster.reshape(h * w, -1)
    out = np.zeros((len(idx), flat.shape[1]), raster.dtype)
    ok = idx >= 0
    out[ok] = flat[idx[ok]]
    return out


def raster_from_gathered(gathered: np.ndarray, width: int, height: int, nranks: int) -> np.ndarray:
    """Inverse: [nranks, n_local_padded*256, C] (the all-gather result) ->
    HxWxC raster. The numpy twin of rvpt_b200_untile()."""
    c = gathered.shape[-1]
    out = np.zeros((height * width, c), gathered.dtype)
    g = gathered.reshape(nranks, -1, c)
    for r in range(nranks):
        idx = slot_to_raster(width, height, r, nranks)
        ok = idx >= 0
        out[idx[ok]] = g[r][ok]
    return out.reshape(height, width, c)
