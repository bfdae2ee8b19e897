"""Pixel partition across ranks: host-side index math (mirrors csrc/render.cu partition_* and the kernels' tile walk).

The image is cut into 8x4-pixel tiles, numbered row-major; tile L belongs to rank ``L % n_ranks`` and is that rank's
tile number ``L // n_ranks``.  Inside a tile, pixel (x, y) sits in slot ``(y % 4) * 8 + (x % 8)`` (one warp lane per
pixel).  A rank's compact buffer holds ``capacity(w, h, n)`` pixels: its tiles back to back, padded with zeros
(image-edge lanes and, for the last ranks, one absent tile).  The RNG stream of a pixel depends only on
(pixel index, frame index) (pathTracerPrograms.cu:721), so any pixel -> rank map renders bit-identical pixels; the
interleave spreads expensive image regions over all ranks (cf. sutil/WorkDistribution.h:62-80).
"""
import numpy as np

TILE_W, TILE_H = 8, 4


def tile_grid(width, height):
    return (width + TILE_W - 1) // TILE_W, (height + TILE_H - 1) // TILE_H


def owned_tiles(width, height, rank, n_ranks):
    tx, ty = tile_grid(width, height)
    tiles = tx * ty
    return (tiles - rank + n_ranks - 1) // n_ranks if rank < tiles else 0


def capacity(width, height, n_ranks):
    """Pixels per rank in the compact exchange buffers (same for every rank)."""
    tx, ty = tile_grid(width, height)
    return ((tx * ty + n_ranks - 1) // n_ranks) * 32


def compact_pixel_index(width, height, rank, n_ranks):
    """int64[capacity]: image pixel index (y*w+x) stored in each compact slot of `rank`, -1 for padding."""
    tx, _ = tile_grid(width, height)
    cap = capacity(width, height, n_ranks)
    slot = np.arange(cap, dtype=np.int64)
    k, lane = slot >> 5, slot & 31
    tile = rank + k * n_ranks
    x = (tile % tx) * TILE_W + (lane & 7)
    y = (tile // tx) * TILE_H + (lane >> 3)
    ok = (k < owned_tiles(width, height, rank, n_ranks)) & (x < width) & (y < height)
    return np.where(ok, y * width + x, -1)


def owner_map(width, height, n_ranks):
    """int32[h, w]: owning rank of every pixel."""
    tx, _ = tile_grid(width, height)
    y, x = np.mgrid[0:height, 0:width]
    return (((y // TILE_H) * tx + (x // TILE_W)) % n_ranks).astype(np.int32)


def pack_numpy(full, width, height, rank, n_ranks):
    """CPU mirror of pt_partition_pack for a (h, w, C) array -> (capacity, C)."""
    idx = compact_pixel_index(width, height, rank, n_ranks)
    flat = full.reshape(width * height, -1)
    out = np.zeros((idx.shape[0], flat.shape[1]), full.dtype)
    m = idx >= 0
    out[m] = flat[idx[m]]
    return out


def unpack_numpy(compact_all, width, height, n_ranks):
    """CPU mirror of pt_partition_unpack: (n_ranks, capacity, C) -> (h, w, C)."""
    c = compact_all.shape[-1]
    out = np.zeros((width * height, c), compact_all.dtype)
    for r in range(n_ranks):
        idx = compact_pixel_index(width, height, r, n_ranks)
        m = idx >= 0
        out[idx[m]] = compact_all[r][m]
    return out.reshape(height, width, c)
