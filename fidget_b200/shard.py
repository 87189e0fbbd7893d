"""Multi-GPU partitioning of the renderers (one process per GPU).

The hot path shards without any data-path collective: a rank renders a band of
root-tile rows (2D) or a Z slab (3D) with the replicated root tape, and ONE
collective (an all-gather of the finished bands / slab images) follows.  The
reference has no distributed layer; its analogue is rayon over root tiles
(fidget-raster/src/lib.rs:152-165)."""
from __future__ import annotations


def band_rows(rank: int, world: int, height: int, root_tile: int = 128):
    """Rows of root tiles [begin, end) rendered by `rank`; bands are equal-sized so that the
    all-gather needs no padding (requires the root-row count to be divisible by `world`)."""
    n_rows = (height + root_tile - 1) // root_tile
    if n_rows % world:
        raise ValueError(f"{n_rows} root-tile rows do not split evenly over {world} ranks")
    per = n_rows // world
    return rank * per, (rank + 1) * per


def band_pixels(rows, width: int, height: int, root_tile: int = 128):
    """(first pixel row, last pixel row exclusive) of a band, clipped to the image."""
    return rows[0] * root_tile, min(rows[1] * root_tile, height)


def z_slab(rank: int, world: int, depth: int, root_tile: int = 128):
    """Voxel range [z_begin, z_end) of `rank`'s slab; slabs are whole root-tile layers, rank 0 at
    the back (z = 0), the last rank in front."""
    n_layers = (depth + root_tile - 1) // root_tile
    if n_layers % world:
        raise ValueError(f"{n_layers} root-tile layers do not split evenly over {world} ranks")
    per = n_layers // world
    return rank * per * root_tile, (rank + 1) * per * root_tile
