"""Multi-GPU partitioning of the renderers (one process per GPU).

The hot path shards without any data-path collective: a rank renders a band of
root-tile rows (2D, or 3D at full depth) or a Z slab (3D) with the replicated root tape, and ONE
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


def _on_torch_stream(shape):
    """The renderers enqueue on the fc_ctx's stream, the collective on torch's current stream: bind
    the context to torch's current stream for the duration of the call so that both are ordered
    (a no-op when the caller has already done so with ``CudaContext.set_stream``)."""
    import torch
    return shape.cuda.on_stream(torch.cuda.current_stream().cuda_stream)


def render2d_bands(shape, cfg, image, gathered, group=None):
    """One sharded 2D frame: this rank renders its band of root-tile rows into `image`
    (a CUDA tensor [H, W] float32), then ONE all-gather assembles all bands into `gathered`.
    Enqueues on torch's current stream (the context is bound to it for the call); returns `gathered`."""
    import torch.distributed as dist
    from dataclasses import replace
    from .shape import render2d
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t0 = (cfg.tile_sizes[0] if cfg.tile_sizes else 128)
    rows = band_rows(rank, world, cfg.height, t0)
    y0, y1 = band_pixels(rows, cfg.width, cfg.height, t0)
    with _on_torch_stream(shape):
        render2d(shape, replace(cfg, root_rows=rows), out=image, asynchronous=True)
        dist.all_gather_into_tensor(gathered, image[y0:y1], group=group)
    return gathered


def render3d_ybands(shape, cfg, image, gathered, group=None):
    """One sharded 3D render in Y bands: this rank renders ALL depths of its band of root-tile rows into
    `image` ([H, W, 4] float32 CUDA tensor viewed as GeometryPixel), then ONE all-gather of the disjoint
    bands assembles the frame in `gathered` -- no merge pass, 1/world of the slab traffic, and the
    surface (where the work is) is spread over the ranks instead of sitting in one Z slab."""
    import torch.distributed as dist
    from dataclasses import replace
    from .shape import render3d
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t0 = (cfg.tile_sizes[0] if cfg.tile_sizes else 128)
    rows = band_rows(rank, world, cfg.height, t0)
    y0, y1 = band_pixels(rows, cfg.width, cfg.height, t0)
    with _on_torch_stream(shape):
        render3d(shape, replace(cfg, root_rows=rows), out=image, asynchronous=True)
        dist.all_gather_into_tensor(gathered, image[y0:y1], group=group)
    return gathered


def render3d_zslabs(shape, cfg, slab, gathered, out, group=None):
    """One sharded 3D render (north star: Z slabs + a single all-gather): this rank renders its
    slab into `slab` ([H, W, 4] float32 CUDA tensor viewed as GeometryPixel, no final clamp), all
    slabs are gathered into `gathered` ([world, H, W, 4]) and merged per pixel into `out`."""
    import ctypes as C
    import torch.distributed as dist
    from dataclasses import replace
    from . import _lib
    from .shape import render3d, _ck
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t0 = (cfg.tile_sizes[0] if cfg.tile_sizes else 128)
    zr = z_slab(rank, world, cfg.depth, t0)
    with _on_torch_stream(shape):
        render3d(shape, replace(cfg, z_range=zr, clamp=False), out=slab, asynchronous=True)
        dist.all_gather_into_tensor(gathered, slab, group=group)
        ptrs = (C.c_void_p * world)(*[gathered[r].data_ptr() for r in range(world)])
        _ck(_lib.load().fc_merge_slabs(shape.cuda._h, ptrs, world, cfg.width, cfg.height, cfg.depth,
                                       C.c_void_p(out.data_ptr())))
    return out
