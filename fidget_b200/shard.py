"""Multi-GPU partitioning of the renderers (one process per GPU).

The hot path shards without any data-path collective while rendering: a rank renders its share of the
root tiles with the replicated root tape, and ONE collective (an all-gather of the finished pieces)
follows.  Three partitions:

* **interleaved root tiles** (`render2d_tiles`, `render3d_tiles`; what `bench.py --gpus N` measures): rank r
  owns the root-tile columns (tx, ty) with tile_owner(tx, ty, N) == r (a spatial hash), at full depth.  Surface-like work is spread
  evenly whatever the model looks like, every column keeps its front-to-back culling, each rank
  contributes 1/N of the image to the gather and nothing has to be merged.  The reference's analogue is
  rayon handing root tiles to worker threads (fidget-raster/src/lib.rs:152-165).
* bands of root-tile rows (`render2d_bands`, `render3d_ybands`): contiguous pieces, no pack/unpack, but
  the busiest band sets the pace.
* Z slabs (`render3d_zslabs`, the north star's wording): N full-size images + a per-pixel merge
  (voxel.rs:535-546 clamp applied after the merge); N times the gather traffic and no culling across slabs."""
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


def tile_owner(tx: int, ty: int, world: int) -> int:
    """Rank that renders root tile (tx, ty): the spatial hash of dev_ops.cuh::tile_owner (a pseudo-random spread
    balances heavy-tailed per-tile cost better than a regular pattern)."""
    return (((tx * 73856093) & 0xFFFFFFFF) ^ ((ty * 19349663) & 0xFFFFFFFF)) % world


def owned_tiles(rank: int, world: int, width: int, height: int, root_tile: int = 128):
    """Root tiles (tx, ty) of `rank`, row-major: the order of its chunk in the all-gather."""
    rx, ry = (width + root_tile - 1) // root_tile, (height + root_tile - 1) // root_tile
    return [(tx, ty) for ty in range(ry) for tx in range(rx) if tile_owner(tx, ty, world) == rank]


def tiles_per_rank(world: int, width: int, height: int, root_tile: int = 128) -> int:
    """Tiles in every rank's all-gather chunk (the largest ownership count; smaller owners pad)."""
    return max(len(owned_tiles(r, world, width, height, root_tile)) for r in range(world))


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


def _render_tiles(render, px_floats, shape, cfg, image, chunk, gathered, group):
    import ctypes as C
    import torch.distributed as dist
    from dataclasses import replace
    from . import _lib
    from .shape import _ck
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t0 = (cfg.tile_sizes[0] if cfg.tile_sizes else 128)
    lib, h = _lib.load(), shape.cuda._h
    with _on_torch_stream(shape):
        render(shape, replace(cfg, interleave=(world, rank)), out=image, asynchronous=True)
        _ck(lib.fc_tiles_pack(h, C.c_void_p(image.data_ptr()), cfg.width, cfg.height, 4 * px_floats, t0, world, rank,
                              C.c_void_p(chunk.data_ptr())))
        dist.all_gather_into_tensor(gathered, chunk, group=group)
        _ck(lib.fc_tiles_unpack(h, C.c_void_p(gathered.data_ptr()), cfg.width, cfg.height, 4 * px_floats, t0, world,
                                C.c_void_p(image.data_ptr())))
    return image


def tile_buffers(world: int, width: int, height: int, px_floats: int, device, root_tile: int = 128):
    """(chunk, gathered) CUDA tensors for `render2d_tiles` (px_floats = 1) / `render3d_tiles` (px_floats = 4)."""
    import torch
    per = tiles_per_rank(world, width, height, root_tile)
    shape = (per, root_tile, root_tile, px_floats)
    return (torch.zeros(shape, dtype=torch.float32, device=device),
            torch.zeros((world * per,) + shape[1:], dtype=torch.float32, device=device))


def render3d_tiles(shape, cfg, image, chunk, gathered, group=None):
    """One sharded 3D render over interleaved root-tile columns: this rank renders its columns (full depth,
    final clamp applied) into `image` ([H, W, 4] float32 CUDA tensor viewed as GeometryPixel), packs them
    into `chunk`, ONE all-gather fills `gathered`, and the unpack writes every pixel of `image` -- the
    complete frame on every rank, byte-identical to a single-GPU render."""
    from .shape import render3d
    return _render_tiles(render3d, 4, shape, cfg, image, chunk, gathered, group)


def render2d_tiles(shape, cfg, image, chunk, gathered, group=None):
    """The 2D counterpart of `render3d_tiles` (`image`: [H, W] float32 CUDA tensor)."""
    from .shape import render2d
    return _render_tiles(render2d, 1, shape, cfg, image, chunk, gathered, group)
