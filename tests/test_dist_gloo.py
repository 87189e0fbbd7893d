"""Host-side logic of the N>1 path on CPU (world_size 2 and 4, gloo): every rank renders ONLY the root tiles
it owns under the tile interleave of fidget_b200.shard (the CPU oracle stands in for the GPU renderer), packs
them into its all-gather chunk, one all-gather runs, and the unpacked frame must equal the single-process
image byte for byte.  The chunk layout used here ([tile][row][pixel], owners in rank order, row-major tiles)
is the one fc_tiles_pack / fc_tiles_unpack implement on the device."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T0 = 128


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(image, tiles, per):
    """image [H, W, C] -> chunk [per, T0, T0, C] (ragged edge tiles are zero-padded)"""
    chunk = np.zeros((per, T0, T0) + image.shape[2:], dtype=image.dtype)
    for k, (tx, ty) in enumerate(tiles):
        piece = image[ty * T0:(ty + 1) * T0, tx * T0:(tx + 1) * T0]
        chunk[k, :piece.shape[0], :piece.shape[1]] = piece
    return chunk


def _unpack(gathered, world, width, height, like):
    from fidget_b200.shard import owned_tiles
    out = np.zeros_like(like)
    for r in range(world):
        for k, (tx, ty) in enumerate(owned_tiles(r, world, width, height, T0)):
            h, w = min(T0, height - ty * T0), min(T0, width - tx * T0)
            out[ty * T0:ty * T0 + h, tx * T0:tx * T0 + w] = gathered[r, k, :h, :w]
    return out


def _worker(rank, world, port, dims, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fidget_b200.shard import owned_tiles, tiles_per_rank
    from oracle import oracle as orc
    if len(dims) == 2:
        width, height = dims
        t = orc.Tape.from_vm(open(os.path.join(ROOT, "models", "hi.vm")).read())
        image = np.zeros((height, width, 1), dtype=np.float32)
    else:
        width, height, depth = dims
        t = orc.Tape.from_vm(open(os.path.join(ROOT, "models", "colonnade.vm")).read())
        image = np.zeros((height, width, 4), dtype=np.float32)
    ry = (height + T0 - 1) // T0      # the oracle enumerates root tiles like the reference: x outer, y inner
    mine = owned_tiles(rank, world, width, height, T0)
    for tx, ty in mine:            # this rank renders its own root tiles and nothing else
        if len(dims) == 2:
            part, _ = orc.render2d(t, width, height, first_root=tx * ry + ty, n_roots=1)
            piece = part[ty * T0:(ty + 1) * T0, tx * T0:(tx + 1) * T0]
            image[ty * T0:ty * T0 + piece.shape[0], tx * T0:tx * T0 + piece.shape[1], 0] = piece
        else:
            part, _ = orc.render3d(t, width, height, depth, first_root=tx * ry + ty, n_roots=1)
            piece = part.view(np.float32).reshape(height, width, 4)[ty * T0:(ty + 1) * T0, tx * T0:(tx + 1) * T0]
            image[ty * T0:ty * T0 + piece.shape[0], tx * T0:tx * T0 + piece.shape[1]] = piece
    per = tiles_per_rank(world, width, height, T0)
    chunk = torch.from_numpy(_pack(image, mine, per))
    gathered = torch.empty((world * per,) + tuple(chunk.shape[1:]), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, chunk)            # the ONE collective of the sharded render
    if rank == world - 1:
        np.save(out_path, _unpack(gathered.numpy().reshape((world, per) + tuple(chunk.shape[1:])), world, width, height, image))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dims", [(2, (512, 512)), (4, (512, 512)), (2, (600, 300))])
def test_tile_interleave_all_gather_2d(tmp_path, world, dims, orc):
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), dims, out), nprocs=world, join=True)
    got = np.load(out)[..., 0]
    text = open(os.path.join(ROOT, "models", "hi.vm")).read()
    want, _ = orc.render2d(orc.Tape.from_vm(text), dims[0], dims[1])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_tile_interleave_all_gather_3d(tmp_path, orc):
    dims = (256, 256, 256)
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), dims, out), nprocs=2, join=True)
    got = np.load(out)
    text = open(os.path.join(ROOT, "models", "colonnade.vm")).read()
    want, _ = orc.render3d(orc.Tape.from_vm(text), *dims, threads=4)
    assert got.tobytes() == want.tobytes()


def test_partition_arithmetic():
    from fidget_b200.shard import band_rows, z_slab, owned_tiles, tiles_per_rank
    rows = [band_rows(r, 8, 4096) for r in range(8)]
    assert rows[0] == (0, 4) and rows[-1] == (28, 32)
    assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    slabs = [z_slab(r, 8, 4096) for r in range(8)]
    assert slabs[0] == (0, 512) and slabs[-1] == (3584, 4096)
    with pytest.raises(ValueError):
        band_rows(0, 3, 4096)
    # the tile interleave is a partition of the root grid
    for world in (1, 2, 3, 4, 8):
        for w, h in ((4096, 4096), (1000, 600), (128, 128)):
            owned = [owned_tiles(r, world, w, h) for r in range(world)]
            flat = sorted(t for o in owned for t in o)
            rx, ry = (w + 127) // 128, (h + 127) // 128
            assert flat == sorted((tx, ty) for ty in range(ry) for tx in range(rx))
            assert tiles_per_rank(world, w, h) == max(len(o) for o in owned)
    assert 128 <= tiles_per_rank(8, 4096, 4096) <= 160       # a hash, not a regular pattern: near n/N, not exactly
    # the C ABI agrees (fc_tiles_per_rank needs no device)
    from fidget_b200 import _lib
    L = _lib.load()
    for world in (1, 2, 3, 8):
        assert L.fc_tiles_per_rank(1000, 600, 128, world) == tiles_per_rank(world, 1000, 600)
