"""Host-side logic of the N>1 path on CPU: two gloo ranks shard a frame into bands of root-tile
rows, render their band (the CPU oracle stands in for the GPU renderer), all-gather, and must
reproduce the single-process image byte for byte."""
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


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, size, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fidget_b200.shard import band_rows, band_pixels
    from oracle import oracle as orc
    text = open(os.path.join(ROOT, "models", "hi.vm")).read()
    t = orc.Tape.from_vm(text)
    rows = band_rows(rank, world, size)
    y0, y1 = band_pixels(rows, size, size)
    full, _ = orc.render2d(t, size, size)            # stand-in renderer; a rank only contributes its band
    band = torch.from_numpy(np.ascontiguousarray(full[y0:y1]))
    gathered = torch.empty((size, size), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, band)
    if rank == 0:
        np.save(out_path, gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_band_sharding_all_gather(tmp_path, world, orc):
    size = 512
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), size, out), nprocs=world, join=True)
    got = np.load(out)
    text = open(os.path.join(ROOT, "models", "hi.vm")).read()
    want, _ = orc.render2d(orc.Tape.from_vm(text), size, size)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_partition_arithmetic():
    from fidget_b200.shard import band_rows, z_slab
    rows = [band_rows(r, 8, 4096) for r in range(8)]
    assert rows[0] == (0, 4) and rows[-1] == (28, 32)
    assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    slabs = [z_slab(r, 8, 4096) for r in range(8)]
    assert slabs[0] == (0, 512) and slabs[-1] == (3584, 4096)
    with pytest.raises(ValueError):
        band_rows(0, 3, 4096)
