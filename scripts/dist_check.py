"""torchrun script (NCCL, N GPUs): the sharded renders of fidget_b200.shard must reproduce the single-GPU images
byte for byte on every rank.  Prints one line per check on rank 0; exits non-zero on a mismatch.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import fidget_b200 as fb
from fidget_b200 import shard

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
cuda = fb.CudaContext(local)
cuda.set_arena_bytes(4 << 30)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ok = True


def check(name, a, b):
    global ok
    same = torch.equal(a.view(torch.int32), b.view(torch.int32))
    t = torch.tensor([int(same)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{name}: {'identical on all ranks' if int(t) else 'MISMATCH'}", flush=True)
    ok = ok and bool(int(t))


def model(name):
    return fb.CudaShape.from_vm(cuda, open(os.path.join(root, "models", name)).read())


# 2D: interleaved tiles and bands
s = model("prospero.vm")
n = 2048
cfg = fb.RenderConfig2D(n, n)
full = torch.zeros((n, n), dtype=torch.float32, device=dev)
fb.render2d(s, cfg, out=full)
img = torch.zeros_like(full)
chunk, gathered = shard.tile_buffers(world, n, n, 1, dev)
shard.render2d_tiles(s, cfg, img, chunk, gathered)
torch.cuda.synchronize()
check(f"2D prospero {n}^2, {world} ranks, interleaved tiles + 1 all-gather", img, full)
if (n // 128) % world == 0:
    g2 = torch.zeros_like(full)
    shard.render2d_bands(s, cfg, img, g2)
    torch.cuda.synchronize()
    check(f"2D prospero {n}^2, {world} ranks, row bands + 1 all-gather", g2, full)

# 3D: interleaved tiles, Y bands, Z slabs
for name, n in (("bear.vm", 512), ("prospero.vm", 1024)):
    s = model(name)
    cfg = fb.RenderConfig3D(n, n, n)
    full = torch.zeros((n, n, 4), dtype=torch.float32, device=dev)
    fb.render3d(s, cfg, out=full)
    img = torch.zeros_like(full)
    chunk, gathered = shard.tile_buffers(world, n, n, 4, dev)
    shard.render3d_tiles(s, cfg, img, chunk, gathered)
    torch.cuda.synchronize()
    check(f"3D {name} {n}^3, {world} ranks, interleaved tile columns + 1 all-gather", img, full)
    if (n // 128) % world == 0:
        g2 = torch.zeros_like(full)
        shard.render3d_ybands(s, cfg, img, g2)
        torch.cuda.synchronize()
        check(f"3D {name} {n}^3, {world} ranks, Y bands + 1 all-gather", g2, full)
        slab = torch.zeros_like(full)
        gs = torch.zeros((world, n, n, 4), dtype=torch.float32, device=dev)
        out = torch.zeros_like(full)
        shard.render3d_zslabs(s, cfg, slab, gs, out)
        torch.cuda.synchronize()
        check(f"3D {name} {n}^3, {world} ranks, Z slabs + 1 all-gather + merge", out, full)
cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
