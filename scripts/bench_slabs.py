"""torchrun script: BASELINE configs[4] -- prospero.vm 3D at N^3 voxels sharded into Z slabs over the ranks,
one NCCL all-gather of the slab images, per-pixel merge on every rank.  Prints one JSON line (rank 0):
step time = max over ranks of CUDA-event time around [render slab, all-gather, merge].

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_slabs.py 4096
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import fidget_b200 as fb
from fidget_b200.shard import render3d_zslabs, z_slab

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else "zslabs"      # or "ybands"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
cuda = fb.CudaContext(local)
cuda.set_arena_bytes(8 << 30)
cuda.set_stream(torch.cuda.current_stream().cuda_stream)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shape = fb.CudaShape.from_vm(cuda, open(os.path.join(root, "models", "prospero.vm")).read())
cfg = fb.RenderConfig3D(n, n, n)
slab = torch.zeros((n, n, 4), dtype=torch.float32, device=dev)
gathered = torch.empty((world, n, n, 4), dtype=torch.float32, device=dev)
out = torch.zeros((n, n, 4), dtype=torch.float32, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
times, render_ms = [], []
for it in range(2 + steps):
    flush.zero_()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    from dataclasses import replace
    if mode == "ybands":
        from fidget_b200.shard import band_rows
        rows = band_rows(rank, world, n, 128)
        fb.render3d(shape, replace(cfg, root_rows=rows), out=slab, asynchronous=True)
        e1.record()
        dist.all_gather_into_tensor(out, slab[rows[0] * 128:rows[1] * 128])
    else:
        fb.render3d(shape, replace(cfg, z_range=z_slab(rank, world, n, 128), clamp=False), out=slab, asynchronous=True)
        e1.record()
        dist.all_gather_into_tensor(gathered, slab)
        import ctypes as C
        from fidget_b200 import _lib
        from fidget_b200.shape import _ck
        ptrs = (C.c_void_p * world)(*[gathered[r].data_ptr() for r in range(world)])
        _ck(_lib.load().fc_merge_slabs(cuda._h, ptrs, world, n, n, n, C.c_void_p(out.data_ptr())))
    e2.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e2), e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if it >= 2:
        times.append(float(t[0])); render_ms.append(float(t[1]))
hit = int((out[..., 3].view(torch.int32) > 0).sum())
if rank == 0:
    ms = sum(times) / len(times)
    print(json.dumps({"config": f"prospero.vm 3D {n}^3, {world} " + ("Y bands (full depth) + all-gather" if mode == "ybands" else "Z slabs + all-gather + merge"), "n_gpus": world, "ms_per_step": ms,
                      "Mvoxels_per_s": n ** 3 / ms / 1e3, "slowest_slab_render_ms": sum(render_ms) / len(render_ms),
                      "gather_merge_ms": ms - sum(render_ms) / len(render_ms), "gathered_MB_per_rank": (1 if mode == "ybands" else world) * n * n * 16 / 1e6,
                      "pixels_hit": hit, "steps": steps}), flush=True)
dist.barrier()
dist.destroy_process_group()
