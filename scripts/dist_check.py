"""torchrun script: sharded renders (2D bands, 3D Z slabs) must equal the single-GPU result byte for byte."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import fidget_b200 as fb
from fidget_b200.shard import render2d_bands, render3d_zslabs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cuda = fb.CudaContext(local)
cuda.set_stream(torch.cuda.current_stream().cuda_stream)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shape = fb.CudaShape.from_vm(cuda, open(os.path.join(root, "models", "prospero.vm")).read())
dev = torch.device("cuda", local)
ok = True
# 2D
n = 2048
img = torch.zeros((n, n), dtype=torch.float32, device=dev)
gat = torch.empty_like(img)
render2d_bands(shape, fb.RenderConfig2D(n, n), img, gat)
torch.cuda.synchronize()
full = torch.zeros((n, n), dtype=torch.float32, device=dev)
fb.render2d(shape, fb.RenderConfig2D(n, n), out=full)
same2d = bool(torch.equal(gat.view(torch.int32), full.view(torch.int32)))
# 3D
col = fb.CudaShape.from_vm(cuda, open(os.path.join(root, "models", "colonnade.vm")).read())
m = 512
slab = torch.zeros((m, m, 4), dtype=torch.float32, device=dev)
g3 = torch.empty((world, m, m, 4), dtype=torch.float32, device=dev)
out3 = torch.zeros((m, m, 4), dtype=torch.float32, device=dev)
render3d_zslabs(col, fb.RenderConfig3D(m, m, m), slab, g3, out3)
torch.cuda.synchronize()
full3 = torch.zeros((m, m, 4), dtype=torch.float32, device=dev)
fb.render3d(col, fb.RenderConfig3D(m, m, m), out=full3)
same3d = bool(torch.equal(out3.view(torch.int32), full3.view(torch.int32)))
print(f"rank {rank}/{world}: 2D bands identical={same2d}  3D slabs identical={same3d}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (same2d and same3d) else 1)
