"""Single render for profiling: python scripts/one_render.py [size] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fidget_b200 as fb
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cuda = fb.CudaContext(0)
shape = fb.CudaShape.from_vm(cuda, open("models/prospero.vm").read())
out = torch.empty((size, size), dtype=torch.float32, device="cuda")
for _ in range(reps):
    fb.render2d(shape, fb.RenderConfig2D(size, size), out=out)
torch.cuda.synchronize()
