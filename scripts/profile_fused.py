import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fidget_b200 as fb
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cuda = fb.CudaContext(0)
s = fb.CudaShape.from_vm(cuda, open(os.path.join(ROOT, "models", "prospero.vm")).read())
img = torch.zeros((4096, 4096), dtype=torch.float32, device="cuda")
for _ in range(3):
    fb.render2d(s, fb.RenderConfig2D(4096, 4096, fused_tail=True), out=img)
cuda.synchronize()
