"""Small workloads for ncu captures (one GPU):  python scripts/profile_targets.py frame2d | slices | volume3d | mesh"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fidget_b200 as fb
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
what = sys.argv[1] if len(sys.argv) > 1 else "frame2d"
cuda = fb.CudaContext(0)
cuda.set_arena_bytes(4 << 30)
model = lambda n: fb.CudaShape.from_vm(cuda, open(os.path.join(ROOT, "models", n)).read())
if what == "frame2d":
    s = model("prospero.vm")
    img = torch.zeros((4096, 4096), dtype=torch.float32, device="cuda")
    for _ in range(3):
        fb.render2d(s, fb.RenderConfig2D(4096, 4096), out=img)
elif what == "slices":
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from bench_slices import csg_tape
    s = fb.CudaShape(cuda, csg_tape())
    n = 1 << 24
    rng = np.random.default_rng(0)
    xyz = [torch.from_numpy(rng.uniform(-1, 1, n).astype(np.float32)).cuda() for _ in range(3)]
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    g = [torch.zeros((n, 4), dtype=torch.float32, device="cuda") for _ in range(3)]
    gout = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    for _ in range(3):
        s.float_slice_eval(xyz, out=out)
        s.grad_slice_eval(g, out=gout)
elif what == "volume3d":
    s = model("prospero.vm")
    img = torch.zeros((1024, 1024, 4), dtype=torch.float32, device="cuda")
    for _ in range(2):
        fb.render3d(s, fb.RenderConfig3D(1024, 1024, 1024), out=img)
    s = model("bear.vm")
    for _ in range(2):
        fb.render3d(s, fb.RenderConfig3D(1024, 1024, 1024), out=img)
elif what == "mesh":
    s = model("gyroid-sphere.vm")
    for _ in range(2):
        fb.mesh(s, 8)
cuda.synchronize()
