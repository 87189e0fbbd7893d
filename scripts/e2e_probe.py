"""Where the end-to-end step goes: tape upload, render into a pinned host image, raw copy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fidget_b200 as fb
cuda = fb.CudaContext(0)
ctx, root = fb.Context.from_text(open("models/prospero.vm").read())
tape = ctx.tape(root)
N = 4096
host = torch.empty((N, N), dtype=torch.float32).pin_memory(); hnp = host.numpy()
dev = torch.empty((N, N), dtype=torch.float32, device="cuda")
cfg = fb.RenderConfig2D(N, N)
def best(fn, reps=20):
    fn(); fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, sorted(ts)[len(ts) // 2] * 1e3
shape = fb.CudaShape(cuda, tape)
print("bytecode() ms", best(lambda: tape.bytecode()))
print("CudaShape create+destroy ms", best(lambda: fb.CudaShape(cuda, tape)))
print("render2d -> device ms", best(lambda: fb.render2d(shape, cfg, out=dev)))
print("render2d -> pinned host ms", best(lambda: fb.render2d(shape, cfg, out=hnp)))
print("raw D2H 67 MB ms", best(lambda: host.copy_(dev, non_blocking=True)))
def e2e():
    s = fb.CudaShape(cuda, tape); fb.render2d(s, cfg, out=hnp)
print("e2e step ms", best(e2e))
