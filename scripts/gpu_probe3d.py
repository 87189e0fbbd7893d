"""Ad-hoc GPU probe: render3d stage timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fidget_b200 as fb
cuda = fb.CudaContext(0)
for name, size in (("bear.vm", 512), ("bear.vm", 1024), ("prospero.vm", 1024), ("colonnade.vm", 1024)):
    shape = fb.CudaShape.from_vm(cuda, open(f"models/{name}").read())
    out = torch.empty((size, size, 4), dtype=torch.float32, device="cuda")
    cfg = fb.RenderConfig3D(size, size, size, timing=True)
    best = None
    for _ in range(4):
        _, st = fb.render3d(shape, cfg, out=out, stats=True)
        if best is None or st["stage_ms"][15] < best["stage_ms"][15]:
            best = st
    ms = best["stage_ms"]
    print(name, size, "total ms %.3f" % ms[15], "levels", [round(v, 3) for v in ms[:5]], "voxels %.3f" % ms[9],
          "normals %.3f" % ms[10], "Mvox/s %.0f" % (size ** 3 / ms[15] / 1e3))
    print("   evaluated", best["evaluated"][:5], "amb", best["ambiguous"][:5], "voxels", best["pixels"], "grads", best["grads"],
          "arena MB %.1f" % (best["arena_bytes_used"] / 1e6))
