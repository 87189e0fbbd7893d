"""Ad-hoc GPU probe: stage timings of render2d on prospero at several sizes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fidget_b200 as fb

cuda = fb.CudaContext(0)
text = open("models/prospero.vm").read()
shape = fb.CudaShape.from_vm(cuda, text)
res = {}
for size in (1024, 4096):
    out = torch.empty((size, size), dtype=torch.float32, device="cuda")
    cfg = fb.RenderConfig2D(size, size, timing=True)
    for _ in range(3):
        _, st = fb.render2d(shape, cfg, out=out, stats=True)
    best = None
    for _ in range(5):
        _, st = fb.render2d(shape, cfg, out=out, stats=True)
        if best is None or st["stage_ms"][15] < best["stage_ms"][15]:
            best = st
    res[size] = best
    print(size, "total ms", best["stage_ms"][15], "levels", best["stage_ms"][:3], "fill", best["stage_ms"][8],
          "pixels", best["stage_ms"][9], "Mpx/s", size * size / best["stage_ms"][15] / 1e3)
    print("   census", best["evaluated"][:3], best["ambiguous"][:3], best["pixels"], "arena MB", best["arena_bytes_used"] / 1e6)
json.dump(res, open("gpurun_out/probe.json", "w"))
