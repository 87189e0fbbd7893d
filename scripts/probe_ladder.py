"""3D render with the reference's tile ladder [128,64,32,16,8] against coarser ladders: same image? how fast?"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fidget_b200 as fb
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cuda = fb.CudaContext(0)
cuda.set_arena_bytes(8 << 30)
for name, n, zr in (("bear.vm", 1024, None), ("prospero.vm", 4096, (3584, 4096)), ("prospero.vm", 1024, None), ("colonnade.vm", 512, None)):
    shape = fb.CudaShape.from_vm(cuda, open(os.path.join(ROOT, "models", name)).read())
    ref = None
    for ts in ((128, 64, 32, 16, 8), (128, 32, 8), (128, 64, 16, 8), (128, 32, 16, 8), (64, 16, 8), (128, 16, 8)):
        out = torch.empty((n, n, 4), dtype=torch.float32, device="cuda")
        kw = dict(tile_sizes=ts, timing=True)
        if zr: kw.update(z_range=zr, clamp=False)
        cfg = fb.RenderConfig3D(n, n, n, **kw)
        best = None
        for _ in range(3):
            _, st = fb.render3d(shape, cfg, out=out, stats=True)
            if best is None or st["stage_ms"][15] < best["stage_ms"][15]: best = st
        img = out.cpu().numpy().view(np.uint32)
        if ref is None: ref = img
        ms = best["stage_ms"]
        print(json.dumps({"model": name, "n": n, "tile_sizes": ts, "ms": round(ms[15], 3), "levels_ms": [round(x, 3) for x in ms[:len(ts)]],
                          "voxels_ms": round(ms[9], 3), "normals_ms": round(ms[10], 3), "same_image": bool(np.array_equal(img, ref)),
                          "n_diff": int((img != ref).any(axis=-1).sum()), "voxel_evals": best["pixels"], "evaluated": best["evaluated"][:len(ts)]}))
