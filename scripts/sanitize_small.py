"""A few small renders and evaluator calls for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py`:
exercises every kernel family once (the f32 interpreters fetch one clause past a tape's end, into its padding)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fidget_b200 as fb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cuda = fb.CudaContext(0)
cuda.set_arena_bytes(64 << 20)
for name in ("prospero.vm", "bear.vm"):
    shape = fb.CudaShape.from_vm(cuda, open(os.path.join(ROOT, "models", name)).read())
    img = fb.render2d(shape, fb.RenderConfig2D(256, 256))
    geo = fb.render3d(shape, fb.RenderConfig3D(128, 128, 128))
    pts = np.random.default_rng(0).uniform(-1, 1, (3, 5000)).astype(np.float32)
    shape.eval_f32(pts[0], pts[1], pts[2]) if hasattr(shape, "eval_f32") else None
    leaves = fb.octree_sample(shape, 4)
    print(name, "ok", float(np.isfinite(img).mean()), len(leaves))
