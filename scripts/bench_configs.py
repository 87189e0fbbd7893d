"""Measurements of the other BASELINE.json configs (not bench.py lines): prints one JSON line each.

  configs[2]  bear.vm 3D heightmap + normals 1024^3
  configs[3]  gyroid-sphere octree sampler, depth 9 (MDC sampling half)
  configs[4]  prospero.vm 3D 4096^3, ONE Z slab of the 8-way split (512 deep) on one GPU + full-depth 1024^3
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fidget_b200 as fb

cuda = fb.CudaContext(0)
cuda.set_arena_bytes(8 << 30)
which = sys.argv[1:] or ["bear", "gyroid", "slab"]


def model(name):
    return open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "models", name)).read()


def time_render3d(shape, cfg, out, reps=5):
    best = None
    for _ in range(reps):
        _, st = fb.render3d(shape, cfg, out=out, stats=True)
        if best is None or st["stage_ms"][15] < best["stage_ms"][15]:
            best = st
    return best


if "bear" in which:
    shape = fb.CudaShape.from_vm(cuda, model("bear.vm"))
    n = 1024
    out = torch.empty((n, n, 4), dtype=torch.float32, device="cuda")
    st = time_render3d(shape, fb.RenderConfig3D(n, n, n, timing=True), out)
    ms = st["stage_ms"]
    from oracle import oracle as orc
    ot = orc.Tape.from_vm(model("bear.vm"))
    threads = os.cpu_count()
    t0 = time.perf_counter()
    orc.render3d(ot, n, n, n, threads=threads)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({"config": "bear.vm 3D heightmap+normals 1024^3", "ms": ms[15], "Mvoxels_per_s": n ** 3 / ms[15] / 1e3,
                      "cpu_oracle_ms": cpu_s * 1e3, "cpu_threads": threads,
                      "levels_ms": ms[:5], "voxels_ms": ms[9], "normals_ms": ms[10], "voxel_evals": st["pixels"],
                      "grad_evals": st["grads"], "arena_MB": st["arena_bytes_used"] / 1e6}))
if "gyroid" in which:
    shape = fb.CudaShape.from_vm(cuda, model("gyroid-sphere.vm"))
    from oracle import oracle as orc
    ot = orc.Tape.from_vm(model("gyroid-sphere.vm"))
    t0 = time.perf_counter()
    orc.octree_sample(ot, 7)
    print(json.dumps({"config": "gyroid-sphere octree sampler depth 7, CPU oracle (1 thread)",
                      "ms": (time.perf_counter() - t0) * 1e3}))
    for depth in (7, 8, 9):
        fb.octree_sample(shape, depth, capacity=8 << 20 if depth == 9 else None)   # warm-up
        t0 = time.perf_counter()
        leaves, st = fb.octree_sample(shape, depth, stats=True, timing=True, capacity=8 << 20 if depth == 9 else None)
        wall = time.perf_counter() - t0
        print(json.dumps({"config": f"gyroid-sphere octree sampler depth {depth}", "device_ms": st["total_ms"],
                          "wall_ms_incl_d2h_and_sort": wall * 1e3, "Mcells_per_s": 8 ** depth / st["total_ms"] / 1e3,
                          "surface_leaves": len(leaves), "float_points": st["float_points"],
                          "grad_points": st["grad_points"], "ambiguous": st["ambiguous"][:depth + 1]}))
if "slab" in which:
    shape = fb.CudaShape.from_vm(cuda, model("prospero.vm"))
    n = 4096
    out = torch.empty((n, n, 4), dtype=torch.float32, device="cuda")
    cfg = fb.RenderConfig3D(n, n, n, z_range=(n - 512, n), clamp=False, timing=True)   # the front slab of 8
    st = time_render3d(shape, cfg, out, reps=3)
    ms = st["stage_ms"]
    print(json.dumps({"config": "prospero.vm 3D 4096^3, front Z slab [3584,4096) of an 8-way split, 1 GPU",
                      "ms": ms[15], "Mvoxels_per_s_slab": n * n * 512 / ms[15] / 1e3, "levels_ms": ms[:5],
                      "voxels_ms": ms[9], "normals_ms": ms[10], "voxel_evals": st["pixels"],
                      "arena_MB": st["arena_bytes_used"] / 1e6}))
