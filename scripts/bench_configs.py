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
which = sys.argv[1:] or ["bear", "gyroid", "mesh", "census", "slab", "volume", "effects"]


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
if "mesh" in which:
    # fc_mesh_build: sampler + QEF vertices + dual walk, everything resident in HBM; STL assembled on the device
    shape = fb.CudaShape.from_vm(cuda, model("gyroid-sphere.vm"))
    for depth in (7, 8, 9):
        fb.mesh(shape, depth)                      # warm-up (buffers)
        t0 = time.perf_counter()
        v, t, info = fb.mesh(shape, depth)
        wall = time.perf_counter() - t0
        print(json.dumps({"config": f"gyroid-sphere mesh depth {depth} (fc_mesh_build: sampler + QEF + dual walk on device)",
                          "sampler_ms": info["sampler_ms"], "qef_and_dual_walk_ms": info["mesh_ms"], "leaves": info["n_leaves"],
                          "vertices": info["n_vertices"], "triangles": info["n_triangles"], "open_edges": info["open_edges"],
                          "wall_ms_incl_readback": wall * 1e3, "readback_MB": (v.nbytes + t.nbytes) / 1e6}))
if "census" in which:
    # cost of the exact reference census (FC_FLAG_EXACT_CENSUS) on top of a render with statistics
    shape = fb.CudaShape.from_vm(cuda, model("bear.vm"))
    n = 1024
    out = torch.empty((n, n, 4), dtype=torch.float32, device="cuda")
    a = time_render3d(shape, fb.RenderConfig3D(n, n, n, timing=True), out)
    b = time_render3d(shape, fb.RenderConfig3D(n, n, n, timing=True, exact_census=True), out)
    print(json.dumps({"config": "bear.vm 1024^3: device census vs exact reference census", "ms_plain": a["stage_ms"][15],
                      "ms_exact_census": b["stage_ms"][15], "evaluated_device": a["evaluated"][:5], "evaluated_reference": b["evaluated"][:5],
                      "voxels_device": a["pixels"], "voxels_reference": b["pixels"]}))
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
if "volume" in which:
    # the N > 1 bench workload on one GPU (bench.py's strong_scaling_base)
    shape = fb.CudaShape.from_vm(cuda, model("prospero.vm"))
    n = 4096
    out = torch.empty((n, n, 4), dtype=torch.float32, device="cuda")
    st = time_render3d(shape, fb.RenderConfig3D(n, n, n, timing=True), out, reps=4)
    ms = st["stage_ms"]
    print(json.dumps({"config": "prospero.vm 3D 4096^3 whole volume, 1 GPU", "ms": ms[15], "levels_ms": ms[:5], "voxels_ms": ms[9],
                      "normals_ms": ms[10], "voxel_evals": st["pixels"], "evaluated": st["evaluated"][:5],
                      "arena_MB": st["arena_bytes_used"] / 1e6}))
if "effects" in which:
    # fidget-raster's viewer post-processing on a device-resident bear.vm 1024^3 heightmap:
    # HBM-bound byte work; GB/s = algorithmic bytes (read 16 B GeometryPixel + write result) / time
    from fidget_b200 import effects as fx
    from oracle import oracle as orc
    shape = fb.CudaShape.from_vm(cuda, model("bear.vm"))
    n = 1024
    geo = torch.empty((n, n, 4), dtype=torch.float32, device="cuda")
    fb.render3d(shape, fb.RenderConfig3D(n, n, n), out=geo)
    den = torch.empty_like(geo)
    ssao = torch.empty((n, n), dtype=torch.float32, device="cuda")
    blur = torch.empty_like(ssao)
    rgb = torch.empty((n, n, 3), dtype=torch.uint8, device="cuda")
    rgba = torch.empty((4096, 4096, 4), dtype=torch.uint8, device="cuda")
    img2d = torch.empty((4096, 4096), dtype=torch.float32, device="cuda")
    fb.render2d(fb.CudaShape.from_vm(cuda, model("prospero.vm")), fb.RenderConfig2D(4096, 4096), out=img2d)
    k, nz = fx.ssao_kernel(64), fx.ssao_noise(256)
    cuda.set_stream(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, reps=10):
        fn()
        best = 1e9
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best
    px = n * n
    cases = [
        ("denoise_normals 1024^2", lambda: fx.denoise_normals(cuda, geo, out=den), px * 32),
        ("compute_ssao 1024^2 (64 samples)", lambda: fx.compute_ssao(cuda, den, n, k, nz, out=ssao), px * 20),
        ("blur_ssao 1024^2", lambda: fx.blur_ssao(cuda, ssao, out=blur), px * 8),
        ("apply_shading(ssao) 1024^2 fused pipeline", lambda: fx.apply_shading(cuda, den, n, True, k, nz, out=rgb), px * 19),
        ("apply_shading(no ssao) 1024^2", lambda: fx.apply_shading(cuda, den, n, False, out=rgb), px * 19),
        ("to_rgba_bitmap 4096^2", lambda: fx.to_rgba_bitmap(cuda, img2d, out=rgba), 4096 * 4096 * 8),
        ("to_rgba_distance 4096^2", lambda: fx.to_rgba_distance(cuda, img2d, out=rgba), 4096 * 4096 * 8),
    ]
    for name, fn, nbytes in cases:
        ms = timed(fn)
        print(json.dumps({"config": "effects: " + name, "ms": ms, "algorithmic_GB_per_s": nbytes / ms / 1e6}))
    fx.apply_shading(cuda, den, n, True, k, nz, out=rgb)
    den_h = den.cpu().numpy().view(fb.GEOMETRY_PIXEL).reshape(n, n)
    t0 = time.perf_counter()
    want = orc.apply_shading(den_h, n, orc.blur_ssao(orc.compute_ssao(den_h, n, k, nz)))
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"config": "effects: apply_shading(ssao) 1024^2, CPU oracle (1 thread)", "ms": cpu_ms,
                      "identical_to_gpu": bool(np.array_equal(want, rgb.cpu().numpy()))}))
