#!/usr/bin/env python
"""Benchmark of the tape-evaluation hot path on B200 ("Measurement").

A step = one pass of the hot path over one frame of synthetic input:
  N = 1 : models/prospero.vm, 2D render 4096x4096 (BASELINE.json configs[1]):
          interval levels [128,32,8] with on-device tape simplification, fill,
          bulk f32 over the surviving leaf tiles.
  N > 1 : default ("weak"): one 4096x4096 Z slice of the same model per GPU (rank r renders the
          voxel layer z_r of a 4096^3 grid), no collective on the data path -- the slices are
          independent units; a step = N slices, value = N * 4096^2 / max-over-ranks time.
          --scaling strong: ONE frame sharded by bands of root-tile rows + ONE NCCL all-gather
          of the bands (total work fixed).  A 0.32 ms frame is a chain of four latency-bound
          launches, so bands do not shorten it (measured: profiles/r01_bench_n2.json).

  python bench.py --gpus N --steps K --warmup W           # CUDA arm
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host threads)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE = 4096
MODEL = "prospero.vm"
METRIC = "Mvoxels/s (prospero 4096^2 2D render, interval + bulk f32)"
ALGO_BYTES_PER_PIXEL = 4  # one RawDistancePixel written per pixel


def model_text():
    with open(os.path.join(ROOT, "models", MODEL)) as f:
        return f.read()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.stop = False
        self.index = index
        self.t = None
        # NVML answers in ~0.1 ms, nvidia-smi (the same counters through a subprocess) in ~50 ms; the
        # timed region of a default run lasts ~10 ms, so NVML is what can sample it more than once
        self.nvml, self.handle = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        p, h = self.nvml, self.handle
        sm = p.nvmlDeviceGetClockInfo(h, p.NVML_CLOCK_SM)
        mx = p.nvmlDeviceGetMaxClockInfo(h, p.NVML_CLOCK_SM)
        try:
            r = p.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = p.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        try:
            w = p.nvmlDeviceGetPowerUsage(h) / 1000.0
        except Exception:
            w = 0.0
        bits = (p.nvmlClocksThrottleReasonHwSlowdown, p.nvmlClocksThrottleReasonHwThermalSlowdown,
                p.nvmlClocksThrottleReasonSwThermalSlowdown, p.nvmlClocksThrottleReasonSwPowerCap)
        self.rows.append([str(sm), str(mx), f"{w:.1f}"] + ["Active" if r & b else "Not Active" for b in bits])

    def _sample(self):
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                  "-i", str(self.index)], capture_output=True, text=True, timeout=10).stdout
            for line in out.strip().splitlines():
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def _run(self):
        while not self.stop:
            if self.nvml is not None:
                try:
                    self._sample_nvml()
                    time.sleep(0.001)
                    continue
                except Exception:
                    self.nvml = None          # fall back to nvidia-smi for the rest of the run
            self._sample()
            time.sleep(0.05)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=15)
        if not self.rows:
            self._sample()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4)
                          if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs"), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(sample_frames=1, threads=None):
    """The CPU oracle (a port of the reference's VmShape path) timed on this host's cores."""
    from oracle import oracle as orc
    threads = threads or os.cpu_count() or 1
    t = orc.Tape.from_vm(model_text())
    orc.render2d(t, 512, 512, threads=threads)  # warm the library / page in
    t0 = time.perf_counter()
    for _ in range(sample_frames):
        orc.render2d(t, SIZE, SIZE, threads=threads)
    dt = (time.perf_counter() - t0) / sample_frames
    return {"value": SIZE * SIZE / dt / 1e6, "unit": "Mvoxels/s", "cores": threads, "kind": "port",
            "sample": f"{sample_frames} full {SIZE}x{SIZE} frame(s) of {MODEL}, tile sizes [128,32,8], "
                      f"{threads} threads over root tiles (oracle/vm.cc render2d)",
            "seconds_per_frame": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import oracle as orc
    t = orc.Tape.from_vm(model_text())
    for _ in range(max(args.warmup, 1)):
        orc.render2d(t, SIZE, SIZE, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.render2d(t, SIZE, SIZE, threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    v = SIZE * SIZE / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mvoxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": args.scaling if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"models/{MODEL} 2D render {SIZE}x{SIZE}, tile sizes [128,32,8], identity camera"},
        "cpu_baseline": {"value": v, "unit": "Mvoxels/s", "cores": threads, "kind": "port",
                         "sample": f"each step = one full {SIZE}x{SIZE} frame, {threads} host threads; the Rust "
                                   "reference cannot be built here (no rustc), so this is the C++ oracle port of "
                                   "VmShape + fidget-raster::pixel::render"},
        "e2e": {"value": v, "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = one Z slice per GPU (default), strong = one frame in bands + all-gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import fidget_b200 as fb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    cuda = fb.CudaContext(local)
    stream = torch.cuda.current_stream()
    cuda.set_stream(stream.cuda_stream)
    ctx, root = fb.Context.from_text(model_text())
    tape = ctx.tape(root)
    shape = fb.CudaShape(cuda, tape)

    from fidget_b200.shard import band_rows
    T0 = 128
    n_rows = SIZE // T0
    strong = world > 1 and args.scaling == "strong"
    rows = band_rows(rank, world, SIZE, T0) if strong else (0, 0)
    # weak scaling: rank r renders voxel layer z_r of the SIZE^3 grid around z = 0 (region.rs:87-108: one voxel = 2/SIZE)
    z_slice = (rank - (world - 1) / 2.0) * (2.0 / SIZE) if world > 1 and not strong else 0.0
    cfg = fb.RenderConfig2D(SIZE, SIZE, root_rows=rows, z=z_slice)
    image = torch.zeros((SIZE, SIZE), dtype=torch.float32, device=dev)
    gathered = torch.empty_like(image) if strong else None
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    from fidget_b200.shard import render2d_bands
    full_cfg = fb.RenderConfig2D(SIZE, SIZE)

    def step():
        if strong:
            render2d_bands(shape, full_cfg, image, gathered)   # band render + ONE all-gather
        else:
            fb.render2d(shape, cfg, out=image, asynchronous=True)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    cuda.synchronize()  # surfaces deferred device errors

    # ---- device-resident timing: K steps, L2 flushed (untimed) between steps ----
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sync_all()
    with ClockSampler(local) as clocks:
        for i in range(args.steps):
            flush.fill_(i & 255)
            starts[i].record(stream)
            step()
            stops[i].record(stream)
        sync_all()
    cuda.synchronize()
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    frames_per_step = 1 if (world == 1 or strong) else world       # weak: every rank renders its own slice
    value = frames_per_step * SIZE * SIZE / (ms_per_step * 1e-3) / 1e6

    # ---- per-kernel timing of one step (CUDA events inside the library) ----
    tcfg = fb.RenderConfig2D(SIZE, SIZE, root_rows=rows, z=z_slice, timing=True)
    stage = np.zeros(16)
    reps = 5
    stats = None
    for _ in range(reps):
        flush.fill_(1)
        _, stats = fb.render2d(shape, tcfg, out=image, stats=True)
        stage += np.array(stats["stage_ms"])
    stage /= reps
    names = {0: "k_interval_root_coop_2d[L0,128px]", 1: "k_interval_level<2>[L1,32px]",
             2: "k_interval_level<2>[L2,8px]", 8: "k_fill_2d (x3)", 9: "k_pixels_2d"}
    dom = max(names, key=lambda k: stage[k])
    frac_rows = (rows[1] - rows[0]) / n_rows if strong else 1.0
    # units decided by one launch of the dominant kernel (DESIGN.md "Measurement")
    if dom in (0, 1, 2):
        tile = [128, 32, 8][dom]
        units = stats["evaluated"][dom] * tile * tile
    elif dom == 9:
        units = stats["pixels"]
    else:
        units = int(SIZE * SIZE * frac_rows) - stats["pixels"]
    peak, peak_src = measured_peaks()
    achieved = units * ALGO_BYTES_PER_PIXEL / (stage[dom] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(names[dom])
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": float(stage[dom]), "units_per_launch": int(units),
                "frame": {"algorithmic_bytes": SIZE * SIZE * ALGO_BYTES_PER_PIXEL,
                          "achieved": SIZE * SIZE * ALGO_BYTES_PER_PIXEL / (ms_per_step * 1e-3) / 1e9,
                          "frac": SIZE * SIZE * ALGO_BYTES_PER_PIXEL / (ms_per_step * 1e-3) / 1e9 / peak},
                "stage_ms": {names[k]: float(stage[k]) for k in names}}

    # ---- end to end through the public API with HOST buffers ----
    host_img = torch.empty((SIZE, SIZE), dtype=torch.float32).pin_memory()
    host_np = host_img.numpy()
    bc = tape.bytecode()

    def e2e_step():
        s = fb.CudaShape(cuda, tape)          # uploads the bytecode (H2D) and builds the device tape
        fb.render2d(s, cfg, out=host_np)      # renders + copies the image back (D2H), synchronous
        return s

    for _ in range(3):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    n_e2e = max(5, min(args.steps, 20))
    for _ in range(n_e2e):
        e2e_step()
    sync_all()
    e2e_dt = (time.perf_counter() - t0) / n_e2e
    et = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_dt = float(et.item())
    e2e = {"value": frames_per_step * SIZE * SIZE / e2e_dt / 1e6, "unit": "Mvoxels/s",
           "h2d_bytes_per_step": int(bc.words.nbytes) * frames_per_step,
           "d2h_bytes_per_step": int(SIZE * SIZE * 4 * frac_rows) * frames_per_step,
           "ms_per_step": e2e_dt * 1e3,
           "note": "fc_tape_create from host bytecode + fc_render2d into a pinned host image, wall clock"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mvoxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"models/{MODEL} 2D render {SIZE}x{SIZE}, tile sizes [128,32,8] (reference VM "
                                   "defaults), identity camera, pixel_perfect=false",
                       "parallelism": "single GPU" if world == 1 else (
                           f"{world} bands of root-tile rows of ONE frame + 1 all-gather" if strong else
                           f"{world} independent {SIZE}x{SIZE} Z slices (layers z_r of a {SIZE}^3 grid), one per GPU, "
                           "no data-path collective; value = all slices / max-over-ranks time"),
                       "l2": "flushed between steps by a 512 MiB fill outside the event-timed regions"},
            "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": int(stats["kernel_launches"]) * args.steps * frames_per_step,
            "roofline": roofline,
        }
        if not args.no_cpu_baseline and world == 1:      # the CPU baseline is an N = 1 measurement
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
