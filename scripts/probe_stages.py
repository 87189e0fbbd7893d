"""Stage times (CUDA events inside the library) of prospero 4096^2, per-level launches and fused tail, under a few
settings -- a diagnostic for scheduling / memory-placement effects.  usage: python scripts/probe_stages.py [arena_gib]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fidget_b200 as fb
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
arena = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cuda = fb.CudaContext(0)
cuda.set_arena_bytes(arena << 30)
cuda.set_stream(torch.cuda.current_stream().cuda_stream)
shape = fb.CudaShape.from_vm(cuda, open(os.path.join(ROOT, "models", "prospero.vm")).read())
img = torch.zeros((4096, 4096), dtype=torch.float32, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for unfused in (True, False):
    cfg = fb.RenderConfig2D(4096, 4096, timing=True, fused_tail=not unfused)
    for _ in range(3):
        fb.render2d(shape, cfg, out=img, stats=True)
    acc = np.zeros(16)
    for _ in range(5):
        flush.fill_(1)
        _, st = fb.render2d(shape, cfg, out=img, stats=True)
        acc += np.array(st["stage_ms"])
    acc /= 5
    print(json.dumps({"arena_gib": arena, "unfused": unfused, "serial_fill": os.environ.get("FIDGET_B200_SERIAL_FILL", "0"),
                      "L0": acc[0], "L1": acc[1], "L2": acc[2], "fill_tail": acc[8], "pixels": acc[9], "fused_tail": acc[12], "fused_phase_end_ms": [round(float(x), 4) for x in acc[1:8]] if not unfused else None,
                      "fused_latest_start_L1_L2_ms": [round(float(x), 4) for x in acc[8:10]] if not unfused else None,
                      "fused_longest_job_L1_L2_ms": [round(float(x), 4) for x in acc[10:12]] if not unfused else None, "total": acc[15],
                      "launches": st["kernel_launches"]}))
