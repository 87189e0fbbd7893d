"""the host tape pipeline on a shape edit -- `.vm` text -> Context -> SsaTape -> register
allocation -> bytecode -> wire blob (csrc/host, C++; the counterpart of Context::from_text + VmData::new +
Bytecode::new) -- timed per model on one host core.  No GPU involved.

  python scripts/bench_host_pipeline.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fidget_b200 as fb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name in ("hi.vm", "bear.vm", "colonnade.vm", "prospero.vm"):
    text = open(os.path.join(ROOT, "models", name)).read()
    reps = 20
    t = {"parse": 0.0, "tape": 0.0, "bytecode+blob": 0.0}
    for _ in range(reps):
        t0 = time.perf_counter()
        ctx, root = fb.Context.from_text(text)
        t1 = time.perf_counter()
        td = ctx.tape(root)
        t2 = time.perf_counter()
        blob = td.serialize()
        t3 = time.perf_counter()
        t["parse"] += t1 - t0; t["tape"] += t2 - t1; t["bytecode+blob"] += t3 - t2
    print(json.dumps({"model": name, "clauses": len(td), "blob_bytes": len(blob),
                      "ms": {k: v / reps * 1e3 for k, v in t.items()}, "total_ms": sum(t.values()) / reps * 1e3}))
