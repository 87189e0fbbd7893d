"""Per-kernel table from an ncu report: duration, DRAM bytes, issue-active %, fp32 (fma) / alu pipe %, warps active.
usage: python scripts/ncu_table.py report.ncu-rep [names.json]  -> JSON on stdout"""
import csv, io, json, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
def col(d, *names):
    for n in names:
        if n in d and d[n] != "":
            try: return float(d[n].replace(",", ""))
            except ValueError: pass
    return None
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    unit_t = rows[1][hdr.index("gpu__time_duration.sum")]
    t = col(d, "gpu__time_duration.sum")
    t_us = t / 1e3 if unit_t in ("ns", "nsecond") else (t if unit_t in ("us", "usecond") else t * 1e3)
    def bytes_of(name):
        v, u = col(d, name), rows[1][hdr.index(name)]
        return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    out.append({"kernel": d["Kernel Name"], "grid": d.get("launch__grid_size"), "block": d.get("launch__block_size"),
                "us": t_us, "dram_bytes": (bytes_of("dram__bytes_read.sum") or 0) + (bytes_of("dram__bytes_write.sum") or 0),
                "issue_active_pct": col(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "pipe_fp32_pct": col(d, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
                                     "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                                     "sm__inst_executed_pipe_fp32.avg.pct_of_peak_sustained_active"),
                "pipe_alu_pct": col(d, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "warps_active_pct": col(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "registers": col(d, "launch__registers_per_thread")})
print(json.dumps(out, indent=1))
