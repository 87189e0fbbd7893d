"""Summarise an ncu report per kernel: key metrics + hottest source lines.
usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp32.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "local_load_requests", "smsp__inst_executed_op_local_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    for k in KEYS:
        if k in d and d[k] != "":
            print(f"  {k} = {d[k]} {rows[1][hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur_file = None; hdr = None; kern = None; data = {}
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "Kernel Name": kern = r[1][:60] + f"#{len(data)}"; data[kern] = []; continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or not hdr or r[0] == "": continue
    try: ln = int(r[0])
    except ValueError: continue
    k = len(hdr) - 4
    d = dict(zip(hdr[4:], r[-k:]))
    try:
        s = float(d["Warp Stall Sampling (All Samples)"] or 0); bar = float(d.get("stall_barrier", 0) or 0)
        inst = float(d["Instructions Executed"] or 0)
    except ValueError: continue
    top_st = sorted(((float(v or 0), kk) for kk, v in d.items() if kk.startswith("stall_") and "Not Issued" not in kk), reverse=True)[:2]
    data.setdefault(kern or "?", []).append((s, bar, inst, cur_file, ln, ",".join(r[1:-k - 2]).strip()[:90], top_st))
for kern, rows_ in data.items():
    tot = sum(x[0] for x in rows_) or 1
    print("\n####", kern, "samples", tot, "warp-inst", sum(x[2] for x in rows_))
    for s, bar, inst, f, ln, txt, st in sorted(rows_, reverse=True)[:top]:
        print(f"{100*s/tot:5.1f}% bar={100*bar/tot:4.1f}% inst={inst:9.0f} {f}:{ln}: {txt}   {[(k, int(v)) for v, k in st]}")
