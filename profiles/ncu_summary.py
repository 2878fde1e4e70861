#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS source page) into the handful of numbers we track.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__block_size", "launch__grid_size", "smsp__inst_executed.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp16.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg"]
for r in rows[2:]:
    print("=" * 100)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:80s} {r[i]:>24s} {units[i]}")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_warp_active.pct") and float(r[i] or 0) > 1.0:
            print(f"{h:80s} {r[i]:>24s}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for k, hi in enumerate(his):
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    end = his[k + 1] - 1 if k + 1 < len(his) else len(rows)
    byop, samp, tot = collections.Counter(), collections.Counter(), 0
    stall = collections.Counter()
    for r in rows[hi + 1:end]:
        try:
            n = int(r[ci["Instructions Executed"]]); s = int(r[ci["# Samples"]])
        except (ValueError, IndexError):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
        op = m.group(2).split(".")[0] if m else "?"
        byop[op] += n; samp[op] += s; tot += n
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                try: stall[h] += int(r[ci[h]])
                except ValueError: pass
    print(f"--- kernel {k}: {tot} warp instructions")
    print("  " + ", ".join(f"{op} {100*n/tot:.1f}%" for op, n in byop.most_common(24)))
    ts = sum(stall.values()) or 1
    print("  stalls: " + ", ".join(f"{h[6:]} {100*n/ts:.1f}%" for h, n in stall.most_common(10)))
