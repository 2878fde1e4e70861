#!/usr/bin/env python
"""Split the SASS of the profiled kernel at BAR.SYNC / backward branches and report warp instructions
executed (and stall samples) per region -- shows how much of the kernel is outside the iteration loop.
usage: python profiles/ncu_regions.py rep.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[hi + 1:]:
    try:
        ins.append((int(r[0], 16), r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])))
    except (ValueError, IndexError):
        pass
base = ins[0][0]
tot = sum(i[2] for i in ins); tots = sum(i[3] for i in ins)
print(f"total warp instructions {tot}, samples {tots}")
start = 0; acc = 0; accs = 0
for k, (a, s, n, sm) in enumerate(ins):
    acc += n; accs += sm
    if "BAR.SYNC" in s or k == len(ins) - 1:
        print(f"[{ins[start][0]-base:#07x}..{a-base:#07x}] {k-start+1:5d} sass  {acc:12d} inst {100*acc/tot:5.1f}%  samples {100*accs/tots:5.1f}%  per-exec {n}")
        start = k + 1; acc = 0; accs = 0
