#!/usr/bin/env python
"""Static SASS size of a kernel between barriers (is the iteration loop inside the 32 KB instruction cache?).
usage: python tools/sass_regions.py <object file> <kernel-name substring>"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, ker = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); ker[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        ker[cur].append(line.split("*/", 1)[1].split(";")[0].strip())
for name, ins in ker.items():
    if sys.argv[2] not in name:
        continue
    print(name, len(ins), "instructions,", len(ins) * 16 // 1024, "KB")
    start = 0
    for k, s in enumerate(ins):
        if "BAR.SYNC" in s or "BAR.RED" in s or k == len(ins) - 1:
            print(f"  [{start:6d}..{k:6d}] {k - start + 1:6d} instr {16 * (k - start + 1) / 1024:6.1f} KB   ends with {s[:40]}")
            start = k + 1
