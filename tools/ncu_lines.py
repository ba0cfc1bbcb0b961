#!/usr/bin/env python
"""Per-source-line stall summary of an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [top_n] [kernel_substring]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kfilter = sys.argv[3] if len(sys.argv) > 3 else ""
by_inst = len(sys.argv) > 4 and sys.argv[4] == "inst"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, lines, seen_blocks = None, {}, 0
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        seen_blocks += 1
        continue
    if r and r[0] == "Function Name":
        fn = r[1]
    if hdr is None or len(r) != len(hdr) or not r[0].strip().isdigit():
        continue
    if kfilter not in fn:
        continue
    key = (fn, int(r[0]))
    if key in lines:
        continue  # only the first profiled instance
    lines[key] = r
si = hdr.index("# Samples")
cols = ["stall_barrier", "stall_long_sb", "stall_lg", "stall_mio", "stall_short_sb", "stall_wait", "stall_math",
        "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_no_inst"]
ci = [hdr.index(c) for c in cols]
ie = hdr.index("Instructions Executed")
tot = sum(float(r[si] or 0) for r in lines.values()) or 1.0
print(f"total samples {tot:.0f} over {len(lines)} source lines")
tot_by = {c: sum(float(r[i] or 0) for r in lines.values()) for c, i in zip(cols, ci)}
print("by reason:", {c: f"{v / tot * 100:.1f}%" for c, v in tot_by.items() if v / tot > 0.01})
keyf = (lambda kv: -float(kv[1][ie] or 0)) if by_inst else (lambda kv: -float(kv[1][si] or 0))
print('total instructions', sum(float(r[ie] or 0) for r in lines.values()))
for (fn, ln), r in sorted(lines.items(), key=keyf)[:top_n]:
    why = {c.replace("stall_", ""): int(float(r[i] or 0)) for c, i in zip(cols, ci) if float(r[i] or 0) / tot > 0.003}
    print(f"{ln:5d} {float(r[si]) / tot * 100:5.1f}%  inst={r[ie]:>8}  {r[1].strip()[:90]:90s} {why}")
