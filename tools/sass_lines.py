#!/usr/bin/env python
"""SASS instruction count per CUDA source line of one kernel (code-size / I-cache budget check).
usage: python tools/sass_lines.py lib.so kernel_substring [top_n]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

so, pat = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, check=True, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout
fn, cur, cnt, files = None, None, collections.Counter(), {}
for line in txt.splitlines():
    m = re.search(r"\.section\s+\.text\.(\S+?),", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', line)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if fn and pat in fn and re.search(r"/\*[0-9a-f]{4,}\*/\s+[A-Z@]", line):
        cnt[cur] += 1
tot = sum(cnt.values())
print(f"{pat}: {tot} SASS instructions = {tot * 16 / 1024:.1f} KB")
for (f, ln), c in cnt.most_common(top_n):
    if f not in files:
        try:
            files[f] = open(f).read().splitlines()
        except OSError:
            files[f] = []
    src = files[f][ln - 1].strip()[:100] if ln - 1 < len(files[f]) else ""
    print(f"{c:5d}  {os.path.basename(f)}:{ln:<5d} {src}")
