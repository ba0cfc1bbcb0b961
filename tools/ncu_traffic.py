#!/usr/bin/env python
"""DRAM traffic of the step kernels from an `ncu --set full` report -> profiles/step_kernel_traffic.json, stamped with the
sha256 of csrc/sy_env.cu so that bench.py can refuse a capture of other kernels.
usage: python tools/ncu_traffic.py report.ncu-rep workload [source note]"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, wl = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
col = {n: hdr.index(n) for n in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


per = {}
for r in rows[2:]:
    k = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    d = per.setdefault(k, dict(n=0, read=0.0, write=0.0, us=0.0))
    d["n"] += 1
    d["read"] += to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    d["write"] += to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    d["us"] += float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)
step = {k: v for k, v in per.items() if "sy_logic" in k or "sy_observe" in k or "sy_step_fused" in k}
total = sum((v["read"] + v["write"]) / v["n"] for v in step.values())
path = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
rec = json.load(open(path)) if os.path.isfile(path) else {}
with open(os.path.join(ROOT, "student_mechanism_design_b200", "csrc", "sy_env.cu"), "rb") as f:
    sha = hashlib.sha256(f.read()).hexdigest()
if rec.get("source_sha256") != sha:
    rec = {}  # captures of another build do not mix
rec["source_sha256"] = sha
rec[wl] = {"dram_bytes_per_launch": total,
           "what": "dram__bytes_read.sum + dram__bytes_write.sum per sy_step call, summed over its kernels (average per launch), ncu --set full, cold L2 per kernel",
           "kernels": {k: {"launches": v["n"], "read_bytes": v["read"] / v["n"], "write_bytes": v["write"] / v["n"], "duration_us": v["us"] / v["n"]} for k, v in step.items()},
           "source": note}
json.dump(rec, open(path, "w"), indent=1)
print(json.dumps(rec[wl], indent=1))
