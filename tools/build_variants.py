#!/usr/bin/env python
"""Build kernel variants of libsy_env.so side by side under variants/ (git-ignored; they travel to the GPU box) so one
gpurun call can time several builds: `python tools/build_variants.py name:-DFLAG=1,-DOTHER=2 ...`.
tools/exp_step_time.py loads each through SY_LIB_PATH (experiments only; bench.py refuses an overridden library)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "student_mechanism_design_b200", "csrc", "sy_env.cu")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared", "-Xcompiler", "-fPIC"]


def main():
    os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
    procs = []
    for spec in sys.argv[1:]:
        name, _, defs = spec.partition(":")
        out = os.path.join(ROOT, "variants", f"libsy_env_{name}.so")
        cmd = ["nvcc", *FLAGS, *[d for d in defs.split(",") if d], "-o", out, SRC]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        if p.wait() != 0:
            raise SystemExit(f"variant {name} failed to build")
        print("built", name)


if __name__ == "__main__":
    main()
