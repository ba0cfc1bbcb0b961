#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_gpu.log 2>&1; echo rc=$? >> gpurun_out/r2k_gpu.log; tail -5 gpurun_out/r2k_gpu.log
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/$name.json')); print('value %.4g ms/step %.4f policy_ms %.4f e2e %.4g' % (d['value'], d['ms_per_step'], d['config']['policy_ms_per_step'], d['e2e']['value']))" 2>&1 | tail -1)"; tail -2 gpurun_out/$name.err; }
run r2k_c5_gnn --workload c5 --steps 200 --no-cpu-baseline
run r2k_c5_mappo --workload c5 --policy mappo --steps 200 --no-cpu-baseline
