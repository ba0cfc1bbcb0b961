#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_gpu.log 2>&1; echo rc=$? >> gpurun_out/r2h_gpu.log; tail -4 gpurun_out/r2h_gpu.log
: > gpurun_out/r2h_steptime.jsonl
timeout 120 python tools/exp_step_time.py --tag c3_default >> gpurun_out/r2h_steptime.jsonl 2>gpurun_out/r2h_err.log
timeout 120 python tools/exp_step_time.py --workload c2 --tag c2_auto >> gpurun_out/r2h_steptime.jsonl 2>>gpurun_out/r2h_err.log
timeout 120 python tools/exp_step_time.py --workload c2 --opt step_kernel=fused --tag c2_fused >> gpurun_out/r2h_steptime.jsonl 2>>gpurun_out/r2h_err.log
timeout 120 python tools/exp_step_time.py --workload c2 --opt step_kernel=two_kernels --tag c2_two >> gpurun_out/r2h_steptime.jsonl 2>>gpurun_out/r2h_err.log
timeout 120 python tools/exp_step_time.py --workload c4 --tag c4_default >> gpurun_out/r2h_steptime.jsonl 2>>gpurun_out/r2h_err.log
cat gpurun_out/r2h_steptime.jsonl; tail -3 gpurun_out/r2h_err.log
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2h_c2.json 2>gpurun_out/r2h_c2.err; python -c "import json; d=json.load(open('gpurun_out/r2h_c2.json')); print('c2 bench', d['value'], d['ms_per_step'], d['config']['loop'])"
