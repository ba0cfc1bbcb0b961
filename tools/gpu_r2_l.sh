#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_gpu.log 2>&1; echo rc=$? >> gpurun_out/r2l_gpu.log; tail -4 gpurun_out/r2l_gpu.log
: > gpurun_out/r2l_steptime.jsonl
timeout 120 python tools/exp_step_time.py --workload c4 --tag c4_lsu >> gpurun_out/r2l_steptime.jsonl 2>gpurun_out/r2l_err.log
timeout 120 python tools/exp_step_time.py --workload c4 --writer bulk --tag c4_bulk >> gpurun_out/r2l_steptime.jsonl 2>>gpurun_out/r2l_err.log
timeout 120 python tools/exp_step_time.py --workload c3 --tag c3 >> gpurun_out/r2l_steptime.jsonl 2>>gpurun_out/r2l_err.log
cat gpurun_out/r2l_steptime.jsonl; tail -3 gpurun_out/r2l_err.log
