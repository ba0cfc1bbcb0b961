#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q > gpurun_out/r2d_fullsize.log 2>&1
echo "fullsize rc=$?" >> gpurun_out/r2d_fullsize.log
tail -15 gpurun_out/r2d_fullsize.log
: > gpurun_out/r2d_steptime.jsonl
timeout 120 python tools/exp_step_time.py --opt step_kernel=two_kernels --tag two_kernels_lsu >> gpurun_out/r2d_steptime.jsonl 2>gpurun_out/r2d_err.log
timeout 120 python tools/exp_step_time.py --opt step_kernel=fused --tag fused >> gpurun_out/r2d_steptime.jsonl 2>>gpurun_out/r2d_err.log
timeout 120 python tools/exp_step_time.py --workload c2 --opt step_kernel=two_kernels --tag c2_two >> gpurun_out/r2d_steptime.jsonl 2>>gpurun_out/r2d_err.log
timeout 120 python tools/exp_step_time.py --workload c2 --opt step_kernel=fused --tag c2_fused >> gpurun_out/r2d_steptime.jsonl 2>>gpurun_out/r2d_err.log
cat gpurun_out/r2d_steptime.jsonl
tail -5 gpurun_out/r2d_err.log
