#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q -k "fused or ragged or two_handles" > gpurun_out/r2e_fullsize.log 2>&1
echo "fullsize rc=$?" >> gpurun_out/r2e_fullsize.log
tail -15 gpurun_out/r2e_fullsize.log
: > gpurun_out/r2e_steptime.jsonl
timeout 120 python tools/exp_step_time.py --opt step_kernel=fused --tag fused >> gpurun_out/r2e_steptime.jsonl 2>gpurun_out/r2e_err.log
cat gpurun_out/r2e_steptime.jsonl
SY_LIB_PATH=variants/libsy_env_fclk.so timeout 200 python tools/exp_fused_clocks.py c3 2>&1 | tail -3
tail -5 gpurun_out/r2e_err.log
