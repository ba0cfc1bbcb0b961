#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deferred.py tests/test_gpu_fullsize.py -m gpu -x -q -k "deferred or timed_path or rollout_random or partial_reset or ragged" > gpurun_out/r2o_gpu.log 2>&1; echo rc=$? >> gpurun_out/r2o_gpu.log; tail -30 gpurun_out/r2o_gpu.log
: > gpurun_out/r2o_steptime.jsonl
timeout 120 python tools/exp_step_time.py --workload c3 --tag c3_lagged >> gpurun_out/r2o_steptime.jsonl 2>gpurun_out/r2o_err.log
timeout 120 python tools/exp_step_time.py --workload c3 --opt lagged_kernel=0 --tag c3_nolag >> gpurun_out/r2o_steptime.jsonl 2>>gpurun_out/r2o_err.log
cat gpurun_out/r2o_steptime.jsonl; tail -3 gpurun_out/r2o_err.log
SY_LIB_PATH=variants/libsy_env_pclk.so timeout 200 python tools/exp_phase_clocks.py > gpurun_out/r2o_pclk.log 2>&1; cat gpurun_out/r2o_pclk.log | tail -34
