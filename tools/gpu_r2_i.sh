#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2i_steptime.jsonl
timeout 120 python tools/exp_step_time.py --tag c3_default >> gpurun_out/r2i_steptime.jsonl 2>gpurun_out/r2i_err.log
SY_LIB_PATH=variants/libsy_env_scalar.so timeout 120 python tools/exp_step_time.py --tag c3_scalar_ones >> gpurun_out/r2i_steptime.jsonl 2>>gpurun_out/r2i_err.log
timeout 120 python tools/exp_step_time.py --opt step_kernel=fused --tag c3_fused >> gpurun_out/r2i_steptime.jsonl 2>>gpurun_out/r2i_err.log
timeout 120 python tools/exp_step_time.py --workload c2 --tag c2_auto >> gpurun_out/r2i_steptime.jsonl 2>>gpurun_out/r2i_err.log
timeout 120 python tools/exp_step_time.py --workload c2 --opt step_kernel=two_kernels --tag c2_two >> gpurun_out/r2i_steptime.jsonl 2>>gpurun_out/r2i_err.log
SY_LIB_PATH=variants/libsy_env_scalar.so timeout 120 python tools/exp_step_time.py --workload c2 --opt step_kernel=two_kernels --tag c2_two_scalar >> gpurun_out/r2i_steptime.jsonl 2>>gpurun_out/r2i_err.log
cat gpurun_out/r2i_steptime.jsonl; tail -3 gpurun_out/r2i_err.log
SY_LIB_PATH=variants/libsy_env_fclk.so timeout 200 python tools/exp_fused_clocks.py c3 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "fused or config2" > gpurun_out/r2i_gpu.log 2>&1; echo rc=$? >> gpurun_out/r2i_gpu.log; tail -3 gpurun_out/r2i_gpu.log
SY_LIB_PATH=variants/libsy_env_scalar.so timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "config3 and lsu" > gpurun_out/r2i_gpu2.log 2>&1; echo rc=$? >> gpurun_out/r2i_gpu2.log; tail -3 gpurun_out/r2i_gpu2.log
