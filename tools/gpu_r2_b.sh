#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q > gpurun_out/r2c_fullsize.log 2>&1
echo "fullsize rc=$?" >> gpurun_out/r2c_fullsize.log
tail -5 gpurun_out/r2c_fullsize.log
: > gpurun_out/r2c_steptime.jsonl
timeout 120 python tools/exp_step_time.py --writer lsu --tag base_lsu >> gpurun_out/r2c_steptime.jsonl 2>gpurun_out/r2c_err.log
timeout 120 python tools/exp_step_time.py --writer bulk --tag base_bulk >> gpurun_out/r2c_steptime.jsonl 2>>gpurun_out/r2c_err.log
for v in bw2 bw3 bw8 d1 d4 bw2w4; do
  SY_LIB_PATH=variants/libsy_env_$v.so timeout 120 python tools/exp_step_time.py --writer bulk --tag $v >> gpurun_out/r2c_steptime.jsonl 2>>gpurun_out/r2c_err.log
done
timeout 120 python tools/exp_step_time.py --workload c4 --writer lsu --tag c4_lsu >> gpurun_out/r2c_steptime.jsonl 2>>gpurun_out/r2c_err.log
timeout 120 python tools/exp_step_time.py --workload c4 --writer bulk --tag c4_bulk >> gpurun_out/r2c_steptime.jsonl 2>>gpurun_out/r2c_err.log
cat gpurun_out/r2c_steptime.jsonl
tail -5 gpurun_out/r2c_err.log
