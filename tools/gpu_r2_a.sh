#!/bin/bash
# round 2, call A: full-size parity (both writer paths) + step-time comparison of the writer variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q > gpurun_out/r2a_fullsize.log 2>&1
echo "fullsize rc=$?" >> gpurun_out/r2a_fullsize.log
tail -5 gpurun_out/r2a_fullsize.log
: > gpurun_out/r2a_steptime.jsonl
timeout 120 python tools/exp_step_time.py --writer lsu --tag base_lsu >> gpurun_out/r2a_steptime.jsonl 2>gpurun_out/r2a_err.log
timeout 120 python tools/exp_step_time.py --writer bulk --tag base_bulk >> gpurun_out/r2a_steptime.jsonl 2>>gpurun_out/r2a_err.log
for v in nbuf2 hint1 w4 w4n2 w2n2 img44k; do
  SY_LIB_PATH=variants/libsy_env_$v.so timeout 120 python tools/exp_step_time.py --writer bulk --tag $v >> gpurun_out/r2a_steptime.jsonl 2>>gpurun_out/r2a_err.log
done
cat gpurun_out/r2a_steptime.jsonl
tail -5 gpurun_out/r2a_err.log
