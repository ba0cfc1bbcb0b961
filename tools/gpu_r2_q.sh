#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2q_steptime.jsonl
for v in b8w4 b6w6 b4w4; do
SY_LIB_PATH=variants/libsy_env_$v.so timeout 120 python tools/exp_step_time.py --workload c3 --opt lagged_kernel=0 --tag $v >> gpurun_out/r2q_steptime.jsonl 2>>gpurun_out/r2q_err.log
done
cat gpurun_out/r2q_steptime.jsonl; tail -3 gpurun_out/r2q_err.log
