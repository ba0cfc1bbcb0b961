#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q -k "lsu or ragged or two_handles" > gpurun_out/r2f_fullsize.log 2>&1
echo "fullsize rc=$?" >> gpurun_out/r2f_fullsize.log
tail -8 gpurun_out/r2f_fullsize.log
: > gpurun_out/r2f_steptime.jsonl
timeout 120 python tools/exp_step_time.py --opt nf_fill=off --tag nofill >> gpurun_out/r2f_steptime.jsonl 2>gpurun_out/r2f_err.log
timeout 120 python tools/exp_step_time.py --tag fill >> gpurun_out/r2f_steptime.jsonl 2>>gpurun_out/r2f_err.log
for v in b12w4 b10w6 pg4k pg1k b12w4pg4k; do
  SY_LIB_PATH=variants/libsy_env_$v.so timeout 120 python tools/exp_step_time.py --tag $v >> gpurun_out/r2f_steptime.jsonl 2>>gpurun_out/r2f_err.log
done
timeout 120 python tools/exp_step_time.py --workload c4 --opt nf_fill=off --tag c4_nofill >> gpurun_out/r2f_steptime.jsonl 2>>gpurun_out/r2f_err.log
timeout 120 python tools/exp_step_time.py --workload c4 --tag c4_fill >> gpurun_out/r2f_steptime.jsonl 2>>gpurun_out/r2f_err.log
timeout 120 python tools/exp_step_time.py --workload c2 --tag c2_fill >> gpurun_out/r2f_steptime.jsonl 2>>gpurun_out/r2f_err.log
cat gpurun_out/r2f_steptime.jsonl
tail -5 gpurun_out/r2f_err.log
