#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q -k "lsu or ragged or two_handles" > gpurun_out/r2g_fullsize.log 2>&1
echo "fullsize rc=$?" >> gpurun_out/r2g_fullsize.log
tail -8 gpurun_out/r2g_fullsize.log
: > gpurun_out/r2g_steptime.jsonl
timeout 120 python tools/exp_step_time.py --opt nf_fill=off --tag nofill >> gpurun_out/r2g_steptime.jsonl 2>gpurun_out/r2g_err.log
timeout 120 python tools/exp_step_time.py --tag split >> gpurun_out/r2g_steptime.jsonl 2>>gpurun_out/r2g_err.log
for v in pg4k pg16k; do
  SY_LIB_PATH=variants/libsy_env_$v.so timeout 120 python tools/exp_step_time.py --tag $v >> gpurun_out/r2g_steptime.jsonl 2>>gpurun_out/r2g_err.log
done
timeout 120 python tools/exp_step_time.py --workload c5 --tag c5_split >> gpurun_out/r2g_steptime.jsonl 2>>gpurun_out/r2g_err.log
timeout 120 python tools/exp_step_time.py --workload c5 --opt nf_fill=off --tag c5_nofill >> gpurun_out/r2g_steptime.jsonl 2>>gpurun_out/r2g_err.log
cat gpurun_out/r2g_steptime.jsonl
tail -5 gpurun_out/r2g_err.log
python tools/exp_two_steps.py c3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 15 --csv --log-file gpurun_out/r2_launch_split.csv python tools/exp_two_steps.py c3 > /dev/null 2>&1
