#!/bin/bash
# re-entry check: GPU parity suite + smoke + step-time variants + default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_gpu.log 2>&1; echo rc=$? >> gpurun_out/r2m_gpu.log; tail -4 gpurun_out/r2m_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; tail -1 gpurun_out/r2m_smoke.log
: > gpurun_out/r2m_steptime.jsonl
for w in c3 c4 c2; do
timeout 120 python tools/exp_step_time.py --workload $w --tag ${w}_default >> gpurun_out/r2m_steptime.jsonl 2>>gpurun_out/r2m_err.log
timeout 120 python tools/exp_step_time.py --workload $w --writer bulk --tag ${w}_bulk >> gpurun_out/r2m_steptime.jsonl 2>>gpurun_out/r2m_err.log
timeout 120 python tools/exp_step_time.py --workload $w --opt step_kernel=fused --tag ${w}_fused >> gpurun_out/r2m_steptime.jsonl 2>>gpurun_out/r2m_err.log
timeout 120 python tools/exp_step_time.py --workload $w --opt nf_fill=on --tag ${w}_split >> gpurun_out/r2m_steptime.jsonl 2>>gpurun_out/r2m_err.log
done
cat gpurun_out/r2m_steptime.jsonl; tail -3 gpurun_out/r2m_err.log
timeout 600 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo bench rc=$?
