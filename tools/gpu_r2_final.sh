#!/bin/bash
# round 2, final build: GPU parity suite + smoke, then tools/gpu_r2_bench.sh (bench lines of every workload, ncu launch list, full captures)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=240 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
bash tools/gpu_r2_bench.sh
