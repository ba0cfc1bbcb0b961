#!/bin/bash
# one GPU round trip: parity tests (per-test timeout: a hung kernel must not eat the GPU budget), smoke, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=240 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python -c "import json; d=json.load(open('gpurun_out/bench.json')); print('value', d['value'], 'ms/step', d['ms_per_step'], 'step_kernels_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"; tail -2 gpurun_out/bench.err
