#!/bin/bash
# one GPU round trip: parity tests, bench line, per-kernel device times (ncu launch list).  Run under gpurun.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 300 --warmup 20 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python -c "import json; d=json.load(open('gpurun_out/bench.json')); print('value', d['value'], 'ms/step', d['ms_per_step'], 'step_kernels_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --e2e-steps 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sy_" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 4 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu.log 2>&1
python tools/ncu_launches.py gpurun_out/launches.csv
