#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deferred.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout=240 > gpurun_out/r2z_gpu.log 2>&1; echo rc=$? >> gpurun_out/r2z_gpu.log; tail -8 gpurun_out/r2z_gpu.log
timeout 300 python bench.py --workload c2 > gpurun_out/r02_bench_c2_n1.json 2> gpurun_out/r02_bench_c2_n1.err; python -c "import json; d=json.load(open('gpurun_out/r02_bench_c2_n1.json')); print('c2 value', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
