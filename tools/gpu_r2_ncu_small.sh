#!/bin/bash
# ncu evidence for the single-launch kernels of latency-bound batches (c2) and the c4 launch list
mkdir -p gpurun_out
timeout 300 python bench.py --workload c2 --steps 10 --warmup 4 --passes 1 --no-cpu-baseline --e2e-steps 3 > gpurun_out/plain_c2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sy_" -c 200 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --workload c2 --steps 10 --warmup 4 --passes 1 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_c2.log 2>&1
python tools/ncu_launches.py gpurun_out/r02_launches_c2.csv
timeout 300 python bench.py --workload c4 --steps 10 --warmup 4 --passes 1 --no-cpu-baseline --e2e-steps 3 > gpurun_out/plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sy_" -c 200 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --workload c4 --steps 10 --warmup 4 --passes 1 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_c4.log 2>&1
python tools/ncu_launches.py gpurun_out/r02_launches_c4.csv
ncu --set full --clock-control none --import-source on -k regex:"sy_rollout_lagged" -c 1 -o gpurun_out/r02_rollout_c2 -f python bench.py --workload c2 --steps 10 --warmup 4 --passes 1 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_c2_full.log 2>&1
tail -2 gpurun_out/ncu_c2_full.log
