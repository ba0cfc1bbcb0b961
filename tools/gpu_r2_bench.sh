#!/bin/bash
# round 2: the driver-style bench lines (+ reference arm), the other workloads, an ncu launch list and a full capture
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/$name.json')); print('value %.4g ms/step %.4f frac %s e2e %.4g' % (d['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'), d['e2e']['value']))" 2>&1 | tail -1)"; }
run r02_bench_c3_steps20 --steps 20 --warmup 5
run r02_bench_c3_n1
run r02_bench_c3_reference_arm --impl reference --steps 20 --warmup 5
run r02_bench_c2_n1 --workload c2
run r02_bench_c4_n1 --workload c4 --steps 300
run r02_bench_c5_gnn_n1 --workload c5 --steps 300
run r02_bench_c5_mappo_n1 --workload c5 --policy mappo --steps 200
timeout 300 python bench.py --steps 10 --warmup 4 --passes 1 --no-cpu-baseline --e2e-steps 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sy_" -c 200 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 10 --warmup 4 --passes 1 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu.log 2>&1
python tools/ncu_launches.py gpurun_out/r02_launches_c3.csv
python tools/exp_two_steps.py c3 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sy_observe|sy_logic" -s 8 -c 2 -o gpurun_out/r02_step_c3 -f python tools/exp_two_steps.py c3 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
python tools/exp_two_steps.py c4 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sy_observe|sy_logic" -s 8 -c 2 -o gpurun_out/r02_step_c4 -f python tools/exp_two_steps.py c4 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
