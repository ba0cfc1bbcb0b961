#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_n2.json 2> gpurun_out/r02_bench_c3_n2.err; echo "n2 rc=$?" > gpurun_out/r02_n2_rc.txt
ls -la gpurun_out/r02_bench_c3_n2.json >> gpurun_out/r02_n2_rc.txt
cat gpurun_out/r02_n2_rc.txt; tail -30 gpurun_out/r02_bench_c3_n2.err | cut -c1-300
