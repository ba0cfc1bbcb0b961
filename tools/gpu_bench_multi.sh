#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_n$N.json 2> gpurun_out/r02_bench_c3_n$N.err; echo "n$N rc=$?"
tail -5 gpurun_out/r02_bench_c3_n$N.err | cut -c1-300; nproc
python -c "
import json
t=open('gpurun_out/r02_bench_c3_n$N.json').read()
d=json.loads(t[t.index('{'):])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['passes_ms'], d['n_gpus'], d['clocks'])
"
