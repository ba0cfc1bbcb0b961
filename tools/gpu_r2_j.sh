#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/$name.json')); print('value %.4g e2e %.4g (halves %s, %s) refshape %.4g' % (d['value'], d['e2e']['value'], d['e2e'].get('half_batches'), d['e2e'].get('loop'), d['e2e_reference_shape']['value']))" 2>&1 | tail -1)"; tail -2 gpurun_out/$name.err; }
run r2j_c3_h1 --steps 100 --no-cpu-baseline --e2e-halves 1
run r2j_c3_h2 --steps 100 --no-cpu-baseline --e2e-halves 2
run r2j_c3_h4 --steps 100 --no-cpu-baseline --e2e-halves 4
run r2j_c3_h8 --steps 100 --no-cpu-baseline --e2e-halves 8
run r2j_c3_h4py --steps 100 --no-cpu-baseline --e2e-halves 4 --e2e-python-loop
