#!/bin/bash
mkdir -p gpurun_out
which compute-sanitizer; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or golden_trace_replay or arbitrary" --timeout=500 > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log; tail -15 gpurun_out/r02_sanitizer_memcheck.log
