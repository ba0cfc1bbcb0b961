#!/bin/bash
mkdir -p gpurun_out
SY_LIB_PATH=variants/libsy_env_fclk.so timeout 200 python tools/exp_fused_clocks.py c3 > gpurun_out/r2n_fclk.log 2>&1; cat gpurun_out/r2n_fclk.log | tail -3
SY_LIB_PATH=variants/libsy_env_pclk.so timeout 200 python tools/exp_phase_clocks.py > gpurun_out/r2n_pclk.log 2>&1; cat gpurun_out/r2n_pclk.log | tail -24
