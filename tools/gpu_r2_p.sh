#!/bin/bash
mkdir -p gpurun_out
SY_LIB_PATH=variants/libsy_env_fclk.so timeout 200 python tools/exp_lagged_clocks.py c3 > gpurun_out/r2p_lag.log 2>&1; tail -2 gpurun_out/r2p_lag.log
SY_LIB_PATH=variants/libsy_env_fclk.so timeout 200 python tools/exp_lagged_clocks.py c3 plain > gpurun_out/r2p_plain.log 2>&1; tail -2 gpurun_out/r2p_plain.log
