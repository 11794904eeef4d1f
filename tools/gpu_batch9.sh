#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_config4.py 20 > gpurun_out/b9_trace4.log 2>&1; grep -v "^$" gpurun_out/b9_trace4.log | tail -16
timeout 300 python tools/check_gemm_f32.py > gpurun_out/b9_check.log 2>&1; tail -2 gpurun_out/b9_check.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b9_tests.log 2>&1; tail -4 gpurun_out/b9_tests.log
