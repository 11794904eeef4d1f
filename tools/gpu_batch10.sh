#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "tall_pivoted or medium_pivoted or tcgen05" > gpurun_out/b10_tests.log 2>&1; tail -5 gpurun_out/b10_tests.log
timeout 300 python tools/trace_config4.py 20 > gpurun_out/b10_trace4.log 2>&1; grep -v "^$" gpurun_out/b10_trace4.log | tail -16
