#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "medium_pivoted or tall_pivoted or wide_pivoted" > gpurun_out/b8_tests.log 2>&1; tail -5 gpurun_out/b8_tests.log
timeout 120 python tools/trace_config3.py 5 > gpurun_out/b8_trace5.log 2>&1; grep "pivqr\|pivoted QR of R" gpurun_out/b8_trace5.log
timeout 300 python tools/trace_config4.py 20 > gpurun_out/b8_trace4.log 2>&1; grep -v "^$" gpurun_out/b8_trace4.log | tail -12
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b8_launches4.csv python tools/trace_config4.py 20 > /dev/null 2>&1
