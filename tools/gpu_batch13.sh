#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_config4.py 20 > gpurun_out/b13_trace4.log 2>&1; grep -v "^$" gpurun_out/b13_trace4.log | tail -12
timeout 120 python tools/trace_config3.py 5 > gpurun_out/b13_trace5.log 2>&1; grep -v "^$" gpurun_out/b13_trace5.log | tail -22 | head -10
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b13_tests.log 2>&1; tail -4 gpurun_out/b13_tests.log
