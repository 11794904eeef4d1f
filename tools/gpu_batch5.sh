#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/b5_tests.log 2>&1; tail -5 gpurun_out/b5_tests.log
timeout 300 python tools/trace_config4.py 20 > gpurun_out/b5_trace4.log 2>&1; grep -c trace gpurun_out/b5_trace4.log
