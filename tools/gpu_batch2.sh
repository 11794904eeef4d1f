#!/bin/bash
# session batch: large-dense tests + timing, stage traces and launch lists of configs 3 and 5
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_large_dense.py -x -q > gpurun_out/b2_tests.log 2>&1; tail -15 gpurun_out/b2_tests.log
timeout 120 python tools/trace_config3.py 3 > gpurun_out/b2_trace3.log 2>&1
timeout 120 python tools/trace_config3.py 5 > gpurun_out/b2_trace5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b2_launches3.csv python tools/trace_config3.py 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b2_launches5.csv python tools/trace_config3.py 5 > /dev/null 2>&1
timeout 500 python tools/bench_dense.py > gpurun_out/b2_dense.log 2>&1; cat gpurun_out/b2_dense.log | tail
