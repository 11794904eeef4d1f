#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b3_tests.log 2>&1; tail -15 gpurun_out/b3_tests.log
timeout 300 python tools/bench_dense.py 4096 2048 > gpurun_out/b3_dense.log 2>&1; tail -3 gpurun_out/b3_dense.log
timeout 300 python bench.py --skip-cpu --skip-e2e --steps 5 > gpurun_out/b3_bench.json 2> gpurun_out/b3_bench.err; tail -c 1500 gpurun_out/b3_bench.json
timeout 120 python tools/trace_config3.py 3 > gpurun_out/b3_trace3.log 2>&1
timeout 120 python tools/trace_config3.py 5 > gpurun_out/b3_trace5.log 2>&1
