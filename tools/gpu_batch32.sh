#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q > gpurun_out/b32_tests.log 2>&1; tail -15 gpurun_out/b32_tests.log
timeout 300 python tools/time_steep.py 2>&1 | tee gpurun_out/b32_steep.log
timeout 600 python bench.py --skip-cpu --skip-sub > gpurun_out/b32_bench.json 2> gpurun_out/b32_bench.err; tail -5 gpurun_out/b32_bench.err; cat gpurun_out/b32_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'])"
