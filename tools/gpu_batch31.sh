#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/b31_multi.log 2>&1; tail -3 gpurun_out/b31_multi.log; grep "multi-gpu" gpurun_out/multi_gpu_worker.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_large_dense.py -x -q 2>&1 | tail -3
timeout 200 python tools/time_config4.py 20 4
