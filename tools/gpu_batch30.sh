#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/b30_multi.log 2>&1; tail -3 gpurun_out/b30_multi.log; grep "multi-gpu" gpurun_out/multi_gpu_worker.log | cut -c1-300; tail -5 gpurun_out/multi_gpu_worker.log | cut -c1-300
