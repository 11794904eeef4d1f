#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/b34_multi.log 2>&1; tail -3 gpurun_out/b34_multi.log; grep "multi-gpu" gpurun_out/multi_gpu_worker.log | cut -c1-330
