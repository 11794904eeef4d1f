#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_large_dense.py -q > gpurun_out/b6_tests.log 2>&1; tail -3 gpurun_out/b6_tests.log
timeout 200 rusty_compression_b200/build/rc_peaks shapes > gpurun_out/b6_peaks.log 2>&1; cat gpurun_out/b6_peaks.log
