#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_side_sms.py 0 4 8 16 2>&1 | tee gpurun_out/b36_side.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_examples.py -x -q 2>&1 | tail -4
