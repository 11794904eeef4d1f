#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_side_sms.py 0 8 16 2>&1 | tee gpurun_out/b37_side.log
timeout 300 python bench.py --skip-cpu --skip-e2e --skip-sub 2>&1 | grep -E "configs\[1\]:" 
