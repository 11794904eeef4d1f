#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/prof_small.py > gpurun_out/b7_plain_small.log 2>&1 || { echo "plain small failed"; tail -5 gpurun_out/b7_plain_small.log; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pivqr_fused --launch-skip 3 -c 1 -o gpurun_out/b7_prof_pivqr_fused python tools/prof_small.py > gpurun_out/b7_ncu1.log 2>&1; tail -2 gpurun_out/b7_ncu1.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chol_inv --launch-skip 2 -c 1 -o gpurun_out/b7_prof_chol python tools/prof_small.py > gpurun_out/b7_ncu2.log 2>&1; tail -2 gpurun_out/b7_ncu2.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jacobi_kernel -c 1 -o gpurun_out/b7_prof_jacobi python tools/prof_small.py > gpurun_out/b7_ncu3.log 2>&1; tail -2 gpurun_out/b7_ncu3.log | cut -c1-200
timeout 600 bash tools/gpu_profile.sh r2b
