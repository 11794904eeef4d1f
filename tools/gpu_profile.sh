#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the ncu launch list, then one
# --set full capture of the dominant kernels.  Usage: tools/gpu_profile.sh <tag>
TAG=${1:-rX}
CMD="python bench.py --steps 1 --warmup 1 --skip-cpu --skip-e2e --skip-sub"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
tail -2 gpurun_out/ncu_list_$TAG.log | cut -c1-300
# the dominant kernels alone (Y = A Omega and Z = A^T Y at the config-2 shape): launch 0 = NN, launch 1 = TN
python tools/prof_gemm.py > gpurun_out/plain_gemm_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:dmma_gemm_kernel -c 2 \
    -o gpurun_out/prof_dmma_$TAG python tools/prof_gemm.py > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-300
