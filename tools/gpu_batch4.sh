#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b4_tests.log 2>&1; tail -5 gpurun_out/b4_tests.log
RC_TF32_RING=2 timeout 300 python tools/check_gemm_f32.py > gpurun_out/b4_check_ring2.log 2>&1; tail -4 gpurun_out/b4_check_ring2.log
RC_TF32_RING=1 timeout 300 python tools/check_gemm_f32.py > gpurun_out/b4_check_ring1.log 2>&1; tail -2 gpurun_out/b4_check_ring1.log
RC_SKIP_SIMT=1 timeout 300 python tools/bench_gemm_f32.py > gpurun_out/b4_gemm_f32.log 2>&1; cat gpurun_out/b4_gemm_f32.log
