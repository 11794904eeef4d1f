#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/b12_multi.log 2>&1; tail -4 gpurun_out/b12_multi.log; tail -8 gpurun_out/multi_gpu_worker.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b12_bench_n2.json 2> gpurun_out/b12_bench_n2.err; tail -c 2500 gpurun_out/b12_bench_n2.json
