#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b14_bench_n2.json 2> gpurun_out/b14_bench_n2.err; tail -c 1500 gpurun_out/b14_bench_n2.json; grep "\[bench\]\|Fatal\|File\|Segmentation" gpurun_out/b14_bench_n2.err | head -40
