#!/bin/bash
# usage: tools/gpu_bench_n.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 600 gpurun_out/bench_n$N.json; grep "\[bench\]\|Fatal\|Error" gpurun_out/bench_n$N.err | tail -12
