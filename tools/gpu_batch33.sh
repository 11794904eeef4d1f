#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/b33_multi.log 2>&1; tail -3 gpurun_out/b33_multi.log; grep "multi-gpu" gpurun_out/multi_gpu_worker.log | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --skip-cpu --skip-sub > gpurun_out/b33_bench_n2.json 2> gpurun_out/b33_bench_n2.err; tail -4 gpurun_out/b33_bench_n2.err; python -c "import json; d=json.loads(open('gpurun_out/b33_bench_n2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['one_step_at_a_time'])"
