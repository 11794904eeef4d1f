#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/b35_tests.log 2>&1; tail -4 gpurun_out/b35_tests.log
timeout 900 python bench.py > gpurun_out/b35_bench.json 2> gpurun_out/b35_bench.err; grep "\[bench\]" gpurun_out/b35_bench.err | tail -14; tail -c 400 gpurun_out/b35_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/b35_bench_ref.json 2> gpurun_out/b35_bench_ref.err; tail -c 600 gpurun_out/b35_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
CMD="python bench.py --steps 1 --warmup 1 --skip-cpu --skip-e2e --skip-sub"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b35_launches.csv $CMD > gpurun_out/b35_ncu_list.log 2>&1
tail -2 gpurun_out/b35_ncu_list.log | cut -c1-200
