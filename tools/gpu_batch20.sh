#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --skip-cpu --steps 10 > gpurun_out/b20_bench.json 2> gpurun_out/b20_bench.err; grep bench gpurun_out/b20_bench.err | tail -12
timeout 200 python tools/time_config4.py 20 5
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b20_tests.log 2>&1; tail -4 gpurun_out/b20_tests.log
