#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_side_sms.py 0 8 4 2>&1 | tee gpurun_out/b39_side.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_examples.py tests/test_gpu_edge_cases.py tests/test_gpu_operator.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --skip-cpu --skip-e2e 2> gpurun_out/b39_bench.err > gpurun_out/b39_bench.json; grep -E "configs\[1\]:" gpurun_out/b39_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/b39_bench.json').read().strip().splitlines()[-1])
print('config3', d['id_config3']['ms'], 'config5', d['id_config5']['ms'])"
