"""Wall time per pass of the config-4 per-GPU share (sample_range_by_rank on a 2^k x 8192 f32 tall shard), several passes
in a row: shows allocator warm-up / steady state.  Usage: tools/time_config4.py [log2 rows] [passes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api
ctx = api.default_context()
m = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 6
a = api.tall_shard_matrix(0, m, 8192, np.float32, 9, m, r0=512, decade_every=64.0)
ctx.synchronize()
for i in range(passes):
    t0 = time.perf_counter()
    q = api.sample_range_by_rank(a, 256, 10, seed=42, device=True)
    ctx.synchronize()
    print(f"pass {i}: {(time.perf_counter() - t0) * 1e3:9.2f} ms", flush=True)
    del q
