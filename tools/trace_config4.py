import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api
ctx = api.default_context()
m = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
a = api.tall_shard_matrix(0, m, 8192, np.float32, 9, m, r0=512, decade_every=64.0)
for i in range(2):
    q = api.sample_range_by_rank(a, 256, 10, seed=42, device=True)
ctx.set_option("trace", 1)
q = api.sample_range_by_rank(a, 256, 10, seed=42, device=True)
