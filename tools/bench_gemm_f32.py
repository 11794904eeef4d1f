"""Times the tcgen05 contractions Y = A X and Z = A^T Y (f32) at the config-3 / config-4 shapes: the default 3xTF32
split, the opt-in bf16 single product (f32_precision = 1) and the SIMT tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rusty_compression_b200 import api
ctx = api.default_context()
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
shapes = [(32768, 32768, 64), (65536, 8192, 266), (65536, 8192, 74)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]]
for (m, n, l) in shapes:
    a = api.DeviceMatrix.random_gaussian((m, n), np.float32, 1)
    x = api.DeviceMatrix.random_gaussian((n, l), np.float32, 2)
    for impl in ((0, 3, 4, 2) if os.environ.get('RC_SKIP_SIMT') else (0, 3, 4, 2, 1)):
        ctx.set_option("gemm_impl", 1 if impl == 1 else 0)
        ctx.set_option("f32_precision", 1 if impl == 2 else 0)
        ctx.set_option("tf32_ring", {3: 1, 4: 2}.get(impl, 0))
        for _ in range(2): y = a.matmat(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        reps = 5
        for _ in range(reps): y = a.matmat(x)
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        yy = y
        for _ in range(2): z = a.conj_matmat(yy)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps): z = a.conj_matmat(yy)
        e1.record(stream); torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / reps
        print(f"   A^T Y: {ms_t:.3f} ms  {2*m*n*l/ms_t/1e9:.1f} TFLOP/s  {m*n*4/ms_t/1e6:.0f} GB/s", flush=True)
        print(f"{m}x{n}x{l} f32 impl={ {0: 'tcgen05-tf32x3', 3: 'tcgen05-tf32x3 deep split ring', 4: 'tcgen05-tf32x3 hi from smem (SS) + deep ring', 2: 'tcgen05-bf16', 1: 'simt'}[impl] }: {ms:.3f} ms  {2*m*n*l/ms/1e9:.1f} TFLOP/s  {m*n*4/ms/1e6:.0f} GB/s", flush=True)
    ctx.set_option("gemm_impl", 0); ctx.set_option("f32_precision", 0); ctx.set_option("tf32_ring", 0)
    del a, x, y
