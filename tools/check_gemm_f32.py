"""f32 / c32 tcgen05 GEMM against numpy (f64) over a sweep of shapes; prints max relative error."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api

rng = np.random.default_rng(0)
if os.environ.get('RC_TF32_RING'):
    api.default_context().set_option('tf32_ring', int(os.environ['RC_TF32_RING']))
shapes = [(4096, 128, 2048), (4096, 128, 128), (4096, 64, 128), (4096, 128, 256), (4096, 256, 128), (4096, 2048, 128),
          (4096, 1024, 74), (4096, 1024, 148), (4096, 1024, 266), (8192, 4096, 320), (1000, 333, 130), (4096, 64, 1024),
          (4096, 96, 2048), (512, 128, 2048), (4096, 128, 384)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
bad = 0
for dtype in (np.float32, np.complex64):
    for (m, k, n) in shapes:
        a = rng.standard_normal((m, k)).astype(dtype)
        x = rng.standard_normal((k, n)).astype(dtype)
        if np.dtype(dtype).kind == "c":
            a = a + 1j * rng.standard_normal((m, k)).astype(np.float32)
            x = x + 1j * rng.standard_normal((k, n)).astype(np.float32)
        ad = api.DeviceMatrix.from_numpy(a)
        y = ad.matmat(x).to_numpy()
        ref = a.astype(np.complex128).dot(x.astype(np.complex128))
        e1 = np.max(np.abs(y - ref)) / np.max(np.abs(ref))
        w = rng.standard_normal((m, n)).astype(dtype)
        z = ad.conj_matmat(w).to_numpy()
        refz = np.conj(a.T).astype(np.complex128).dot(w.astype(np.complex128))
        e2 = np.max(np.abs(z - refz)) / np.max(np.abs(refz))
        flag = "" if (e1 < 1e-5 and e2 < 1e-5) else "   <-- BAD"
        bad += bool(flag)
        print(f"{np.dtype(dtype).name:10s} A {m}x{k}  X {k}x{n}:  A X err {e1:.2e}   A^H W ({k}x{n}) err {e2:.2e}{flag}", flush=True)
print("BAD" if bad else "ALL OK")
