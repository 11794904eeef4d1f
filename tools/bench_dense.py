"""SURVEY 8(f) rank 2 at size: QR::compute_from and SVD::compute_from of a large dense matrix on the B200 next to
?geqp3 + ?orgqr / ?gesdd (scipy LAPACK, all host cores) on the same matrix.  Usage: tools/bench_dense.py [m n] [--no-cpu]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from rusty_compression_b200 import api
from oracle import reference_path as ref
from test_gpu_large_dense import dense_matrix

args = [x for x in sys.argv[1:] if not x.startswith("--")]
shapes = [(int(args[0]), int(args[1]))] if len(args) >= 2 else [(4096, 2048), (8192, 4096)]
cpu = "--no-cpu" not in sys.argv
ctx = api.default_context()
for shape in shapes:
    a = dense_matrix(shape, np.float64, seed=21, decades=8.0)
    d = api.DeviceMatrix.from_numpy(a)
    for name, fn in (("QR::compute_from ", lambda: api.QR.compute_from(d)), ("SVD::compute_from", lambda: api.SVD.compute_from(d))):
        fn(); ctx.synchronize()
        t0 = time.perf_counter(); out = fn(); ctx.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        line = f"{shape[0]}x{shape[1]} f64 {name}: B200 {ms:9.1f} ms"
        if cpu:
            t0 = time.perf_counter()
            o = ref.QR.compute_from(a) if name.startswith("QR") else ref.SVD.compute_from(a)
            sec = time.perf_counter() - t0
            if name.startswith("QR"):
                same = bool(np.array_equal(np.asarray(out.ind), np.asarray(o.ind)))
                dd, d0 = np.abs(np.diag(np.asarray(out.r))), np.abs(np.diag(o.r))
                line += (f" | CPU ?geqp3+?orgqr {sec * 1e3:9.1f} ms ({os.cpu_count()} cores) | pivots identical: {same}, "
                         f"max rel dev of |r_ii| {np.max(np.abs(dd - d0) / d0):.1e}")
            else:
                s, s0 = out.s_f64(), np.asarray(o.s, dtype=np.float64)
                line += f" | CPU ?gesdd {sec * 1e3:9.1f} ms ({os.cpu_count()} cores) | max rel dev of s {np.max(np.abs(s - s0) / s0):.1e}"
        print(line, flush=True)
        del out
    d.free()
