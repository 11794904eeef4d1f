import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api
rng = np.random.default_rng(0)
m, k, n = (int(v) for v in sys.argv[1].split("x"))
a = rng.standard_normal((m, k)).astype(np.float32)
x = rng.standard_normal((k, n)).astype(np.float32)
ad = api.DeviceMatrix.from_numpy(a)
for rep in range(2):
    y = ad.matmat(x).to_numpy()
    ref = a.astype(np.float64).dot(x.astype(np.float64))
    err = np.abs(y - ref) / np.max(np.abs(ref))
    nch = (n + 127) // 128
    npad = ((n + nch - 1) // nch + 31) // 32 * 32
    mt = (m + 127) // 128
    grid = np.zeros((mt, nch))
    for i in range(mt):
        for c in range(nch):
            blk = err[i * 128:(i + 1) * 128, c * npad:(c + 1) * npad]
            grid[i, c] = blk.max() if blk.size else 0
    bad = np.argwhere(grid > 1e-4)
    print(f"rep {rep}: {len(bad)} bad (m-tile, chunk) items of {mt * nch}; npad {npad}")
    items = sorted(int(i * nch + c) for i, c in bad)
    print("  bad item ids t (t = mt * nchunks + ch):", items[:60])
    print("  t mod 148:", sorted(set(t % 148 for t in items))[:60])
    print("  t div 148:", sorted(set(t // 148 for t in items)))
    if len(bad):
        i, c = bad[0]
        blk = err[i * 128:(i + 1) * 128, c * npad:(c + 1) * npad]
        rows = np.nonzero(blk.max(axis=1) > 1e-4)[0]; cols = np.nonzero(blk.max(axis=0) > 1e-4)[0]
        print("  first bad item rows", rows[:10], "n", len(rows), "cols", cols[:10], "n", len(cols))
        good = np.nonzero(blk.max(axis=1) <= 1e-4)[0]
        print("  good rows of that item:", good.tolist())
        for (i2, c2) in bad[:6]:
            blk = err[i2 * 128:(i2 + 1) * 128, c2 * npad:(c2 + 1) * npad]
            print("   item", int(i2 * nch + c2), "bad rows", int((blk.max(axis=1) > 1e-4).sum()), "max err", float(blk.max()), "median err", float(np.median(blk)))
