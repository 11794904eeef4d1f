"""Philox4x32-10 + Box-Muller Gaussian generator (numpy), bit-compatible with the
device generator in ``rusty_compression_b200/csrc/philox.cuh``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Replaces, for seeded/sharded runs, the reference's ``rand_distr::Normal`` ziggurat
draw (src/random_matrix.rs:120-145).  Semantics kept from the reference:
* row-major fill order (``map_inplace`` on a C-order array, :122-124);
* all normals are drawn in f64 and then cast (:123, :140-141);
* complex entries are N(0,1) + i N(0,1), *not* variance-normalised (:139-143).

Counter layout (so that any shard can regenerate any element):
    counter = (lo32(e), hi32(e), stream, 0),  key = (lo32(seed), hi32(seed))
with ``e`` the row-major element index.  The four output words give
    u1 = ((w0 >> 5) * 2^26 + (w1 >> 6) + 1) * 2^-53   in (0, 1]
    u2 = ((w2 >> 5) * 2^26 + (w3 >> 6))     * 2^-53   in [0, 1)
    r = sqrt(-2 ln u1);  re = r cos(2 pi u2);  im = r sin(2 pi u2)
Real matrices use ``re`` only; complex matrices use (re, im).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32-valued arrays
    (held as uint64 for the 32x32->64 multiplies).  Returns four uint64 arrays
    holding 32-bit words."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _uniform_pair(e, seed, stream):
    e = np.asarray(e, dtype=np.uint64)
    w0, w1, w2, w3 = philox4x32_10(e & MASK32, e >> np.uint64(32),
                                   np.uint64(stream & 0xFFFFFFFF), np.uint64(0),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    two26 = np.float64(67108864.0)
    two_m53 = np.float64(2.0 ** -53)
    u1 = ((w0 >> np.uint64(5)).astype(np.float64) * two26
          + (w1 >> np.uint64(6)).astype(np.float64) + 1.0) * two_m53
    u2 = ((w2 >> np.uint64(5)).astype(np.float64) * two26
          + (w3 >> np.uint64(6)).astype(np.float64)) * two_m53
    return u1, u2


def gaussian_pair(e, seed, stream=0):
    """Two independent N(0,1) f64 samples for every element index in ``e``."""
    u1, u2 = _uniform_pair(e, seed, stream)
    r = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    return r * np.cos(ang), r * np.sin(ang)


_DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128,
       np.dtype(np.float32): np.float32, np.dtype(np.float64): np.float64,
       np.dtype(np.complex64): np.complex64, np.dtype(np.complex128): np.complex128}


def random_gaussian(shape, dtype, seed, stream=0, row_offset=0, chunk=1 << 22):
    """Seeded equivalent of ``RandomMatrix::random_gaussian``
    (src/random_matrix.rs:21, 96-145).  ``row_offset`` lets a row shard
    regenerate its slice of a larger matrix (element index uses global rows)."""
    dtype = np.dtype(_DT[np.dtype(dtype) if not isinstance(dtype, str) else dtype])
    rows, cols = shape
    out = np.empty((rows, cols), dtype=dtype)
    flat = out.reshape(-1)
    n = rows * cols
    base = row_offset * cols
    for s in range(0, n, chunk):
        t = min(n, s + chunk)
        e = np.arange(base + s, base + t, dtype=np.uint64)
        re, im = gaussian_pair(e, seed, stream)
        if dtype.kind == "c":
            real_t = np.float32 if dtype == np.complex64 else np.float64
            flat[s:t] = re.astype(real_t) + 1j * im.astype(real_t)
        else:
            flat[s:t] = re.astype(dtype)
    return out
