"""Pins the oracle's Philox4x32-10 to the published Random123 known-answer vectors."""
import numpy as np

from oracle.philox import philox4x32_10, random_gaussian

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        got = tuple(int(x) for x in philox4x32_10(*ctr, *key))
        assert got == want


def test_gaussian_moments_and_shard_consistency():
    g = random_gaussian((4096, 64), "d", seed=42)
    assert abs(g.mean()) < 0.01 and abs(g.std() - 1.0) < 0.01
    # a row shard regenerates exactly its slice
    part = random_gaussian((100, 64), "d", seed=42, row_offset=1000)
    assert np.array_equal(part, g[1000:1100])
    # complex: independent N(0,1) real and imaginary parts (variance 2 in total)
    z = random_gaussian((4096, 16), "z", seed=3)
    assert abs(z.real.std() - 1.0) < 0.02 and abs(z.imag.std() - 1.0) < 0.02
    assert abs(np.mean(z.real * z.imag)) < 0.02
    # f32 is the f64 draw cast down (src/random_matrix.rs:123)
    s = random_gaussian((128, 8), "s", seed=42)
    assert np.array_equal(s, random_gaussian((128, 8), "d", seed=42).astype(np.float32))
