// Philox4x32-10 counter-based generator + Box-Muller, bit-compatible with oracle/philox.py.
// Replaces rand_distr::Normal (src/random_matrix.rs:120-145) for seeded / sharded runs:
// counter = (lo32(e), hi32(e), stream, 0), key = (lo32(seed), hi32(seed)), e = row-major
// element index, so any row shard regenerates exactly its slice of Omega.
#pragma once
#include <cstdint>

__host__ __device__ __forceinline__ void rc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Two independent N(0,1) doubles for element index e.
__device__ __forceinline__ void rc_philox_gaussian_pair(uint64_t e, uint64_t seed, uint32_t stream, double& re, double& im) {
    uint32_t w[4];
    rc_philox4x32_10((uint32_t)e, (uint32_t)(e >> 32), stream, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    const double two26 = 67108864.0, two_m53 = 1.1102230246251565404e-16;
    double u1 = ((double)(w[0] >> 5) * two26 + (double)(w[1] >> 6) + 1.0) * two_m53;   // (0, 1]
    double u2 = ((double)(w[2] >> 5) * two26 + (double)(w[3] >> 6)) * two_m53;         // [0, 1)
    double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    re = r * c;
    im = r * s;
}
