"""Index arithmetic of the DMMA GEMM fragment loads (rusty_compression_b200/csrc/gemm_dmma.cu, dmma_consumer), checked
on CPU: the per-lane address tables (six constants + immediates) must address exactly the bytes the direct formula of
the 128-byte-swizzled TMA boxes gives, for every lane, warp, column group, k slot and tile width, and the reads of a
warp must be bank-conflict free.  Pure integer work, so the bar is bit-exact."""
import itertools

import pytest

BM, BK = 64, 16


def direct_b_offset(col, k):
    """B tile = boxes of [16 k][16 doubles], 128-byte rows, 16-byte chunk index XOR (k & 7)."""
    box, cb = col >> 4, col & 15
    return box * 2048 + k * 128 + ((((cb >> 1) ^ (k & 7))) << 4) + ((cb & 1) << 3)


def direct_a_nn_offset(row, kchunk):
    """A tile (NN) = [64 rows][16 doubles]: one LDS.128 per row at chunk (4 kg + t4) ^ (row & 7)."""
    return row * 128 + ((kchunk ^ (row & 7)) << 4)


@pytest.mark.parametrize("bn", [16, 32, 48, 64, 80, 96])
def test_fragment_address_tables_match_the_swizzle_formula(bn):
    wn, cg = bn // 2, bn // 16
    for warp, lane in itertools.product(range(8), range(32)):
        g, t4 = lane >> 2, lane & 3
        w1 = (warp ^ (warp >> 2)) & 1
        wm0, wn0 = (warp >> 1) * 16, w1 * wn
        rho = 4 * (g & 1) + (g >> 1)
        L = [(2 * t4 + h) * 128 + ((((g >> 1) ^ (2 * t4 + h))) << 4) + ((g & 1) << 3) for h in range(2)]
        boff = [[(((w1 * cg + jp) >> 1) * 2048) + (L[h] ^ (((w1 * cg + jp) & 1) << 6)) for jp in range(2)] for h in range(2)]
        a_nn = [(wm0 + rho) * 128 + ((t4 ^ rho) << 4)]
        a_nn.append(a_nn[0] ^ 64)
        a_tn = [[(warp >> 1) * 2048 + (L[h] ^ (i << 6)) for i in range(2)] for h in range(2)]
        for kg, h in itertools.product(range(2), range(2)):
            k = kg * 8 + 2 * t4 + h
            for j in range(cg):
                assert boff[h][j & 1] + (j >> 1) * 2048 + kg * 1024 == direct_b_offset(wn0 + j * 8 + g, k)
            for i in range(2):       # transposed A: the same box layout with column = tile row index
                assert a_tn[h][i] + kg * 1024 == direct_b_offset(wm0 + i * 8 + g, k)
        for kg, i in itertools.product(range(2), range(2)):
            assert a_nn[kg] + i * 1024 == direct_a_nn_offset(wm0 + i * 8 + rho, kg * 4 + t4)
        # tail columns BN - 8 + 2 c2 (TAILW = 2, 4): one LDS.128 per (h, c2)
        for kg, h, c2 in itertools.product(range(2), range(2), range(2)):
            k = kg * 8 + 2 * t4 + h
            toff = (bn // 16 - 1) * 2048 + (2 * t4 + h) * 128 + ((4 ^ (2 * t4 + h)) << 4)
            assert (toff ^ (c2 << 4)) + kg * 1024 == direct_b_offset(bn - 8 + 2 * c2, k)


def test_warp_roles_cover_the_tile_and_balance_the_subpartitions():
    roles = {((w >> 1), (w ^ (w >> 2)) & 1) for w in range(8)}
    assert roles == {(r, c) for r in range(4) for c in range(2)}          # 4 row groups x 2 column halves, each once
    for sp in range(4):                                                   # warp & 3 = SM sub-partition
        halves = sorted((w ^ (w >> 2)) & 1 for w in range(8) if w & 3 == sp)
        assert halves == [0, 1]                                           # one full and one ragged-tail warp each


@pytest.mark.parametrize("trans", [False, True])
def test_fragment_loads_are_bank_conflict_free(trans):
    """Shared memory serves 128 bytes (32 banks x 4 bytes) per cycle: a warp-wide LDS.64 runs in two half-warp phases
    and is conflict free when the 16 words of each phase fall in 16 distinct 8-byte bank pairs; an LDS.128 runs in four
    quarter-warp phases of 8 x 16 bytes."""
    for warp in range(8):
        w1 = (warp ^ (warp >> 2)) & 1
        for kg, h, j in itertools.product(range(2), range(2), range(5)):
            for half in range(2):
                words = set()
                for lane in range(16 * half, 16 * half + 16):
                    g, t4 = lane >> 2, lane & 3
                    col = (warp >> 1) * 16 + (j & 1) * 8 + g if trans else w1 * 40 + j * 8 + g      # transposed A uses the same boxes
                    off = direct_b_offset(col, kg * 8 + 2 * t4 + h)
                    words.add((off % 128) // 8)
                assert len(words) == 16
        if not trans:
            for kg, i in itertools.product(range(2), range(2)):
                for quarter in range(4):
                    chunks = set()
                    for lane in range(8 * quarter, 8 * quarter + 8):
                        g, t4 = lane >> 2, lane & 3
                        rho = 4 * (g & 1) + (g >> 1)
                        off = direct_a_nn_offset((warp >> 1) * 16 + i * 8 + rho, kg * 4 + t4)
                        chunks.add((off % 128) // 16)
                    assert len(chunks) == 8


@pytest.mark.parametrize("n", [2, 3, 7, 64, 74, 129])
def test_jacobi_round_robin_schedule_visits_every_pair_once_per_sweep(n):
    """jacobi_kernel (csrc/jacobi.cu): round r, slot k rotates the pair (a, b) below; within a round the pairs are
    disjoint (one warp each, no races on columns), over a sweep every unordered pair of the n columns appears once."""
    npad = n + (n & 1)
    half = npad // 2
    seen = set()
    for rnd in range(npad - 1):
        cols = set()
        for k in range(half):
            if k == 0:
                a, b = npad - 1, rnd
            else:
                a, b = (rnd + k) % (npad - 1), (rnd - k + (npad - 1)) % (npad - 1)
            p, q = min(a, b), max(a, b)
            assert p != q and p not in cols and q not in cols
            cols.update((p, q))
            if q < n:                      # the padding column of an odd n is skipped
                assert (p, q) not in seen
                seen.add((p, q))
    assert len(seen) == n * (n - 1) // 2
