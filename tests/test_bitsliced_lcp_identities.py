"""not gpu: the logic k_derive (csrc/unpack.cu) and k_cluster_scan (csrc/scan.cu) rest on, modelled in numpy bit for bit:
the LCP as 7 bit planes + the plane A (bit x = lcp[x-1] > lcp[x]) in 64-position groups, written range by range in any
order; "lcp >= k" as the 7-step bit-sliced compare whose step is the 3-input function with truth table 0x8E; the START / END
masks from G, A and the four neighbour bits.  The kernels themselves are compared with the oracle on the GPU
(tests/test_gpu_parity.py, tests/test_streaming_gpu.py)."""
import numpy as np

from tests import helpers as H

U64 = np.uint64
ONES = U64(0xFFFFFFFFFFFFFFFF)


def lop3(a, b, c, table):
    """the GPU's 3-input logic instruction: bit i of the result = table[(a_i << 2) | (b_i << 1) | c_i]"""
    out = np.zeros_like(a)
    for idx in range(8):
        if (table >> idx) & 1:
            ta = a if idx & 4 else ~a
            tb = b if idx & 2 else ~b
            tc = c if idx & 1 else ~c
            out |= ta & tb & tc
    return out


def derive(groups, lcp, a, b):
    """k_derive on the positions [a, b): plane bits of [a, b), A bits of [a, b] (bit b: its left neighbour is new).
    groups[g, w] = word w of the group of positions [64 g, 64 g + 64)"""
    n = len(lcp)
    for x in range(a, min(b + 1, n)):
        g, bit = x >> 6, U64(x & 63)
        if x < b:
            v = min(int(lcp[x]), 127)
            for p in range(7):
                groups[g, p] = (groups[g, p] & ~(U64(1) << bit)) | (U64((v >> p) & 1) << bit)
        prev = int(lcp[x - 1]) if x > 0 else 0
        groups[g, 7] = (groups[g, 7] & ~(U64(1) << bit)) | (U64(prev > int(lcp[x])) << bit)


def scan_masks(groups, n, k):
    """the mask section of k_cluster_scan for an eBWT held by one shard: START / END as booleans for positions 0..n-1"""
    kk = min(k, 128)
    km = [ONES if (kk >> p) & 1 else U64(0) for p in range(7)]
    lt = np.zeros(len(groups), dtype=U64)
    for p in range(7):
        lt = lop3(groups[:, p], np.full_like(lt, km[p]), lt, 0x8E)
    G = ~lt if kk < 128 else np.zeros_like(lt)
    A = groups[:, 7]
    ng = len(groups)
    g_p = np.concatenate([G[1:] & U64(1), [U64(0)]])          # G / A of the position after my 64 (zero padding after the array)
    a_p = np.concatenate([A[1:] & U64(1), [U64(0)]])
    g_m1 = np.concatenate([[U64(0)], G[:-1] >> U64(63)])      # ... of the position before them (zero padding before position 0)
    a_m1 = np.concatenate([[U64(0)], A[:-1] >> U64(63)])
    Gn = (G >> U64(1)) | (g_p << U64(63))
    An = (A >> U64(1)) | (a_p << U64(63))
    E = G & ((A & ~An) | ~Gn)
    e_prev = g_m1 & ((a_m1 & ~A) | ~G) & U64(1)
    # the init special cases (global position 0), the host tail rule (END(n-1) left out), positions past n
    E[0] &= ~U64(1)
    e_prev[0] = U64(0)
    if (G[0] & U64(1)) and not (G[0] & U64(2)):
        E[0] |= U64(2)
    pos = np.arange(ng * 64, dtype=np.int64).reshape(ng, 64)
    vm = np.zeros(ng, dtype=U64)
    for j in range(64):
        vm |= (pos[:, j] < n).astype(U64) << U64(j)
    last = n - 1
    E[last >> 6] &= ~(U64(1) << U64(last & 63))
    E &= vm
    Gp = (G << U64(1)) | g_m1
    Ep = (E << U64(1)) | e_prev
    S = G & (~Gp | Ep) & vm
    bits = lambda M: ((M[:, None] >> np.arange(64, dtype=U64)[None, :]) & U64(1)).astype(bool).reshape(-1)[:n]
    return bits(S), bits(E)


def test_lt_step_truth_table():
    """0x8E = (k_b ? ~x | lt : ~x & lt) with a = x, b = k_b mask, c = lt"""
    for x in (0, 1):
        for kb in (0, 1):
            for lt in (0, 1):
                want = ((1 - x) | lt) if kb else ((1 - x) & lt)
                assert (0x8E >> ((x << 2) | (kb << 1) | lt)) & 1 == want


def test_bitsliced_compare_all_values_all_k():
    v = np.arange(128)
    groups = np.zeros((2, 8), dtype=U64)
    derive(groups, v, 0, 128)
    for k in list(range(0, 131)) + [200, 255, 256, 70000]:
        kk = min(k, 128)
        lt = np.zeros(2, dtype=U64)
        for p in range(7):
            lt = lop3(groups[:, p], np.full(2, ONES if (kk >> p) & 1 else 0, dtype=U64), lt, 0x8E)
        G = ~lt if kk < 128 else np.zeros_like(lt)
        got = ((G[:, None] >> np.arange(64, dtype=U64)[None, :]) & U64(1)).astype(bool).reshape(-1)
        assert np.array_equal(got, v >= k), k


def test_masks_equal_the_stencil_ranges_in_any_order():
    rng = np.random.default_rng(11)
    for it in range(40):
        n = int(rng.choice([2, 3, 63, 64, 65, 127, 128, 129, 1000, 4099]))
        k = int(rng.choice([1, 2, 3, 16, 70, 127, 128, 300]))
        lcp = np.minimum(H.random_lcp(rng, n, min(k, 100), it % 5), 127).astype(np.uint32)
        ng = (n + 63) // 64 + 1
        groups = rng.integers(0, 1 << 63, size=(ng, 8)).astype(U64)  # stale contents everywhere
        groups[(n >> 6):, :] = 0                                      # ... except past the end: the phantom / pad is derived too
        cuts = sorted(set([0, n] + [int(c) for c in rng.integers(0, n + 1, size=3)]))
        ranges = list(zip(cuts[:-1], cuts[1:]))
        rng.shuffle(ranges)
        padded = np.concatenate([lcp, np.zeros(ng * 64 - n, dtype=np.uint32)])
        for a, b in ranges:
            derive(groups, padded, a, b)
        derive(groups, padded, n, ng * 64)  # the pad (zeros): what the seal's phantom fill / the zeroed allocation leave there
        S, E = scan_masks(groups, n, k)
        ws, we = H.flags(lcp, k)
        assert np.array_equal(E, we), (it, n, k)
        assert np.array_equal(S, ws), (it, n, k)
