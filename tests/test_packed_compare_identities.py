"""not gpu: the arithmetic identities k_lcp_flags8 (csrc/cluster.cu) rests on, checked exhaustively in numpy with 32-bit
wrap-around -- the kernel itself is compared with the oracle on the GPU (tests/test_gpu_parity.py::test_cluster_narrow_lcp)."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def gather_b7(x):
    """bit 7 of bytes 0..3 -> bits 28..31 (one multiply)"""
    return ((x & np.uint64(0x80808080)) * np.uint64(0x00204081)) & M32


def test_ge_by_biased_add_all_k_all_values():
    v = np.arange(128, dtype=np.uint64)
    words = v | (v[::-1] << np.uint64(8)) | (np.uint64(127) << np.uint64(16)) | (np.uint64(0) << np.uint64(24))
    for k in list(range(1, 131)) + [200, 255, 256, 300, 70000]:
        kk = min(k, 128)
        t = (words + np.uint64((128 - kk) * 0x01010101)) & M32
        nib = gather_b7(t) >> np.uint64(28)
        want = (v >= k).astype(np.uint64) | ((v[::-1] >= k).astype(np.uint64) << np.uint64(1)) | (np.uint64(127 >= k) << np.uint64(2))
        assert np.array_equal(nib, want), k


def test_strictly_greater_than_previous_as_multiply_add():
    """(prev-word | 0x80808080) - cur - 0x01010101  ==  255 * cur + (w_prev >> 24) + 0x7f7f7f7f  (mod 2^32), and its bit 7 per
    byte says lcp[j-1] > lcp[j]; exhaustive over every (previous byte, byte) pair in every byte lane"""
    a, b = np.meshgrid(np.arange(128, dtype=np.uint64), np.arange(128, dtype=np.uint64), indexing="ij")
    a, b = a.ravel(), b.ravel()  # a = previous position's value, b = this position's
    rng = np.random.default_rng(3)
    for lane in range(4):
        other = rng.integers(0, 128, size=(len(a), 4)).astype(np.uint64)
        cur_bytes = other.copy()
        cur_bytes[:, lane] = b
        prev_last = rng.integers(0, 128, size=len(a)).astype(np.uint64)  # byte 3 of the previous word
        if lane == 0:
            prev_last = a
        else:
            cur_bytes[:, lane - 1] = a
        cur = sum(cur_bytes[:, j] << np.uint64(8 * j) for j in range(4))
        w_prev = (rng.integers(0, 1 << 24, size=len(a)).astype(np.uint64)) | (prev_last << np.uint64(24))
        prv = ((cur << np.uint64(8)) | (w_prev >> np.uint64(24))) & M32
        direct = ((prv | np.uint64(0x80808080)) - cur - np.uint64(0x01010101)) & M32
        fma = (cur * np.uint64(255) + ((w_prev * np.uint64(256)) >> np.uint64(32)) + np.uint64(0x7F7F7F7F)) & M32
        assert np.array_equal(direct, fma)
        nib = gather_b7(fma) >> np.uint64(28)
        prev_of = np.concatenate([prev_last[:, None], cur_bytes[:, :3]], axis=1)
        want = sum((prev_of[:, j] > cur_bytes[:, j]).astype(np.uint64) << np.uint64(j) for j in range(4))
        assert np.array_equal(nib, want), lane


def test_gather_ignores_low_bits():
    rng = np.random.default_rng(4)
    x = rng.integers(0, 1 << 32, size=200000, dtype=np.uint64)
    nib = gather_b7(x) >> np.uint64(28)
    want = ((x >> np.uint64(7)) & np.uint64(1)) | (((x >> np.uint64(15)) & np.uint64(1)) << np.uint64(1)) | \
           (((x >> np.uint64(23)) & np.uint64(1)) << np.uint64(2)) | (((x >> np.uint64(31)) & np.uint64(1)) << np.uint64(3))
    assert np.array_equal(nib, want)
