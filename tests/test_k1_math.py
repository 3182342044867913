"""CPU: the two pieces of arithmetic the row form of K1 (pyascore_b200/csrc/pa_bin_rows.cuh) leans on, restated in numpy.

(1) Counting with FSET.BF: a compare leaves the bits of 1.0f (0x3f800000 = 127 << 23) or 0, the kernel adds those words as
    integers modulo 2^32 and recovers the count as ((sum >> 23) * 383) & 511 -- valid for counts below 512 because
    383 * 127 = 1 (mod 512).
(2) Ties by a sum: with c_i = number of keys of the bin greater than key i, the c_i of a bin of n peaks are a permutation
    of 0 .. n-1 when the keys are distinct, and add up to less than n (n - 1) / 2 as soon as two keys are equal (or NaN).
"""
import numpy as np


def test_fset_bits_count_decode():
    one = np.uint64(0x3f800000)
    for n in range(512):
        s = (np.uint64(n) * one) & np.uint64(0xffffffff)
        assert int(((s >> np.uint64(23)) * np.uint64(383)) & np.uint64(511)) == n
    assert (383 * 127) % 512 == 1
    # 512 wraps to 0: why runs of 512 or more peaks are declined
    s = (np.uint64(512) * one) & np.uint64(0xffffffff)
    assert int(((s >> np.uint64(23)) * np.uint64(383)) & np.uint64(511)) == 0


def _counts(keys):
    k = np.asarray(keys, np.float32)
    return (k[None, :] > k[:, None]).sum(1)          # NaN compares false both ways, as FSET does


def test_rank_sum_detects_every_tie():
    rng = np.random.default_rng(3)
    for _ in range(300):
        n = int(rng.integers(1, 40))
        keys = rng.permutation(np.arange(n, dtype=np.float32) * 1.5 + 7.)      # distinct
        c = _counts(keys)
        assert sorted(c) == list(range(n)) and c.sum() == n * (n - 1) // 2
        if n >= 2:
            t = keys.copy()
            i, j = rng.choice(n, 2, replace=False)
            t[i] = t[j]                                                         # one tie
            assert _counts(t).sum() < n * (n - 1) // 2
            t = keys.copy()
            t[int(rng.integers(n))] = np.nan                                    # a NaN key
            assert _counts(t).sum() < n * (n - 1) // 2
            z = keys.copy()
            z[i], z[j] = 0., -0.                                                # +0 and -0 compare equal
            assert _counts(z).sum() < n * (n - 1) // 2
    # several bins: the spectrum's total matches exactly when every bin is tie-free
    bins = [rng.permutation(np.arange(m, dtype=np.float32)) for m in (3, 17, 1, 9)]
    want = sum(len(b) * (len(b) - 1) // 2 for b in bins)
    assert sum(_counts(b).sum() for b in bins) == want
    bins[1][4] = bins[1][5]
    assert sum(_counts(b).sum() for b in bins) < want
