"""CPU: the control flow of the device landmark selection's rss cut (cge_select.cu: cut_segment) restated in
NumPy and held to the host mirror of the reference (cge_jl_b200.landmarks.split_cluster_rss,
landmarks.jl:155-210).  The device code never materialises the reference's index lists: it sorts the
projection z once and works on RANK RANGES -- "gray" is always a contiguous range of the sorted order, every
accepted half is a range, and the final order of the segment is a stable regrouping by range.  This test
checks that formulation (including first-occurrence argmin / argmax under ties, the median rule, the final
assignment of the leftover range and the member order inside both children) on random clusters with
duplicated rows; the kernels themselves are covered by tests/test_gpu_select.py."""
import importlib

import numpy as np

lm = importlib.import_module("cge_jl_b200.landmarks")


def _moments(m, w):
    if m.shape[0] == 0:
        return np.zeros(1 + 2 * m.shape[1])
    return np.concatenate([[w.sum()], (w[:, None] * m).sum(0), (w[:, None] * m * m).sum(0)])


def _total(v, d):
    return float((v[1 + d:] - v[1:1 + d] ** 2 / v[0]).sum())


def rank_range_rss_cut(m, w):
    """cut_segment(), rule rss: returns the members of the two children in their new order."""
    n, d = m.shape
    if n == 2:
        return [0], [1]
    z = lm._pc1(m, w)
    perm = np.argsort(z, kind="stable")  # the device's stable radix sort of (z, position)
    zs = z[perm]
    rb = n - 1  # first occurrence of the maximum = first rank holding the top value
    while rb > 0 and zs[rb - 1] == zs[n - 1]:
        rb -= 1
    if rb != n - 1:  # tied maximum: rotate it to the last rank
        tail = list(perm[rb:])
        perm[rb:] = tail[1:] + tail[:1]
    rng = lambda a, b: _moments(m[perm[a:b]], w[perm[a:b]])  # noqa: E731 -- k_sel_moments on a rank range
    rss_low, rss_high = rng(0, 1), rng(n - 1, n)
    lo, hi = 1, n - 1  # gray = ranks [lo, hi)
    acc_low, acc_high = [], []

    def median(a, b):
        c = b - a
        return zs[a + c // 2] if c % 2 else 0.5 * (zs[a + c // 2 - 1] + zs[a + c // 2])

    if hi > lo:
        med = median(0, n)
        while True:
            p = lo + int(np.searchsorted(zs[lo:hi], med, side="left"))
            low_tmp, high_tmp = rss_low + rng(lo, p), rss_high + rng(p, hi)
            if _total(low_tmp, d) < _total(high_tmp, d):
                if p == lo:
                    break
                rss_low = low_tmp
                acc_low.append((lo, p))
                lo = p
            else:
                if p == hi:
                    break
                rss_high = high_tmp
                acc_high.append((p, hi))
                hi = p
            if lo == hi:
                break
            med = median(lo, hi)
    gray_low = False
    if hi > lo:
        g = rng(lo, hi)
        gray_low = (max(_total(rss_low + g, d), _total(rss_high, d))
                    < max(_total(rss_low, d), _total(rss_high + g, d)))
    groups_low = [(0, 1)] + acc_low + ([(lo, hi)] if hi > lo and gray_low else [])
    groups_high = [(n - 1, n)] + acc_high + ([(lo, hi)] if hi > lo and not gray_low else [])
    rank = np.empty(n, dtype=np.int64)
    rank[perm] = np.arange(n)

    def members(groups):  # stable regrouping: group by group, previous member order inside a group
        return [i for a, b in groups for i in range(n) if a <= rank[i] < b]

    low, high = members(groups_low), members(groups_high)
    assert len(low) == (hi if gray_low else lo)  # the device's cut position
    return low, high


def test_rank_range_formulation_equals_the_reference_lists():
    rng = np.random.default_rng(1)
    compared = 0
    for t in range(1500):
        n, d = int(rng.integers(3, 40)), int(rng.integers(1, 9))
        m = rng.normal(size=(n, d))
        w = rng.uniform(0.5, 3.0, size=n)
        if t % 5 == 0:
            m[1] = m[0]  # a duplicated row: ties in z
        if t % 7 == 0:
            m[int(rng.integers(0, n))] = m.max(axis=0) + 5.0  # a far point
            m[int(rng.integers(0, n))] = m[int(rng.integers(0, n))]
        try:
            a, b = lm.split_cluster_rss(m, w)
        except RuntimeError:
            continue
        a2, b2 = rank_range_rss_cut(m, w)
        assert list(a) == a2 and list(b) == b2, (t, n, d)
        compared += 1
    assert compared > 1400
