"""Landmark coarsening: host-side mirror of ``/root/reference/src/landmarks.jl``.

Not on the B200 path (north_star keeps landmark selection on the host); it is mirrored
here because every landmark-mode input of the scorer is produced by it and Julia is not
in this image.  All vertex / landmark ids handled here are 1-based like the reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_EPS = float(np.finfo(np.float64).eps)


# ---- priority queue on (what, value), min-heap on value (landmarks.jl:5-46) -------------
def _put(pq, what, value):
    pq.append((what, value))
    i = len(pq)  # 1-based position of the hole
    while i // 2 >= 1:
        j = i // 2
        if value < pq[j - 1][1]:
            pq[i - 1] = pq[j - 1]
            i = j
        else:
            break
    pq[i - 1] = (what, value)


def _pop(pq):
    x = pq[0]
    y = pq.pop()
    if pq:
        i = 1
        n = len(pq)
        while 2 * i <= n:
            l = 2 * i
            r = l + 1
            j = l if (r > n or pq[l - 1][1] < pq[r - 1][1]) else r
            if pq[j - 1][1] < y[1]:
                pq[i - 1] = pq[j - 1]
                i = j
            else:
                break
        pq[i - 1] = y
    return x


# ---- weighted SSE bookkeeping, vectorised over embedding dimensions (landmarks.jl:50-66) --
class _W:
    __slots__ = ("ss", "s", "ws")

    def __init__(self, ss, s, ws):
        self.ss, self.s, self.ws = ss, s, ws

    @staticmethod
    def of(x, w):
        """WSSE of the rows ``x`` (r x d) with weights ``w`` (r), one entry per dimension."""
        d = x.shape[1]
        if x.shape[0] == 0:
            return _W(np.zeros(d), np.zeros(d), np.zeros(d))
        return _W((w[:, None] * x * x).sum(0), (w[:, None] * x).sum(0), np.full(d, w.sum()))

    def __add__(self, o):
        return _W(self.ss + o.ss, self.s + o.s, self.ws + o.ws)

    def __sub__(self, o):
        return _W(self.ss - o.ss, self.s - o.s, self.ws - o.ws)

    def total(self):
        return float((self.ss - self.s ** 2 / self.ws).sum())


def _w_mean(m, w):
    return (m * w[:, None]).sum(0) / w.sum()


# The sign of an eigenvector is LAPACK's to choose (it decides which child of a cut is "low", i.e. the
# numbering of the landmarks, not the partition).  CANONICAL_SIGN = True fixes it the way the device
# selection (cge_b200_landmarks_select) does -- largest-magnitude component positive -- so that the two
# can be compared label by label; the default keeps whatever LAPACK returns, like the reference.
CANONICAL_SIGN = False
# who answers the d x d eigenproblem of a cut when the selection runs on the device (landmarks(..., device=)):
# "lapack" = numpy.linalg.eigh through the library's callback (this mirror's own routine), "builtin" = the
# library's solver with the canonical sign
DEVICE_EIG = os.environ.get("CGE_B200_EIG", "lapack")


def _pc1(m, w):
    """Projection on the first principal component of the weighted, centred rows."""
    y = (m - _w_mean(m, w)) * np.sqrt(w)[:, None]
    _, vec = np.linalg.eigh(y.T @ y)
    v = vec[:, -1]
    if CANONICAL_SIGN and v[int(np.argmax(np.abs(v)))] < 0:
        v = -v
    return y @ v


def total_rss(m, w):
    return _W.of(m, w).total()


def split_cluster_rss(m, w):
    """landmarks.jl:155-210 -- returns two 0-based index arrays into the rows of ``m``."""
    assert m.shape[0] > 1
    if m.shape[0] == 2:
        return np.array([0]), np.array([1])
    z = _pc1(m, w)
    a, b = int(np.argmin(z)), int(np.argmax(z))
    if a == b:
        raise RuntimeError("Trying to split homogenous cluster")
    l1, l2 = [a], [b]
    gray = np.array([i for i in range(z.shape[0]) if i != a and i != b], dtype=np.int64)
    rss_low = _W(m[a] ** 2 * w[a], m[a] * w[a], np.full(m.shape[1], w[a]))
    rss_high = _W(m[b] ** 2 * w[b], m[b] * w[b], np.full(m.shape[1], w[b]))
    med = np.median(z)
    while True:
        mask = z[gray] < med
        t1, t2 = gray[mask], gray[~mask]
        low_tmp = rss_low + _W.of(m[t1], w[t1])
        high_tmp = rss_high + _W.of(m[t2], w[t2])
        if low_tmp.total() < high_tmp.total():
            if t1.size == 0:
                break
            rss_low = low_tmp
            l1.extend(t1.tolist())
            gray = t2.copy()
        else:
            if t2.size == 0:
                break
            rss_high = high_tmp
            l2.extend(t2.tolist())
            gray = t1.copy()
        if gray.size == 0:
            break
        med = np.median(z[gray])
    if gray.size > 0:
        g = _W.of(m[gray], w[gray])
        low_tmp, high_tmp = rss_low + g, rss_high + g
        if max(low_tmp.total(), rss_high.total()) < max(rss_low.total(), high_tmp.total()):
            l1.extend(gray.tolist())
        else:
            l2.extend(gray.tolist())
    return np.asarray(l1, dtype=np.int64), np.asarray(l2, dtype=np.int64)


def split_cluster_rss2(m, w):
    """landmarks.jl:91-147 (sorting variant)."""
    assert m.shape[0] > 1
    if m.shape[0] == 2:
        return np.array([0]), np.array([1])
    z = _pc1(m, w)
    p = np.argsort(z, kind="stable")
    n = p.shape[0]
    one = lambda i: _W.of(m[p[i]:p[i] + 1], w[p[i]:p[i] + 1])  # noqa: E731
    low, high = 0, n - 1
    rss_low, rss_high = one(0), one(n - 1)
    while low + 1 < high:
        if rss_low.total() < rss_high.total():
            low += 1
            rss_low = rss_low + one(low)
        else:
            high -= 1
            rss_high = rss_high + one(high)
    moved_low = False
    while low > 0:
        lt, ht = rss_low - one(low), rss_high + one(low)
        if max(lt.total(), ht.total()) < max(rss_low.total(), rss_high.total()):
            moved_low = True
            low -= 1
            high -= 1
            rss_low, rss_high = lt, ht
        else:
            break
    if not moved_low:
        while high < n - 1:
            lt, ht = rss_low + one(high), rss_high - one(high)
            if max(lt.total(), ht.total()) < max(rss_low.total(), rss_high.total()):
                low += 1
                high += 1
                rss_low, rss_high = lt, ht
            else:
                break
    return p[: low + 1], p[high:]


def _split_by_threshold(z, thr):
    low, high = [], []
    for i, c in enumerate(z):
        if c == thr:
            (low if len(low) < len(high) else high).append(i)
        else:
            (low if c < thr else high).append(i)
    return np.asarray(low, dtype=np.int64), np.asarray(high, dtype=np.int64)


def split_cluster_size(m, w):
    """landmarks.jl:218-238."""
    assert m.shape[0] > 1
    if m.shape[0] == 2:
        return np.array([0]), np.array([1])
    z = _pc1(m, w)
    return _split_by_threshold(z, np.median(z))


def split_cluster_diameter(m, w):
    """landmarks.jl:247-267."""
    assert m.shape[1] > 1
    if m.shape[0] == 2:
        return np.array([0]), np.array([1])
    z = _pc1(m, w)
    return _split_by_threshold(z, (z.min() + z.max()) / 2)


def _split_and_put(pq, idxs, embedding, w, rule):
    low, high = rule(embedding[idxs - 1], w[idxs - 1])
    for part in (idxs[low], idxs[high]):
        if part.size > 1:
            _put(pq, part, -total_rss(embedding[part - 1], w[part - 1]))
        elif part.size == 1:
            _put(pq, part, _EPS)  # singletons go to the tail of the queue
        else:
            raise RuntimeError("Unexpected empty cluster generated")


def runsplit(embedding, w, initial_clusters, n, s, rule):
    """landmarks.jl:279-345 -- returns 0-based landmark ids per vertex."""
    pq = []
    for cluster in sorted(initial_clusters, key=lambda c: c.tolist()):
        if cluster.size <= s:
            for j in cluster:
                _put(pq, np.array([j], dtype=np.int64), _EPS)
        else:
            local = []
            _put(local, cluster, -total_rss(embedding[cluster - 1], w[cluster - 1]))
            while len(local) < s:
                idxs, _ = _pop(local)
                _split_and_put(local, idxs, embedding, w, rule)
            while local:
                idxs, sse = _pop(local)
                _put(pq, idxs, sse)
    while len(pq) < n:
        idxs, _ = _pop(pq)
        _split_and_put(pq, idxs, embedding, w, rule)
    group = np.full(embedding.shape[0], -1, dtype=np.int64)
    for i, (what, _) in enumerate(pq):
        group[what - 1] = i
    assert (group >= 0).all()
    return group


def _idx(n, i, j):
    return n * (i - 1) - (i - 1) * (i - 2) // 2 + j - i + 1


def landmarks(edges, weights, vweights, clusters, comm, embedding, verbose, land, forced,
              method, directed, device=None):
    """Mirror of ``landmarks(...)`` (landmarks.jl:365-465).

    Returns ``(dii, embed, cluster, landmark_edges, weights, lweight, v_to_l)``; all ids 1-based.
    ``device``: a :class:`cge_jl_b200.divergence.Scorer`; ``runsplit`` (``cge_b200_landmarks_select``,
    SURVEY.md 8(f) F4; rss, size and diameter rules) and the aggregation after it (landmarks.jl:387-463,
    ``cge_b200_landmarks_aggregate``) then run on the GPU.
    """
    if verbose:
        print("Starts landmark generation")
    rows_embed, dim = embedding.shape
    unique_rows = (device.unique_rows(embedding) if device is not None
                   else np.unique(embedding, axis=0).shape[0])
    if land > unique_rows:
        print(f"Warning: Requested number of clusters larger than unique no. embeddings. "
              f"Truncating to {unique_rows} landmarks.", file=sys.stderr)
        land = unique_rows
    rule = {split_cluster_rss: "rss", split_cluster_size: "size",
            split_cluster_diameter: "diameter"}.get(method)
    if device is not None and rule is not None:  # the cuts run on the GPU
        lm = device.landmarks_select(embedding, vweights, clusters, land, forced, rule, eig=DEVICE_EIG)[0] + 1
    else:
        lm = runsplit(embedding, vweights, clusters, land, forced, method) + 1
    if verbose:
        print("Landmarks generated")
    N = int(lm.max())
    if verbose:
        print(f"Using {N} landmarks")
    if device is not None:
        dii, embed, cluster, landmark_edges, lw, lweight = device.landmarks_aggregate(
            lm, vweights, comm, embedding, edges, weights, directed, N)
        return dii, embed, cluster, landmark_edges, lw, lweight, lm
    return aggregate_host(lm, edges, weights, vweights, comm, embedding, directed) + (lm,)


def aggregate_host(lm, edges, weights, vweights, comm, embedding, directed):
    """landmarks.jl:387-463 on the host, in the reference's order of additions (``np.add.at`` and
    ``np.cumsum`` add sequentially; products and sums are rounded separately as in the Julia loops).
    The CPU side of tests/test_gpu_landmarks.py; returns everything but ``v_to_l``."""
    N = int(lm.max())
    dim = embedding.shape[1]
    l0 = lm - 1
    lweight = np.zeros(N)
    np.add.at(lweight, l0, vweights)
    embed = np.zeros((N, dim))
    np.add.at(embed, l0, vweights[:, None] * embedding)
    embed /= lweight[:, None]
    # d_ii: unweighted squared deviations over the landmark's weight, then sqrt (landmarks.jl:407-423);
    # the inner sum over the dimensions runs left to right like the reference's `dist +=`
    dii = np.zeros(N)
    np.add.at(dii, l0, np.cumsum((embed[l0] - embedding) ** 2, axis=1)[:, -1])
    pos = lweight > 0
    dii[pos] = np.sqrt(dii[pos] / lweight[pos])
    cluster = np.zeros(N, dtype=np.int64)
    cluster[l0] = comm[:, 0]  # last member wins, as in the sequential loop
    cluster = cluster.reshape(-1, 1)
    a = l0[edges[:, 0] - 1]
    b = l0[edges[:, 1] - 1]
    wedges = np.zeros((N, N))
    if directed:
        np.add.at(wedges, (a, b), weights)
        ii, jj = np.nonzero(wedges > 0)  # row-major order == N*(i-1)+j order
    else:
        np.add.at(wedges, (np.minimum(a, b), np.maximum(a, b)), weights)
        ii, jj = np.nonzero(np.triu(wedges) > 0)  # row-major upper triangle == idx order
    landmark_edges = np.stack([ii + 1, jj + 1], axis=1).astype(np.int64)
    lw = wedges[ii, jj].copy()
    return dii, embed, cluster, landmark_edges, lw, lweight
