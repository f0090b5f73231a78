"""Deterministic synthetic inputs for the benchmark and the scale tests:
planted-partition graphs with clustered random embeddings, in the reference's conventions
(1-based ids, ``parseargs``-style arrays)."""
from __future__ import annotations

import numpy as np


def planted_partition(n, k=64, d=32, degree=16, p_in=0.75, seed=0, directed=False,
                      weighted=False):
    """n vertices in k equal communities (vertex i -> community i*k//n), expected degree
    ``degree`` with ``p_in`` of the edges inside the community, plus a ring per community so
    that no vertex is isolated.  Embedding x_i = mu_c(i) + 0.7*N(0,I_d), mu_c ~ 0.5*N(0,I_d).

    Returns ``(edges[m,2] int64 1-based, eweights[m], vweights[n], comm[n,1] 1-based, embed[n,d])``.
    """
    rng = np.random.default_rng(seed)
    comm0 = (np.arange(n, dtype=np.int64) * k) // n
    start = np.searchsorted(comm0, np.arange(k))
    size = np.diff(np.append(start, n))
    # ring inside each community
    nxt = np.arange(n, dtype=np.int64) + 1
    last = start + size - 1
    nxt[last] = start
    ring = np.stack([np.arange(n, dtype=np.int64), nxt], axis=1)
    ring = ring[ring[:, 0] != ring[:, 1]]
    m_target = n * degree // 2
    u = rng.integers(0, n, size=m_target)
    intra = rng.random(m_target) < p_in
    v_in = start[comm0[u]] + (rng.random(m_target) * size[comm0[u]]).astype(np.int64)
    v_out = rng.integers(0, n, size=m_target)
    v = np.where(intra, v_in, v_out)
    e = np.concatenate([ring, np.stack([u, v], axis=1)])
    e = e[e[:, 0] != e[:, 1]]
    if not directed:
        e = np.stack([e.min(axis=1), e.max(axis=1)], axis=1)
    e = np.unique(e, axis=0)
    if directed:  # orient at random
        flip = rng.random(e.shape[0]) < 0.5
        e[flip] = e[flip][:, ::-1]
        e = np.unique(e, axis=0)
    w = rng.uniform(0.5, 2.0, size=e.shape[0]) if weighted else np.ones(e.shape[0])
    edges = e + 1
    vw = np.zeros(n)
    np.add.at(vw, e[:, 0], w)
    np.add.at(vw, e[:, 1], w)
    mu = 0.5 * rng.normal(size=(k, d))
    emb = mu[comm0] + 0.7 * rng.normal(size=(n, d))
    return edges, w, vw, (comm0 + 1).reshape(-1, 1), emb


def abcd_like(n, k=64, d=128, gamma=2.5, beta=1.5, dmin=5, dmax=50, xi=0.2, seed=0):
    """ABCD-style undirected benchmark graph (BASELINE config 4): power-law degrees
    (exponent ``gamma`` in [dmin, dmax]), power-law community sizes (exponent ``beta``), a fraction
    ``xi`` of every vertex's edge stubs leaves its community.  Same return convention as
    :func:`planted_partition`.  Vertices of a community are contiguous; sizes are very unequal,
    which exercises community boundaries inside tiles and tiny / huge communities."""
    rng = np.random.default_rng(seed)
    # community sizes ~ s^-beta on [n/(8k), 4n/k], rescaled to sum to n, every community >= 2
    lo, hi = max(2.0, n / (8.0 * k)), 4.0 * n / k
    u = rng.random(k)
    sizes = (lo ** (1 - beta) + u * (hi ** (1 - beta) - lo ** (1 - beta))) ** (1 / (1 - beta))
    sizes = np.maximum(2, np.floor(sizes * n / sizes.sum())).astype(np.int64)
    sizes[np.argmax(sizes)] += n - sizes.sum()
    start = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    comm0 = np.repeat(np.arange(k, dtype=np.int64), sizes)
    u = rng.random(n)
    deg = (dmin ** (1 - gamma) + u * (dmax ** (1 - gamma) - dmin ** (1 - gamma))) ** (1 / (1 - gamma))
    deg = np.floor(deg).astype(np.int64)
    src = np.repeat(np.arange(n, dtype=np.int64), (deg + 1) // 2)  # each stub pair -> one edge
    inside = rng.random(src.shape[0]) >= xi
    v_in = start[comm0[src]] + (rng.random(src.shape[0]) * sizes[comm0[src]]).astype(np.int64)
    v_out = rng.integers(0, n, size=src.shape[0])
    dst = np.where(inside, v_in, v_out)
    nxt = np.arange(n, dtype=np.int64) + 1          # ring per community: no isolated vertex
    nxt[start + sizes - 1] = start
    e = np.concatenate([np.stack([np.arange(n, dtype=np.int64), nxt], axis=1),
                        np.stack([src, dst], axis=1)])
    e = e[e[:, 0] != e[:, 1]]
    e = np.unique(np.stack([e.min(axis=1), e.max(axis=1)], axis=1), axis=0)
    w = np.ones(e.shape[0])
    vw = np.zeros(n)
    np.add.at(vw, e[:, 0], w)
    np.add.at(vw, e[:, 1], w)
    mu = 0.5 * rng.normal(size=(k, d))
    emb = mu[comm0] + 0.7 * rng.normal(size=(n, d))
    return e + 1, w, vw, (comm0 + 1).reshape(-1, 1), emb
