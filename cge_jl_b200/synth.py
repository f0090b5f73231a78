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
