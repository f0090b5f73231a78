"""Shared helpers for the parity tests: fixtures, synthetic graphs, comparison rules."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-9  # north_star: scores within 1e-9 relative, best alpha identical


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name))
    return z["edges"], z["eweights"], z["vweights"], z["comm"], z["embedding"]


def clusters_of(comm):
    by = {}
    for v, c in enumerate(comm[:, 0], start=1):
        by.setdefault(int(c), []).append(v)
    return [np.asarray(v, dtype=np.int64) for v in by.values()]


def empty_landmark_args():
    """The empty arrays CGE_CLI.jl:5-9 passes in exact mode."""
    return (np.zeros(0), np.zeros(0, dtype=np.int64), np.zeros((0, 0), dtype=np.int64),
            np.zeros(0), np.zeros((0, 0)))


def planted_partition(n, k, d, seed, directed=False, weighted=False, deg=8):
    """Small synthetic planted-partition graph + clustered embedding (1-based ids)."""
    rng = np.random.default_rng(seed)
    comm0 = rng.integers(0, k, size=n)
    comm0[:k] = np.arange(k)  # every community non-empty
    edges = set()
    # a ring so that every vertex has in- and out-degree >= 1
    for v in range(n):
        edges.add((v, (v + 1) % n))
    while len(edges) < n * deg // 2:
        u = int(rng.integers(0, n))
        if rng.random() < 0.75:
            cand = np.flatnonzero(comm0 == comm0[u])
            v = int(cand[rng.integers(0, cand.size)])
        else:
            v = int(rng.integers(0, n))
        if u == v:
            continue
        e = (u, v) if directed else (min(u, v), max(u, v))
        edges.add(e)
    e = np.array(sorted(edges), dtype=np.int64) + 1
    w = rng.uniform(0.5, 2.0, size=e.shape[0]) if weighted else np.ones(e.shape[0])
    mu = rng.normal(size=(k, d)) * 0.5
    emb = mu[comm0] + 0.7 * rng.normal(size=(n, d))
    vw = np.zeros(n)
    np.add.at(vw, e[:, 0] - 1, w)
    np.add.at(vw, e[:, 1] - 1, w)
    return e, w, vw, (comm0 + 1).reshape(-1, 1), emb


def assert_parity(out, stats, ref_out, ref_tr, check_auc=True):
    """GPU result vs oracle: identical control flow, scores within RTOL."""
    n_run = int(ref_tr.n_alpha_run)
    assert int(stats.n_alpha_run) == n_run
    assert list(stats.iters)[:n_run] == list(ref_tr.iters)[:n_run], "fixed-point pass counts differ"
    assert out.shape == ref_out.shape
    assert out[0] == ref_out[0], "best alpha (global) differs"
    np.testing.assert_allclose(out[1:4], ref_out[1:4], rtol=RTOL, atol=0)
    d_gpu, d_ref = np.array(list(stats.div)), np.array(list(ref_tr.div))
    assert np.array_equal(np.isnan(d_gpu), np.isnan(d_ref))
    ok = ~np.isnan(d_ref)
    np.testing.assert_allclose(d_gpu[ok], d_ref[ok], rtol=RTOL, atol=0)
    if check_auc:
        assert out[4] == ref_out[4], "best alpha (local) differs"
        np.testing.assert_allclose(out[5:7], ref_out[5:7], rtol=RTOL, atol=1e-15)
        a_gpu, a_ref = np.array(list(stats.auc)), np.array(list(ref_tr.auc))
        assert np.array_equal(np.isnan(a_gpu), np.isnan(a_ref))
        ok = ~np.isnan(a_ref)
        np.testing.assert_allclose(a_gpu[ok], a_ref[ok], rtol=RTOL, atol=1e-15)
