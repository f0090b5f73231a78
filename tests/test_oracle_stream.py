"""The streaming oracle (oracle/cge_oracle_stream.c: no O(n^2) arrays, all cores, q^m instead of one
pow per alpha and pair) against the line-by-line oracle of the reference: identical pass counts and
best alphas, scores within 1e-11 -- undirected and directed, cached and recomputed q."""
import numpy as np
import pytest

import oracle
from cge_jl_b200 import divergence as dv
from util import load_fixture, planted_partition


@pytest.mark.parametrize("case,directed,threads,budget", [
    ("test115", False, 1, 0), ("test115", True, 3, 0), ("pp700", False, 4, 0),
    ("pp700", True, 0, 0), ("pp700", True, 2, 1 << 30), ("pp333", False, 2, 1 << 30)])
def test_streaming_oracle_matches_the_sequential_oracle(case, directed, threads, budget):
    if case == "test115":
        edges, ew, vw, comm, emb = load_fixture("test115_weighted.npz" if directed else "test115.npz")
    else:
        n = int(case[2:])
        edges, ew, vw, comm, emb = planted_partition(n, 5, 20, seed=n + 5, directed=directed,
                                                     weighted=True)
    n = emb.shape[0]
    samples = dv.draw_samples(edges, ew, n, 1500, 42, directed, True)
    f = oracle.wgcl_directed if directed else oracle.wgcl
    ref, tr = f(edges, ew, comm, emb, np.zeros(n), vw, samples=samples)
    out, ts = oracle.wgcl_stream(edges, ew, comm, emb, vw, samples=samples, directed=directed,
                                 n_threads=threads, mem_budget=budget)
    assert ts.cached == (1 if budget else 0)
    assert ts.n_alpha_run == tr.n_alpha_run and list(ts.iters) == list(tr.iters)
    assert out[0] == ref[0] and out[4] == ref[4]
    np.testing.assert_allclose(out, ref, rtol=1e-11, atol=0)
    np.testing.assert_allclose(np.array(ts.div), np.array(tr.div), rtol=1e-11, equal_nan=True)
    np.testing.assert_allclose(np.array(ts.auc), np.array(tr.auc), rtol=1e-11, equal_nan=True)
    assert ts.hi == tr.hi and ts.lo == tr.lo == 0.0


def test_cached_and_recomputed_q_give_the_same_bits():
    edges, ew, vw, comm, emb = planted_partition(500, 4, 33, seed=9, directed=True, weighted=True)
    samples = dv.draw_samples(edges, ew, 500, 800, 42, True, True)
    a, ta = oracle.wgcl_stream(edges, ew, comm, emb, vw, samples=samples, directed=True, n_threads=3)
    b, tb = oracle.wgcl_stream(edges, ew, comm, emb, vw, samples=samples, directed=True, n_threads=3,
                               mem_budget=1 << 30)
    assert (ta.cached, tb.cached) == (0, 1)
    assert np.array_equal(a, b) and list(ta.iters) == list(tb.iters)
    assert np.array_equal(np.array(ta.div), np.array(tb.div), equal_nan=True)
