"""the device sampler of non-edges (cge_b200_sample_non_edges) against the
definition the reference samples from -- NE = all pairs minus the edge Set (divergence.jl:121-137,
405-421), drawn uniformly with replacement (:193-194, 209)."""
import numpy as np
import pytest

from cge_jl_b200 import divergence as dv
from test_gpu_parity import run_pair
from util import assert_parity, empty_landmark_args, load_fixture, planted_partition

pytestmark = pytest.mark.gpu


def _edge_codes(edges, n, directed):
    e = np.asarray(edges, dtype=np.int64)
    a, b = (e[:, 0], e[:, 1]) if directed else (e.min(axis=1), e.max(axis=1))
    return set((a * (n + 1) + b).tolist())


@pytest.mark.parametrize("directed", [False, True])
def test_samples_are_non_edges_and_deterministic(scorer, directed):
    n = 5000
    edges = planted_partition(n, 8, 4, seed=3, directed=directed)[0]
    codes = _edge_codes(edges, n, directed)
    ni, nj, draws = scorer.sample_non_edges(edges, n, 4000, n_sets=3, seed=42, directed=directed,
                                            return_draws=True)
    assert ni.shape == nj.shape == (3, 4000)
    assert ni.min() >= 1 and nj.min() >= 1 and ni.max() <= n and nj.max() <= n
    assert np.all(ni != nj)
    if not directed:
        assert np.all(ni < nj)                                   # tuples (i, j) with i < j, :122-127
    else:
        assert (ni > nj).any() and (ni < nj).any()               # ordered pairs, :407-412
    assert not (set((ni * (n + 1) + nj).ravel().tolist()) & codes)
    assert 1.0 <= draws < 1.05                                   # sparse graph: almost no rejection
    again = scorer.sample_non_edges(edges, n, 4000, n_sets=3, seed=42, directed=directed)
    assert np.array_equal(again[0], ni) and np.array_equal(again[1], nj)
    other = scorer.sample_non_edges(edges, n, 4000, n_sets=3, seed=43, directed=directed)
    assert not np.array_equal(other[0], ni)
    assert not np.array_equal(ni[0], ni[1])                      # the sets differ from each other
    # 0-based ids in, 0-based ids out, same draws
    zi, zj = scorer.sample_non_edges(edges - 1, n, 4000, n_sets=3, seed=42, directed=directed,
                                     index_base=0)
    assert np.array_equal(zi + 1, ni) and np.array_equal(zj + 1, nj)


@pytest.mark.parametrize("directed", [False, True])
def test_uniform_over_the_non_edges(scorer, directed):
    n, K = 30, 400_000
    rng = np.random.default_rng(5)
    edges = rng.integers(1, n + 1, size=(160, 2))                # duplicates and self loops included
    codes = _edge_codes(edges, n, directed)
    all_pairs = [(i, j) for i in range(1, n + 1) for j in range(1, n + 1)
                 if (i != j if directed else i < j) and i * (n + 1) + j not in codes]
    ni, nj = scorer.sample_non_edges(edges, n, K, seed=7, directed=directed)
    got, counts = np.unique(ni[0] * (n + 1) + nj[0], return_counts=True)
    assert set(got.tolist()) == {i * (n + 1) + j for i, j in all_pairs}   # every non-edge, nothing else
    expect = K / len(all_pairs)
    chi2 = float(((counts - expect) ** 2 / expect).sum())
    dof = len(all_pairs) - 1
    assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof)
    if directed:  # an edge (u, v) does not remove (v, u)
        u, v = next((a, b) for a, b in edges if a != b and b * (n + 1) + a not in codes)
        assert ((ni[0] == v) & (nj[0] == u)).any()


def test_dense_graphs(scorer):
    n = 12
    full = np.array([(i, j) for i in range(1, n + 1) for j in range(i + 1, n + 1)])
    with pytest.raises(RuntimeError, match="collection must be non-empty"):
        scorer.sample_non_edges(full, n, 10)
    ni, nj, draws = scorer.sample_non_edges(full[1:], n, 50, seed=1, return_draws=True)  # one non-edge
    assert np.all(ni == 1) and np.all(nj == 2) and draws > 20
    # the directed candidates are twice as many: the same edge list leaves half of them free
    di, dj = scorer.sample_non_edges(full, n, 500, seed=1, directed=True)
    assert np.all(di > dj)
    with pytest.raises(RuntimeError, match="outside"):
        scorer.sample_non_edges(np.array([[1, n + 1]]), n, 10)


def test_scores_with_device_drawn_negatives_match_the_oracle(scorer):
    """draw_samples(device=...) feeds the same problem struct; parity bars as everywhere else."""
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    samples = dv.draw_samples(edges, ew, 115, 1500, 42, False, True, device=scorer)
    host = dv.draw_samples(edges, ew, 115, 1500, 42, False, True)
    assert all(np.array_equal(a, b) for a, b in zip(samples[:3], host[:3]))  # positives unchanged
    out, stats, ref, tr = run_pair(scorer, False, edges, ew, comm, emb, np.zeros(115), vw,
                                   samples=samples)
    assert_parity(out, stats, ref, tr)


def test_large_graph_uses_the_device_sampler(scorer, monkeypatch):
    """Above NE_MATERIALIZE_LIMIT vertices wGCL draws its negatives through the C ABI."""
    calls = []
    real = dv.Scorer.sample_non_edges

    def spy(self, *a, **k):
        calls.append(a[1])
        return real(self, *a, **k)

    monkeypatch.setattr(dv.Scorer, "sample_non_edges", spy)
    monkeypatch.setattr(dv, "NE_MATERIALIZE_LIMIT", 500)
    n = 700
    edges, ew, vw, comm, emb = planted_partition(n, 5, 20, seed=705)
    out = dv.wGCL(edges, ew, comm, emb, np.zeros(n), vw, *empty_landmark_args(), False, 42, 2000,
                  False, scorer=scorer)
    assert calls == [n] and np.all(np.isfinite(out)) and 0.0 < out[5] < 1.0
