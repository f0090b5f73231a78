"""the aggregation half of ``landmarks()`` (landmarks.jl:387-463) on the device
(``cge_b200_landmarks_aggregate``) against the host restatement of the reference's loops
(``cge_jl_b200.landmarks.aggregate_host``): bit-identical centroids, weights, d_ii, landmark
communities and weighted landmark edge lists -- integer work and FP64 sums in the reference's order,
so the bar is equality, not a tolerance."""
import numpy as np
import pytest

from cge_jl_b200 import divergence as dv
from cge_jl_b200.landmarks import aggregate_host, landmarks, split_cluster_rss
from util import clusters_of, load_fixture, planted_partition

pytestmark = pytest.mark.gpu


def _same(dev, host):
    names = ("dii", "embed", "cluster", "landmark_edges", "weights", "lweight")
    for nm, a, b in zip(names, dev, host):
        assert a.shape == b.shape, nm
        assert np.array_equal(a, b), f"{nm} differs: max |diff| {np.max(np.abs(a - b))}"


@pytest.mark.parametrize("directed", [False, True])
@pytest.mark.parametrize("n,k,d,N", [(115, 0, 0, 20), (3000, 7, 33, 150), (20000, 16, 128, 700)])
def test_device_aggregation_is_bit_identical(scorer, n, k, d, N, directed):
    if k == 0:
        edges, ew, vw, comm, emb = load_fixture("test115_weighted.npz")
    else:
        edges, ew, vw, comm, emb = planted_partition(n, k, d, seed=n + N, directed=directed,
                                                     weighted=True)
    rng = np.random.default_rng(N)
    lm = rng.integers(1, N + 1, size=emb.shape[0])
    lm[rng.permutation(emb.shape[0])[:N]] = np.arange(1, N + 1)  # every landmark non-empty
    host = aggregate_host(lm, edges, ew, vw, comm, emb, directed)
    dev = scorer.landmarks_aggregate(lm, vw, comm, emb, edges, ew, directed, N)
    _same(dev, host)
    # column-major embedding (what Julia passes) gives the same bits
    dev_f = scorer.landmarks_aggregate(lm, vw, comm, np.asfortranarray(emb), edges, ew, directed, N)
    _same(dev_f, host)


def test_landmarks_with_device_aggregation_feeds_the_scorer(scorer):
    """landmarks(..., device=scorer) == landmarks(...) on the host, and the scorer's result from the
    device-built landmark graph equals the one from the host-built graph."""
    import importlib
    lm_mod = importlib.import_module("cge_jl_b200.landmarks")
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    args = (edges, ew, vw, clusters_of(comm), comm, emb, False, 20, 1, split_cluster_rss, False)
    # selection runs on the device as well.  This fixture's clusters (5..14 vertices
    # in 32 dimensions) have rank-deficient covariances, for which LAPACK's eigenvector SIGN is not stable
    # under rounding-level differences of the matrix, so both sides use the fixed sign convention here
    # (tests/test_gpu_select.py compares the LAPACK-callback path)
    lm_mod.CANONICAL_SIGN, lm_mod.DEVICE_EIG = True, "builtin"
    try:
        host, dev = landmarks(*args), landmarks(*args, device=scorer)
    finally:
        lm_mod.CANONICAL_SIGN, lm_mod.DEVICE_EIG = False, "lapack"
    for a, b in zip(dev, host):
        assert np.array_equal(a, b)
    dii, lemb, lcomm, ledges, lw, lweight, v2l = dev
    samples = dv.draw_samples(edges, ew, 115, 800, 42, False, False)
    out = dv.wGCL(ledges, lw, lcomm, lemb, dii, lweight, vw, v2l, edges, ew, emb, False, 42, 800, False,
                  samples=samples, scorer=scorer)
    dii, lemb, lcomm, ledges, lw, lweight, v2l = host
    ref = dv.wGCL(ledges, lw, lcomm, lemb, dii, lweight, vw, v2l, edges, ew, emb, False, 42, 800, False,
                  samples=samples, scorer=scorer)
    assert np.array_equal(out, ref)


def test_out_of_range_ids_are_refused(scorer):
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    lm = np.ones(115, dtype=np.int64)
    lm[3] = 9  # beyond n_landmarks = 4
    with pytest.raises(RuntimeError, match="out of range"):
        scorer.landmarks_aggregate(lm, vw, comm, emb, edges, ew, False, 4)
