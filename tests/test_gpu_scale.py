"""GPU: sizes the oracle cannot reach in seconds -- parity through size-independent properties
(SURVEY.md section 7): the fixed point converged for the last alpha (max |w - S| <= 0.001 on the
probe of the final S), all drivers and both regimes agree with each other, pads and ragged last
tiles do not leak (n deliberately not a multiple of 128)."""
import numpy as np
import pytest

from cge_jl_b200 import divergence as dv
from cge_jl_b200.synth import planted_partition
from util import empty_landmark_args

pytestmark = pytest.mark.gpu
EMPTY = empty_landmark_args()


def _run(scorer, directed, data, samples, driver, regime, n):
    edges, ew, vw, comm, emb = data
    f = dv.wGCL_directed if directed else dv.wGCL
    out, st = f(edges, ew, comm, emb, np.zeros(n), vw, *EMPTY, False, 42, samples[0].shape[1],
                False, samples=samples, return_stats=True, scorer=scorer, driver=driver,
                regime=regime)
    return out, st


@pytest.mark.parametrize("directed", [False, True])
def test_20k_properties_and_driver_agreement(scorer, directed):
    n = 20011
    data = planted_partition(n, k=24, d=48, seed=31, directed=directed, weighted=directed)
    edges, ew, vw = data[0], data[1], data[2]
    samples = dv.draw_samples(edges, ew, n, 5000, 42, directed, True)
    out, st = _run(scorer, directed, data, samples, 2, 1, n)
    assert np.all(np.isfinite(out)) and out[0] * 4 == round(out[0] * 4)
    assert st.n_alpha_run >= 6 and min(list(st.iters)[: st.n_alpha_run]) >= 1
    # converged degrees on the final probe
    if directed:
        din, dout = np.zeros(n), np.zeros(n)
        np.add.at(dout, edges[:, 0] - 1, ew)
        np.add.at(din, edges[:, 1] - 1, ew)
        sin, sout = scorer.debug_read(3, n), scorer.debug_read(4, n)
        assert np.abs(din - sin)[din > 0].max() <= 0.001
        assert np.abs(dout - sout)[dout > 0].max() <= 0.001
    else:
        assert np.abs(vw - scorer.debug_read(3, n)).max() <= 0.001
    assert np.isclose(out[6], 1.96 * np.sqrt(out[5] * (1 - out[5]) / 5000))
    # the host-loop driver walks through the same passes and lands on the same numbers
    out1, st1 = _run(scorer, directed, data, samples, 1, 1, n)
    assert list(st1.iters) == list(st.iters)
    assert out1[0] == out[0] and out1[4] == out[4]
    np.testing.assert_allclose(out1, out, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("directed,d", [(False, 40), (True, 40), (False, 12), (True, 24)])
def test_stored_and_recompute_regimes_agree(scorer, directed, d):
    """300 tiles: every CTA walks 2-3 tiles, so the recompute kernel's cross-tile staging (the
    next tile's first chunk requested during the last chunk of the current one) runs with 1, 2
    and 3 chunks per tile.  Two vertices share an embedding row (distance exactly 0)."""
    n = 3001
    data = planted_partition(n, k=9, d=d, seed=77, directed=directed, weighted=True)
    data[4][1234] = data[4][77]
    samples = dv.draw_samples(data[0], data[1], n, 3000, 42, directed, True)
    a, sa = _run(scorer, directed, data, samples, 2, 1, n)
    b, sb = _run(scorer, directed, data, samples, 2, 4, n)  # difference form
    assert sa.regime == 1 and sb.regime == 4 and sb.matrix_bytes == 0
    assert list(sa.iters) == list(sb.iters) and a[0] == b[0] and a[4] == b[4]
    np.testing.assert_allclose(b, a, rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(np.array(list(sb.div)), np.array(list(sa.div)), rtol=1e-11,
                               equal_nan=True)


@pytest.mark.parametrize("directed,d", [(False, 40), (True, 24)])
def test_stored_and_row_norm_dot_form_agree(scorer, directed, d):
    """Regime 3 against the stored regime on 300 tiles, with an exactly duplicated and a 1e-9-close
    embedding row (the pairs the dot form hands back to the difference form)."""
    n = 3001
    data = planted_partition(n, k=9, d=d, seed=77, directed=directed, weighted=True)
    data[4][1234] = data[4][77]
    data[4][55] = data[4][56] + 1e-9
    samples = dv.draw_samples(data[0], data[1], n, 3000, 42, directed, True)
    a, sa = _run(scorer, directed, data, samples, 2, 1, n)
    b, sb = _run(scorer, directed, data, samples, 2, 3, n)
    assert sa.regime == 1 and sb.regime == 3 and sb.matrix_bytes == 0
    assert list(sa.iters) == list(sb.iters) and a[0] == b[0] and a[4] == b[4]
    np.testing.assert_allclose(b, a, rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(np.array(list(sb.div)), np.array(list(sa.div)), rtol=1e-11,
                               equal_nan=True)


def test_q_matrix_symmetric_and_in_range_at_scale(scorer):
    n = 2500
    data = planted_partition(n, k=6, d=24, seed=5)
    dv.wGCL(data[0], data[1], data[3], data[4], np.zeros(n), data[2], *EMPTY, False, 42, 0, False,
            samples=None, scorer=scorer, max_alphas=1)
    q = scorer.debug_read(0, n)
    assert np.array_equal(q, q.T)
    assert q.min() == 0.0 and q.max() == 1.0 and np.all(np.diag(q) == 1.0)
    emb = data[4]
    i, j = np.random.default_rng(0).integers(0, n, size=(2, 2000))
    d = np.sqrt(((emb[i] - emb[j]) ** 2).sum(1))
    sq = ((emb ** 2).sum(1))
    D2 = sq[:, None] + sq[None, :] - 2 * emb @ emb.T
    hi = np.sqrt(D2.max())
    want = (1 - d / hi) ** 0.25
    ok = i != j
    np.testing.assert_allclose(q[i[ok], j[ok]], want[ok], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("n,d,k", [(3001, 40, 9), (5000, 128, 16), (700, 7, 3)])
def test_tensor_core_diameter_filter_is_exact(scorer, monkeypatch, n, d, k):
    """Landmark mode needs the exact diameter of the ORIGINAL graph (divergence.jl:113).  The
    tcgen05 filter + FP64 verification must return the very same double as the all-FP64 pass, and
    only a small share of the tiles may need verification."""
    from cge_jl_b200.landmarks import landmarks, split_cluster_rss
    from util import clusters_of
    edges, ew, vw, comm, emb = planted_partition(n, k=k, d=d, seed=n + d)
    emb[17] *= 1.5  # an outlier row pair decides the diameter
    lm = landmarks(edges, ew, vw, clusters_of(comm), comm, emb, False, 4 * k, 4, split_cluster_rss,
                   False)
    dii, lemb, lcomm, ledges, lw, lweight, v2l = lm
    samples = dv.draw_samples(edges, ew, n, 2000, 42, False, False)
    res = {}
    for label, dmin in (("fp64", "1000000000"), ("filter", "0")):
        monkeypatch.setenv("CGE_B200_DIAM_MIN", dmin)
        out, st = dv.wGCL(ledges, lw, lcomm, lemb, dii, lweight, vw, v2l, edges, ew, emb, False, 42,
                          2000, False, samples=samples, return_stats=True, scorer=scorer)
        res[label] = (out, float(st.hi_full), int(st.diam_candidate_tiles), int(st.launches))
    assert res["fp64"][2] == -1 and res["filter"][2] >= 1
    nb = (n + 127) // 128
    assert res["filter"][2] <= max(4, nb * (nb + 1) // 2 // 8)
    assert res["filter"][1] == res["fp64"][1]                      # bit-identical diameter
    assert np.array_equal(res["filter"][0], res["fp64"][0])
    D2 = ((emb[:, None, :] - emb[None, :, :]) ** 2).sum(-1) if n <= 3001 else None
    if D2 is not None:
        assert np.isclose(res["filter"][1], np.sqrt(D2.max()), rtol=1e-14)


def test_diameter_filter_with_a_large_common_offset(scorer, monkeypatch):
    """A common offset of the embedding inflates the norms but not the distances: the packed copy
    is centred, so the filter still prunes and the diameter stays exact."""
    from cge_jl_b200.landmarks import landmarks, split_cluster_rss
    from util import clusters_of
    n = 2000
    edges, ew, vw, comm, emb = planted_partition(n, k=5, d=32, seed=3)
    emb = emb + 1000.0
    lm = landmarks(edges, ew, vw, clusters_of(comm), comm, emb, False, 20, 4, split_cluster_rss, False)
    dii, lemb, lcomm, ledges, lw, lweight, v2l = lm
    samples = dv.draw_samples(edges, ew, n, 500, 42, False, False)
    monkeypatch.setenv("CGE_B200_DIAM_MIN", "0")
    out, st = dv.wGCL(ledges, lw, lcomm, lemb, dii, lweight, vw, v2l, edges, ew, emb, False, 42, 500,
                      False, samples=samples, return_stats=True, scorer=scorer)
    D2 = ((emb[:, None, :] - emb[None, :, :]) ** 2).sum(-1)
    assert np.isclose(st.hi_full, np.sqrt(D2.max()), rtol=1e-12)
    assert 1 <= st.diam_candidate_tiles <= 16


def test_abcd_like_unequal_communities(scorer):
    """ABCD-style input (BASELINE config 4's family): power-law degrees and very unequal community
    sizes (2 .. thousands of vertices) -- community boundaries fall inside tiles everywhere."""
    from cge_jl_b200.synth import abcd_like
    n = 20011
    data = abcd_like(n, k=64, d=32, seed=4)
    edges, ew, vw = data[0], data[1], data[2]
    samples = dv.draw_samples(edges, ew, n, 5000, 42, False, True)
    out, st = _run(scorer, False, data, samples, 2, 1, n)
    assert np.all(np.isfinite(out)) and st.n_alpha_run >= 6
    assert np.abs(vw - scorer.debug_read(3, n)).max() <= 0.001
    out1, st1 = _run(scorer, False, data, samples, 1, 1, n)
    assert list(st1.iters) == list(st.iters)
    np.testing.assert_allclose(out1, out, rtol=1e-12, atol=1e-15)
    # the B matrix bins survive the irregular community layout: recompute regime agrees too
    small = abcd_like(2500, k=40, d=16, seed=9)
    s2 = dv.draw_samples(small[0], small[1], 2500, 2000, 42, False, True)
    a, sa = _run(scorer, False, small, s2, 2, 1, 2500)
    b, sb = _run(scorer, False, small, s2, 2, 2, 2500)
    assert list(sa.iters) == list(sb.iters)
    np.testing.assert_allclose(b, a, rtol=1e-11, atol=1e-15)


def test_sub_block_spot_check_of_a_recompute_run(scorer, monkeypatch):
    """What scripts/run_config.py --config 5 --exact does at 10^6 vertices (SURVEY.md section 7 "(iv)
    sub-block spot checks at 1M"), at a size the suite can afford: after two alphas of a recompute run
    with super-tiles and part of the matrix kept in HBM, the degree sums S_i of a few vertices are
    recomputed from scratch in NumPy over all their partners and must match the device's last pass."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from spotcheck import degree_sum_spot_check
    monkeypatch.setenv("CGE_B200_RC_SB", "4")
    monkeypatch.setenv("CGE_B200_STORE_MB", "600")
    n = 30000
    data = planted_partition(n, k=12, d=64, seed=31)
    samples = dv.draw_samples(data[0], data[1], n, 2000, 42, False, True)
    out, st = dv.wGCL(data[0], data[1], data[3], data[4], np.zeros(n), data[2], *EMPTY, False, 42, 2000,
                      False, samples=samples, return_stats=True, scorer=scorer, max_alphas=2, regime=2)
    assert st.regime == 2 and 0 < st.matrix_bytes < 8 * n * n // 2 and st.n_alpha_run == 2
    rows = np.random.default_rng(3).integers(0, n, size=6)
    worst = degree_sum_spot_check(data[4], scorer.debug_read(1, n), scorer.debug_read(3, n), data[2],
                                  float(st.hi), 0.5, rows)
    assert worst <= 1e-11
