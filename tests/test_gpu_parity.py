"""GPU parity tests: libcge_b200.so (through the C ABI) against the CPU oracle on identical
inputs and identical sampled pairs.  Bars (north_star): best alpha identical, fixed-point pass
counts identical, scores within 1e-9 relative."""
import numpy as np
import pytest

import oracle
from cge_jl_b200 import divergence as dv
from cge_jl_b200.landmarks import landmarks, split_cluster_rss
from util import (RTOL, assert_parity, clusters_of, empty_landmark_args, load_fixture,
                  planted_partition)

pytestmark = pytest.mark.gpu
EMPTY = empty_landmark_args()
README_GOLDEN = [6.25, 0.002961243353776198]  # /root/reference/README.md:99, elements 1-2


def run_pair(scorer, directed, edges, ew, comm, emb, dist, vw, lm_args=None, split=False,
             seed=42, K=2000, samples="draw", driver=0, regime=0):
    """Score on the GPU and with the oracle using the same sampled pairs."""
    if lm_args is None:
        init_vw, v2l, init_edges, init_ew, init_emb = EMPTY
        adj = (edges, ew, int(edges.max()))
    else:
        init_vw, v2l, init_edges, init_ew, init_emb = lm_args
        adj = (init_edges, init_ew, init_vw.shape[0])
    if samples == "draw":
        samples = dv.draw_samples(adj[0], adj[1], adj[2], K, seed, directed, exact=lm_args is None)
    f_gpu = dv.wGCL_directed if directed else dv.wGCL
    out, stats = f_gpu(edges, ew, comm, emb, dist, vw, init_vw, v2l, init_edges, init_ew,
                       init_emb, split, seed, K, False, samples=samples, return_stats=True,
                       scorer=scorer, driver=driver, regime=regime)
    assert driver == 0 or stats.driver == driver or regime >= 2
    assert stats.regime == (regime or 1)
    f_ref = oracle.wgcl_directed if directed else oracle.wgcl
    ref, tr = f_ref(edges, ew, comm, emb, dist, vw, init_vw if lm_args else None,
                    v2l if lm_args else None, init_emb if lm_args else None, split, samples)
    return out, stats, ref, tr


def test_q_matrix_matches_definition(scorer):
    """The stored matrix is q = (1 - (D-lo)/(hi-lo))^(1/4) with D from auxilary.jl:14-20."""
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    n = 115
    out, stats = dv.wGCL(edges, ew, comm, emb, np.zeros(n), vw, *EMPTY, False, 42, 0, False,
                         samples=None, return_stats=True, scorer=scorer, max_alphas=1)
    q = scorer.debug_read(0, n)
    D = np.sqrt(((emb[:, None, :] - emb[None, :, :]) ** 2).sum(-1))
    np.fill_diagonal(D, 0.0)
    lo, hi = D.min(), D.max()
    assert np.isclose(stats.lo, lo, rtol=1e-15, atol=0) and np.isclose(stats.hi, hi, rtol=1e-14)
    want = (1.0 - (D - lo) / (hi - lo)) ** 0.25
    np.testing.assert_allclose(q, want, rtol=1e-12, atol=1e-12)
    assert np.array_equal(q, q.T)


def test_one_pass_degrees(scorer):
    """After the passes of alpha = 0.25 the S vector equals T_i * sum_j T_j * q_ij for the T of
    the previous pass -- checked through the invariant of the final pass: |w - S| <= 0.001."""
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    n = 115
    dv.wGCL(edges, ew, comm, emb, np.zeros(n), vw, *EMPTY, False, 42, 0, False, samples=None,
            scorer=scorer, max_alphas=1)
    S = scorer.debug_read(3, n)
    assert np.max(np.abs(vw - S)) <= 0.001
    T = scorer.debug_read(1, n)
    assert np.all(T > 0)


DRIVERS = pytest.mark.parametrize("driver", [1, 2, 3], ids=["hostloop", "persistent", "ring"])


@DRIVERS
def test_exact_undirected_test_graph(scorer, driver):
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    out, stats, ref, tr = run_pair(scorer, False, edges, ew, comm, emb, np.zeros(115), vw,
                                   driver=driver)
    assert_parity(out, stats, ref, tr)
    assert out[0] == 3.25 and np.isclose(out[1], 0.006929334486296551, rtol=RTOL)


def test_exact_undirected_split_global(scorer):
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    out, stats, ref, tr = run_pair(scorer, False, edges, ew, comm, emb, np.zeros(115), vw,
                                   split=True)
    assert_parity(out, stats, ref, tr)
    assert out[2] > 0 and out[3] > 0
    assert np.isclose(out[1], (out[2] + out[3]) / 2, rtol=1e-12)


@DRIVERS
def test_exact_directed_weighted_test_graph(scorer, driver):
    edges, ew, vw, comm, emb = load_fixture("test115_weighted.npz")
    out, stats, ref, tr = run_pair(scorer, True, edges, ew, comm, emb, np.zeros(115), vw,
                                   driver=driver)
    assert_parity(out, stats, ref, tr)
    assert out[0] == 5.5 and np.isclose(out[1], 0.008821457041054588, rtol=RTOL)


def test_exact_directed_split_global(scorer):
    edges, ew, vw, comm, emb = load_fixture("test115_weighted.npz")
    out, stats, ref, tr = run_pair(scorer, True, edges, ew, comm, emb, np.zeros(115), vw,
                                   split=True)
    assert_parity(out, stats, ref, tr)


def test_unseeded_sample_sets_per_alpha(scorer):
    """seed = -1: a fresh sample set per alpha (n_sets = 40), divergence.jl:184 `seed != -1 &&`."""
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    samples = dv.draw_samples(edges, ew, 115, 500, -1, False, True)
    assert samples[0].shape == (40, 500)
    out, stats, ref, tr = run_pair(scorer, False, edges, ew, comm, emb, np.zeros(115), vw,
                                   seed=-1, K=500, samples=samples)
    assert_parity(out, stats, ref, tr)


@pytest.mark.parametrize("directed", [False, True])
def test_landmark_mode_test_graph(scorer, directed):
    """runtests.jl:4-7,95-96 configuration (-l 20 -f 1 -m rss) through landmarks() and wGCL."""
    edges, ew, vw, comm, emb = load_fixture("test115_weighted.npz" if directed else "test115.npz")
    dii, lemb, lcomm, ledges, lw, lweight, v2l = landmarks(
        edges, ew, vw, clusters_of(comm), comm, emb, False, 20, 1, split_cluster_rss, directed)
    out, stats, ref, tr = run_pair(scorer, directed, ledges, lw, lcomm, lemb, dii, lweight,
                                   lm_args=(vw, v2l, edges, ew, emb))
    assert_parity(out, stats, ref, tr)
    assert out[0] <= 10.0  # the reference's own assertion, runtests.jl:101
    assert np.isclose(stats.hi_full, tr.hi_full, rtol=1e-15)


@DRIVERS
@pytest.mark.parametrize("directed,n,k,d", [(False, 700, 5, 20), (True, 520, 7, 33),
                                            (False, 129, 3, 16), (False, 256, 40, 8)])
def test_synthetic_multi_tile(scorer, directed, n, k, d, driver):
    """Several 128-tiles, ragged last tile, d not a multiple of 16, many small communities."""
    edges, ew, vw, comm, emb = planted_partition(n, k, d, seed=n + k, directed=directed,
                                                 weighted=True)
    out, stats, ref, tr = run_pair(scorer, directed, edges, ew, comm, emb, np.zeros(n), vw, K=3000,
                                   driver=driver)
    assert_parity(out, stats, ref, tr)


@pytest.mark.parametrize("regime", [2, 4], ids=["dot", "diff"])
@pytest.mark.parametrize("driver", [1, 2], ids=["hostloop", "persistent"])
@pytest.mark.parametrize("directed,n,k,d", [(False, 115, 0, 0), (True, 115, 0, 0),
                                            (False, 700, 5, 20), (True, 520, 7, 33)])
def test_recompute_regime(scorer, directed, n, k, d, driver, regime):
    """No stored matrix: distances are re-derived from the embedding in every pass (north_star
    kernel (a)), from the row norms and dot products of the centred embedding (regime 2, the
    default) or in the reference's difference form (regime 4); same parity bars against the oracle
    as the stored regime."""
    if k == 0:
        edges, ew, vw, comm, emb = load_fixture("test115_weighted.npz" if directed else "test115.npz")
    else:
        edges, ew, vw, comm, emb = planted_partition(n, k, d, seed=n + k, directed=directed,
                                                     weighted=True)
    out, stats, ref, tr = run_pair(scorer, directed, edges, ew, comm, emb, np.zeros(n), vw,
                                   driver=driver, regime=regime)
    assert stats.matrix_bytes == 0
    assert_parity(out, stats, ref, tr)


@pytest.mark.parametrize("sb", [2, 4, 8])
@pytest.mark.parametrize("directed,driver", [(False, 2), (True, 2), (False, 1)])
def test_recompute_super_tiles(scorer, directed, driver, sb, monkeypatch):
    """Super-tiles of sb x sb tiles (what 10^5..10^6 vertices use to keep the partial-sum slots
    small) forced on a 1 100-vertex graph: 9 tile rows make whole, ragged and diagonal super-tiles;
    d = 70 gives 5 staging chunks per tile, more than the 4 ring stages."""
    monkeypatch.setenv("CGE_B200_RC_SB", str(sb))
    n = 1100
    edges, ew, vw, comm, emb = planted_partition(n, 6, 70, seed=sb + 5 * directed, directed=directed,
                                                 weighted=True)
    out, stats, ref, tr = run_pair(scorer, directed, edges, ew, comm, emb, np.zeros(n), vw,
                                   driver=driver, regime=2)
    assert stats.matrix_bytes == 0 and stats.regime == 2
    assert_parity(out, stats, ref, tr)


@pytest.mark.parametrize("directed,n,k,d", [(False, 115, 0, 0), (False, 700, 5, 20),
                                            (True, 520, 7, 33)])
def test_recompute_row_norm_dot_form(scorer, directed, n, k, d):
    """Regime 3: the recompute regime with d^2 = n_i + n_j - 2 x_i.x_j on the centred embedding
    (extrema and sampled pairs in the same arithmetic); same parity bars against the oracle."""
    if k == 0:
        edges, ew, vw, comm, emb = load_fixture("test115.npz")
    else:
        edges, ew, vw, comm, emb = planted_partition(n, k, d, seed=n + k, directed=directed,
                                                     weighted=True)
    out, stats, ref, tr = run_pair(scorer, directed, edges, ew, comm, emb, np.zeros(n), vw,
                                   driver=2, regime=3)
    assert stats.matrix_bytes == 0 and stats.regime == 3
    assert_parity(out, stats, ref, tr)


@pytest.mark.parametrize("sb,store_mb", [(1, 1), (2, 2), (2, 100)])
@pytest.mark.parametrize("directed", [False, True])
def test_recompute_store_what_fits(scorer, directed, sb, store_mb, monkeypatch):
    """"Store what fits": the leading super-tiles of the rank keep their q tiles in HBM (here a
    budget of 1-2 MB = 8-16 of the 45 tiles, or everything) and are read in every pass, the rest is
    recomputed; the mix must meet the same parity bars."""
    monkeypatch.setenv("CGE_B200_RC_SB", str(sb))
    monkeypatch.setenv("CGE_B200_STORE_MB", str(store_mb))
    n = 1100
    edges, ew, vw, comm, emb = planted_partition(n, 6, 40, seed=31 + directed, directed=directed,
                                                 weighted=True)
    out, stats, ref, tr = run_pair(scorer, directed, edges, ew, comm, emb, np.zeros(n), vw,
                                   driver=2, regime=2)
    assert stats.regime == 2
    tiles = stats.matrix_bytes // 131072
    assert stats.matrix_bytes % 131072 == 0 and (tiles == 45 if store_mb == 100 else 0 < tiles <= 8 * store_mb)
    assert_parity(out, stats, ref, tr)


def test_recompute_branch_free_math_stays_within_2_ulp(scorer):
    """The recompute epilogue's short branch-free forms -- sqrt from the MUFU.RSQ64H seed and one cubic
    step, (hi - D) * 1/(hi - lo) for the normalisation -- against the correctly rounded sqrt and
    divide the stored regime and the reference use: 2^26 pseudo-random operands in the epilogue's
    ranges, none farther than 2 ulp, zeros and operands under 2^-943 exactly 0, D = hi exactly 0."""
    assert scorer.selftest_math(1 << 26, seed=2024) == (0, 0)


def test_recompute_regime_landmarks_with_diagonal(scorer):
    """Landmark mode has a non-zero diagonal (d_ii) and lo > 0 possible: recompute vs stored."""
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    dii, lemb, lcomm, ledges, lw, lweight, v2l = landmarks(
        edges, ew, vw, clusters_of(comm), comm, emb, False, 20, 1, split_cluster_rss, False)
    out, stats, ref, tr = run_pair(scorer, False, ledges, lw, lcomm, lemb, dii, lweight,
                                   lm_args=(vw, v2l, edges, ew, emb), regime=2)
    assert_parity(out, stats, ref, tr)


def test_sample_sets_must_cover_the_alpha_grid(scorer):
    """n_sets is 1 (seeded: the same pairs for every alpha) or one set per evaluated alpha; anything
    in between would index past the sample buffers from alpha n_sets + 1 on and is refused."""
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    n = vw.shape[0]
    samples = dv.draw_samples(edges, ew, n, 300, -1, False, True)  # 40 sets
    two = tuple(a[:2] for a in samples)
    p, keep = dv.make_problem(edges, ew, comm, emb, np.zeros(n), vw, None, None, None, False, False,
                              two)
    with pytest.raises(RuntimeError, match="n_sets"):
        scorer.upload(p, keep)
    p, keep = dv.make_problem(edges, ew, comm, emb, np.zeros(n), vw, None, None, None, False, False,
                              two, max_alphas=2)  # two sets do cover a two-alpha run
    scorer.upload(p, keep)
    out, st = scorer.run()
    assert st.n_alpha_run == 2 and np.all(np.isfinite(out))


def test_star_graph_early_exit(scorer):
    n = 6
    edges = np.array([[1, j] for j in range(2, n + 1)])
    out = dv.wGCL_directed(edges, np.ones(n - 1), np.ones((n, 1), dtype=np.int64),
                           np.random.default_rng(1).normal(size=(n, 4)), np.zeros(n), np.ones(n),
                           *EMPTY, False, 42, 100, False, scorer=scorer)
    assert out.tolist() == [-1.0, 0, 0, 0, 0, 0]  # divergence.jl:332-334


def test_assertions_match_reference(scorer):
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    with pytest.raises(AssertionError, match="No. communities not matching no. vertices"):
        dv.wGCL(edges, ew, comm[:-1], emb, np.zeros(115), vw, *EMPTY, False, 42, 10, False,
                scorer=scorer)
    with pytest.raises(AssertionError, match="Distances vector length"):
        dv.wGCL(edges, ew, comm, emb, np.zeros(114), vw, *EMPTY, False, 42, 10, False,
                scorer=scorer)
    # the same checks inside the library (what a Julia ccall would hit)
    p, keep = dv.make_problem(edges, ew, comm[:-1], emb, np.zeros(115), vw, None, None, None,
                              False, False, None)
    with pytest.raises(AssertionError, match="No. communities"):
        scorer.upload(p, keep)


def test_readme_golden_landmarks_10k(scorer):
    """BASELINE.json configs[0]: 10k example, -l 200 --seed 42 (README.md:88-100)."""
    edges, ew, vw, comm, emb = load_fixture("example10k.npz")
    dii, lemb, lcomm, ledges, lw, lweight, v2l = landmarks(
        edges, ew, vw, clusters_of(comm), comm, emb, False, 200, 4, split_cluster_rss, False)
    out, stats, ref, tr = run_pair(scorer, False, ledges, lw, lcomm, lemb, dii, lweight,
                                   lm_args=(vw, v2l, edges, ew, emb), K=10000)
    assert out[0] == README_GOLDEN[0]
    assert abs(out[1] - README_GOLDEN[1]) / README_GOLDEN[1] < RTOL
    assert_parity(out, stats, ref, tr)
    # elements 5-7 depend on Julia's RNG (parity unpinned); they must be statistically consistent
    assert 0.0 <= out[5] < 0.01 and out[4] >= 7.5


@DRIVERS
def test_exact_10k_against_frozen_oracle(scorer, driver):
    """BASELINE.json configs[1]: 10k example --force-exact, vs tests/golden/oracle_example10k_exact.npz."""
    import os
    from util import GOLDEN
    g = np.load(os.path.join(GOLDEN, "oracle_example10k_exact.npz"))
    edges, ew, vw, comm, emb = load_fixture("example10k.npz")
    samples = tuple(g[k].astype(np.int64) if k != "pos_w" else g[k]
                    for k in ("pos_i", "pos_j", "pos_w", "neg_i", "neg_j"))
    out, stats = dv.wGCL(edges, ew, comm, emb, np.zeros(10000), vw, *EMPTY, False, 42, 10000,
                         False, samples=samples, return_stats=True, scorer=scorer, driver=driver)

    class Tr:
        n_alpha_run = int(g["n_alpha_run"])
        iters, div, auc = g["iters"].tolist(), g["div"].tolist(), g["auc"].tolist()
    assert_parity(out, stats, g["out"], Tr)
    assert int(stats.fp_sweeps) == 1026 and out[0] == 10.0  # SURVEY section 6 probe
    assert np.isclose(stats.lo, float(g["lo"])) and np.isclose(stats.hi, float(g["hi"]), rtol=1e-14)


# ---------------------------------------------------------------------------------------------
# edge cases (the reference's tests cover none of these; the oracle defines the answer)
# ---------------------------------------------------------------------------------------------
def _tiny(n, edges, d=3, seed=0, k=2, w=None):
    rng = np.random.default_rng(seed)
    edges = np.asarray(edges, dtype=np.int64)
    ew = np.ones(len(edges)) if w is None else np.asarray(w, dtype=np.float64)
    vw = np.zeros(n)
    np.add.at(vw, edges[:, 0] - 1, ew)
    np.add.at(vw, edges[:, 1] - 1, ew)
    comm = (np.arange(n) % k + 1).reshape(-1, 1)
    return edges, ew, vw, comm, rng.normal(size=(n, d))


@pytest.mark.parametrize("driver", [1, 2], ids=["hostloop", "persistent"])
def test_tiny_graphs(scorer, driver):
    """Three and four vertices: a single ragged tile, one community pair."""
    for n, e in ((3, [[1, 2], [2, 3]]), (4, [[1, 2], [2, 3], [3, 4]])):
        edges, ew, vw, comm, emb = _tiny(n, e)
        out, stats, ref, tr = run_pair(scorer, False, edges, ew, comm, emb, np.zeros(n), vw, K=7,
                                       driver=driver)
        assert_parity(out, stats, ref, tr)


def test_self_loops_multi_edges_and_single_sample(scorer):
    """Self-loops stay in E (divergence.jl:131-134 keeps (i,i,w)), duplicate edges are sampled with
    multiplicity, K = 1."""
    e = [[1, 2], [2, 3], [3, 4], [4, 5], [5, 1], [2, 2], [1, 2], [3, 5]]
    edges, ew, vw, comm, emb = _tiny(5, e, w=[1, 2, 1, 0.5, 1, 3, 1, 1])
    for K in (1, 33):
        out, stats, ref, tr = run_pair(scorer, False, edges, ew, comm, emb, np.zeros(5), vw, K=K)
        assert_parity(out, stats, ref, tr)


def test_duplicate_embedding_rows_and_isolated_vertex(scorer):
    """Identical rows give D = 0 off the diagonal (q = 1); a vertex below max(edges) without
    edges has weight 0 and its T decays (divergence.jl:162 with vweights[i] = 0)."""
    edges, ew, vw, comm, emb = planted_partition(150, 3, 5, seed=9)
    edges = edges[(edges[:, 0] != 7) & (edges[:, 1] != 7)]      # vertex 7 becomes isolated
    ew = np.ones(edges.shape[0])
    vw = np.zeros(150)
    np.add.at(vw, edges[:, 0] - 1, ew)
    np.add.at(vw, edges[:, 1] - 1, ew)
    assert vw[6] == 0 and edges.max() == 150
    emb[10] = emb[11] = emb[140]
    out, stats, ref, tr = run_pair(scorer, False, edges, ew, comm, emb, np.zeros(150), vw, K=500)
    assert_parity(out, stats, ref, tr)


def test_directed_vertices_without_in_or_out_edges(scorer):
    """Tin/Tout start at 0 where the degree is 0 and are never updated (divergence.jl:399-402,453,457)."""
    e = [[1, 2], [1, 3], [2, 3], [3, 4], [4, 2], [5, 1], [5, 4], [2, 6]]
    edges, ew, vw, comm, emb = _tiny(6, e, d=4, seed=3, w=[1, 1.5, 2, 1, 1, 0.7, 1, 1])
    out, stats, ref, tr = run_pair(scorer, True, edges, ew, comm, emb, np.zeros(6), vw, K=50)
    assert out.shape == (7,)
    assert_parity(out, stats, ref, tr)


def test_column_major_embedding_is_accepted_without_copy(scorer):
    """Julia passes a column-major Matrix{Float64}; strides are part of the ABI."""
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    samples = dv.draw_samples(edges, ew, 115, 300, 42, False, True)
    a = dv.wGCL(edges, ew, comm, emb, np.zeros(115), vw, *EMPTY, False, 42, 300, False,
                samples=samples, scorer=scorer)
    emb_f = np.asfortranarray(emb)
    assert emb_f.strides == (8, 8 * 115)
    b = dv.wGCL(edges, ew, comm, emb_f, np.zeros(115), vw, *EMPTY, False, 42, 300, False,
                samples=samples, scorer=scorer)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("split", [False, True])
def test_deferred_b_sweep_is_the_same_computation(scorer, split, monkeypatch):
    """The B sweep of alpha fused into the first fixed-point pass of alpha + 1/4 (k_bfp): identical passes
    and T bit for bit (the degree-sum arithmetic is tile_pass_u's), per-alpha global scores equal to the
    B atomics' 1e-13, oracle parity -- on a problem whose patience counters switch the local score off
    early, so deferred, stand-alone and skipped B sweeps all occur."""
    n = 900
    edges, ew, vw, comm, emb = planted_partition(n, 6, 16, seed=41, weighted=True)
    samples = dv.draw_samples(edges, ew, n, 800, 42, False, True)
    monkeypatch.setenv("CGE_B200_RT_EXPONENT", "0")  # the per-alpha kernels, as on large problems
    runs = {}
    for fuse in ("0", "1"):
        monkeypatch.setenv("CGE_B200_FUSE_B", fuse)
        out, st = dv.wGCL(edges, ew, comm, emb, np.zeros(n), vw, *EMPTY, split, 42, 800, False,
                          samples=samples, return_stats=True, scorer=scorer, driver=2, regime=1)
        runs[fuse] = (out, st, scorer.debug_read(1, n))
    (a, sa, ta), (b, sb, tb) = runs["0"], runs["1"]
    assert sa.b_fused == 0 and sb.b_fused >= 5 and sb.b_sweeps == sa.b_sweeps
    assert list(sa.iters) == list(sb.iters) and np.array_equal(ta, tb)
    assert np.array_equal(a[4:], b[4:]) and a[0] == b[0]
    np.testing.assert_allclose(b, a, rtol=1e-12)
    np.testing.assert_allclose(np.array(list(sb.div)), np.array(list(sa.div)), rtol=1e-12, equal_nan=True)
    ref, tr = oracle.wgcl(edges, ew, comm, emb, np.zeros(n), vw, samples=samples, split=split)
    assert_parity(b, sb, ref, tr)


def test_run_is_bit_reproducible(scorer):
    """Fixed reduction orders everywhere in the fixed point: repeated runs agree bit for bit on
    everything that does not pass through the B atomics, and to 1e-13 on the global score."""
    edges, ew, vw, comm, emb = planted_partition(700, 5, 20, seed=705, weighted=True)
    samples = dv.draw_samples(edges, ew, 700, 1000, 42, False, True)
    runs = [dv.wGCL(edges, ew, comm, emb, np.zeros(700), vw, *EMPTY, False, 42, 1000, False,
                    samples=samples, return_stats=True, scorer=scorer) for _ in range(3)]
    for out, st in runs[1:]:
        assert list(st.iters) == list(runs[0][1].iters)
        assert np.array_equal(out[4:], runs[0][0][4:]) and out[0] == runs[0][0][0]
        np.testing.assert_allclose(out[1], runs[0][0][1], rtol=1e-13)
    T = scorer.debug_read(1, 700)
    dv.wGCL(edges, ew, comm, emb, np.zeros(700), vw, *EMPTY, False, 42, 1000, False,
            samples=samples, scorer=scorer)
    assert np.array_equal(T, scorer.debug_read(1, 700))


@pytest.mark.parametrize("seed", range(12))
def test_random_small_graphs(scorer, seed):
    """Randomised parity sweep: size, dimension, community count, weights, direction, split flag and
    driver all vary with the seed; every case must meet the full parity bars against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(20, 420))
    k = int(rng.integers(1, min(n // 4, 30) + 1))
    d = int(rng.integers(2, 70))
    directed = bool(seed % 2)
    edges, ew, vw, comm, emb = planted_partition(n, k, d, seed=seed, directed=directed,
                                                 weighted=bool(rng.integers(0, 2)),
                                                 deg=int(rng.integers(4, 12)))
    out, stats, ref, tr = run_pair(scorer, directed, edges, ew, comm, emb, np.zeros(n), vw,
                                   split=bool(rng.integers(0, 2)), K=int(rng.integers(50, 1500)),
                                   driver=int(rng.integers(1, 4)), regime=int(rng.integers(1, 3)))
    assert_parity(out, stats, ref, tr)


def test_cli_end_to_end(tmp_path):
    """`python -m cge_jl_b200` with the CGE_CLI.jl flags (example/CGE_CLI.jl:1-25): text files ->
    parseargs -> landmarks -> scorer -> printed 7-vector, in a fresh process."""
    import ast
    import os
    import subprocess
    import sys
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    (tmp_path / "g.edgelist").write_text("\n".join(f"{a - 1} {b - 1}" for a, b in edges))
    (tmp_path / "g.ecg").write_text("\n".join(str(c - 1) for c in comm[:, 0]) + "\n")
    (tmp_path / "g.emb").write_text(
        "\n".join(" ".join([str(i)] + [repr(float(x)) for x in emb[i]]) for i in range(115)) + "\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = [sys.executable, "-m", "cge_jl_b200", "-g", str(tmp_path / "g.edgelist"), "-c",
            str(tmp_path / "g.ecg"), "-e", str(tmp_path / "g.emb"), "--seed", "42",
            "--samples-local", "500"]
    for extra in ([], ["-l", "20", "-f", "1"], ["-d"]):
        r = subprocess.run(base + extra, cwd=root, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        vec = ast.literal_eval(r.stdout.strip().splitlines()[-1])
        assert len(vec) == 7 and 0.25 <= vec[0] <= 10.0 and vec[1] > 0
        assert set(r.stderr.strip()) <= {"."} or "Info" in r.stderr   # progress dots (divergence.jl:140)
    # exact undirected run equals the API call with the same seed
    direct = dv.wGCL(edges, ew, comm, emb, np.zeros(115), vw, *EMPTY, False, 42, 500, False)
    r = subprocess.run(base, cwd=root, capture_output=True, text=True, timeout=300)
    vec = ast.literal_eval(r.stdout.strip().splitlines()[-1])
    np.testing.assert_allclose(vec, direct, rtol=1e-12)
    bad = subprocess.run(base[:3] + ["-g", "/nonexistent"], cwd=root, capture_output=True, text=True)
    assert bad.returncode == 1 and "Usage" in bad.stdout
