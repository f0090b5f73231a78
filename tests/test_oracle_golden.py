"""CPU: pins the oracle (and the host-side landmarks mirror) to the reference's only published
known answer, README.md:99, and checks the small helper functions (idx, dist, JS)."""
import numpy as np
import pytest

import oracle
from cge_jl_b200.landmarks import landmarks, split_cluster_rss
from util import clusters_of, load_fixture

README_GOLDEN = [6.25, 0.002961243353776198, 0.0, 0.0, 9.75, 0.0017000000000000348,
                 0.000807441501038938]  # /root/reference/README.md:99


@pytest.fixture(scope="module")
def cfg1():
    edges, ew, vw, comm, emb = load_fixture("example10k.npz")
    lm = landmarks(edges, ew, vw, clusters_of(comm), comm, emb, False, 200, 4,
                   split_cluster_rss, False)
    return (edges, ew, vw, comm, emb), lm


def test_landmark_count(cfg1):
    # -l 200 with the default 4 forced splits of 64 communities gives 256 landmarks
    _, (dii, lemb, lcomm, ledges, lw, lweight, v2l) = cfg1
    assert lemb.shape == (256, 32) and dii.shape == (256,)
    assert ledges.min() == 1 and ledges.max() == 256
    assert v2l.min() == 1 and v2l.max() == 256 and v2l.shape == (10000,)
    assert np.isclose(lweight.sum(), 2 * 41536)


def test_readme_golden_global(cfg1):
    """Elements 1-2 of README.md:99 are RNG-free: the oracle must reproduce them."""
    (edges, ew, vw, comm, emb), (dii, lemb, lcomm, ledges, lw, lweight, v2l) = cfg1
    out, tr = oracle.wgcl(ledges, lw, lcomm, lemb, dii, lweight, vw, v2l, emb, False, None)
    assert out[0] == README_GOLDEN[0]
    assert abs(out[1] - README_GOLDEN[1]) / README_GOLDEN[1] < 1e-12
    assert out[2] == 0.0 and out[3] == 0.0
    assert list(tr.iters)[:15] == [49, 39, 39, 38, 38, 37, 37, 36, 35, 34, 33, 32, 32, 31, 30]
    assert tr.n_alpha_run == 30  # global early stop: best at 6.25, patience 5


def test_readme_golden_local_error_formula():
    # element 7 is a pure function of element 6 (divergence.jl:217)
    auc = README_GOLDEN[5]
    assert np.isclose(1.96 * np.sqrt(auc * (1 - auc) / 10000), README_GOLDEN[6], rtol=1e-12)


def test_idx_matches_definition():
    # auxilary.jl:57-59: packed row-major upper triangle with diagonal, 1-based
    n = 7
    k = 1
    for i in range(1, n + 1):
        for j in range(i, n + 1):
            assert oracle.idx(n, i, j) == k
            k += 1


def test_dist_and_js():
    rng = np.random.default_rng(0)
    e = rng.normal(size=(5, 9))
    assert oracle.dist(2, 2, e) == 0.0
    assert np.isclose(oracle.dist(1, 4, e), np.linalg.norm(e[0] - e[3]), rtol=1e-15)
    c, b = rng.uniform(0, 5, 10), rng.uniform(0, 5, 10)
    p, q = (c + 1) / (c.sum() + 10), (b + 1) / (b.sum() + 10)
    m = (p + q) / 2
    want = 0.5 * np.sum(p * np.log(p / m) + q * np.log(q / m))
    assert np.isclose(oracle.js(c, b), want, rtol=1e-13)
    assert oracle.js(c, c) == 0.0
    mask = np.arange(10) % 3 == 0
    ci, bi = c[mask], b[mask]
    p, q = (ci + 1) / (ci.sum() + ci.size), (bi + 1) / (bi.sum() + bi.size)
    m = (p + q) / 2
    want = 0.5 * np.sum(p * np.log(p / m) + q * np.log(q / m))
    assert np.isclose(oracle.js(c, b, mask, True), want, rtol=1e-13)


def test_oracle_small_graph_values():
    """Frozen oracle values for the reference's 115-node test graph (SURVEY section 6 probes)."""
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    out, tr = oracle.wgcl(edges, ew, comm, emb, np.zeros(115), vw)
    assert out[0] == 3.25
    assert np.isclose(out[1], 0.006929334486296551, rtol=1e-10)
    edges, ew, vw, comm, emb = load_fixture("test115_weighted.npz")
    out, tr = oracle.wgcl_directed(edges, ew, comm, emb, np.zeros(115), vw)
    assert out.shape == (7,) and out[0] == 5.5
    assert np.isclose(out[1], 0.008821457041054588, rtol=1e-10)


def test_oracle_star_graph():
    n = 6
    edges = np.array([[1, j] for j in range(2, n + 1)])
    out, _ = oracle.wgcl_directed(edges, np.ones(n - 1), np.ones((n, 1), dtype=np.int64),
                                  np.random.default_rng(1).normal(size=(n, 4)), np.zeros(n),
                                  np.ones(n))
    assert out.tolist() == [-1.0, 0, 0, 0, 0, 0]  # divergence.jl:332-334
