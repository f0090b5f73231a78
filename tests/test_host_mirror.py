"""CPU: the host-side mirror of the reference's interface (parseargs, landmarks, sampling).
Modelled on /root/reference/test/runtests.jl:4-93 (three embedding formats, both .ecg formats,
the four split rules) but with value checks, which the reference's tests lack."""
import numpy as np
import pytest

from cge_jl_b200 import landmarks as landmarks_fn
from cge_jl_b200 import parseargs
from cge_jl_b200.divergence import draw_samples
from cge_jl_b200.landmarks import (split_cluster_diameter, split_cluster_rss, split_cluster_rss2,
                                   split_cluster_size)
from util import clusters_of, load_fixture, planted_partition


@pytest.fixture()
def files(tmp_path):
    """A 60-vertex graph written in the reference's text formats (0-based edgelist, 1-col ecg,
    2-col 1-based ecg, node2vec / ordered / unordered embeddings)."""
    edges, w, vw, comm, emb = planted_partition(60, 4, 6, seed=5, weighted=True)
    e0 = edges - 1
    (tmp_path / "g.edgelist").write_text("\n".join(f"{a} {b}" for a, b in e0))  # no trailing \n
    (tmp_path / "gw.edgelist").write_text("\n".join(f"{a}\t{b}\t{float(x)!r}" for (a, b), x in zip(e0, w)) + "\n")
    (tmp_path / "c1.ecg").write_text("\n".join(str(c - 1) for c in comm[:, 0]) + "\n")
    order = np.random.default_rng(1).permutation(60)
    (tmp_path / "c2.ecg").write_text("\n".join(f"{i + 1}\t{comm[i, 0]}" for i in order) + "\n")
    rows = [" ".join([str(i)] + [repr(float(x)) for x in emb[i]]) for i in range(60)]
    (tmp_path / "ordered.emb").write_text("\n".join(rows) + "\n")
    (tmp_path / "unordered.emb").write_text("\n".join(rows[i] for i in order) + "\n")
    (tmp_path / "n2v.emb").write_text("60 6\n" + "\n".join(rows[i] for i in order) + "\n")
    return tmp_path, edges, w, vw, comm, emb


def test_parseargs_formats(files):
    d, edges, w, vw, comm, emb = files
    a = parseargs(["-g", str(d / "g.edgelist"), "-c", str(d / "c1.ecg"), "-e", str(d / "n2v.emb"),
                   "-l", "20", "-f", "1", "-m", "rss"])
    b = parseargs(["-g", str(d / "g.edgelist"), "-c", str(d / "c2.ecg"), "-e", str(d / "ordered.emb")])
    c = parseargs(["-g", str(d / "gw.edgelist"), "-c", str(d / "c2.ecg"), "-e",
                   str(d / "unordered.emb"), "-d", "--split-global", "--seed", "7",
                   "--samples-local", "123", "-v", "-m", "Diameter "])
    for t in (a, b, c):
        assert t[0].dtype == np.int64 and t[0].min() == 1            # runtests.jl:23-24
        assert np.array_equal(t[0], edges)
        assert t[3].shape == (60, 1) and t[3].min() == 1             # :29-31
        assert np.array_equal(t[3], comm)
        assert np.array_equal(t[5], emb)
    assert np.all(a[1] == 1.0) and np.allclose(c[1], w)
    assert np.allclose(c[2], vw)
    assert a[7] == 20 and a[8] == 1 and a[9] is split_cluster_rss and len(a[4]) == 4
    assert b[7] == -1 and b[8] == 4 and b[4] == [] and b[12] == -1 and b[13] == 10000
    assert c[6] is True and c[10] is True and c[11] is True and c[12] == 7 and c[13] == 123
    assert c[9] is split_cluster_diameter


def test_parseargs_landmark_defaults(files):
    d = files[0]
    base = ["-g", str(d / "g.edgelist"), "-c", str(d / "c1.ecg"), "-e", str(d / "ordered.emb")]
    assert parseargs(base + ["-l"])[7] == round(4 * np.sqrt(60))      # auxilary.jl:179-183
    assert parseargs(base + ["-l", "--seed", "3"])[7] == round(4 * np.sqrt(60))
    t = parseargs(base + ["-f", "5"])                                  # auxilary.jl:186-189
    assert t[7] == 1 and t[8] == 5


def test_parseargs_errors_exit_1(files, capsys):
    d = files[0]
    with pytest.raises(SystemExit) as e:
        parseargs(["-e", str(d / "ordered.emb")])
    assert e.value.code == 1
    assert "Usage" in capsys.readouterr().out                          # auxilary.jl:221-246
    with pytest.raises(SystemExit):
        parseargs(["-g", str(d / "missing"), "-e", str(d / "ordered.emb")])
    with pytest.raises(SystemExit):   # no -c and no Louvain binaries in this build
        parseargs(["-g", str(d / "g.edgelist"), "-e", str(d / "ordered.emb")])


@pytest.mark.parametrize("rule", [split_cluster_rss, split_cluster_rss2, split_cluster_size,
                                  split_cluster_diameter])
def test_landmarks_rules(rule):
    """runtests.jl:43-93 for each split rule, plus the invariants of landmarks.jl:387-463."""
    edges, ew, vw, comm, emb = load_fixture("test115.npz")
    dii, lemb, lcomm, ledges, lw, lweight, v2l = landmarks_fn(
        edges, ew, vw, clusters_of(comm), comm, emb, False, 20, 1, rule, False)
    N = lemb.shape[0]
    assert N >= 20 and ledges.dtype == np.int64 and ledges.min() == 1 and ledges.max() <= N
    assert lcomm.shape == (N, 1) and dii.shape == (N,) and v2l.shape == (115,)
    assert np.all(ledges[:, 0] <= ledges[:, 1]) and np.all(lw > 0)
    assert np.isclose(lw.sum(), ew.sum()) and np.isclose(lweight.sum(), vw.sum())
    for L in range(1, N + 1):  # landmark = weighted centroid of one community's vertices
        mem = np.flatnonzero(v2l == L)
        assert mem.size >= 1 and len(set(comm[mem, 0])) == 1 and lcomm[L - 1, 0] == comm[mem[0], 0]
        cen = (emb[mem] * vw[mem, None]).sum(0) / vw[mem].sum()
        assert np.allclose(lemb[L - 1], cen)
        assert np.isclose(dii[L - 1], np.sqrt(((emb[mem] - cen) ** 2).sum() / vw[mem].sum()))


def test_landmarks_directed_keeps_orientation():
    edges, ew, vw, comm, emb = load_fixture("test115_weighted.npz")
    out = landmarks_fn(edges, ew, vw, clusters_of(comm), comm, emb, False, 20, 1,
                       split_cluster_rss, True)
    ledges, lw = out[3], out[4]
    assert np.isclose(lw.sum(), ew.sum()) and (ledges[:, 0] > ledges[:, 1]).any()


def test_draw_samples_semantics():
    edges, ew, vw, comm, emb = planted_partition(80, 4, 4, seed=2, weighted=True)
    n, K = 80, 400
    und = {(min(a, b), max(a, b)) for a, b in edges}
    pi, pj, pw, ni, nj = draw_samples(edges, ew, n, K, 42, False, True)
    assert pi.shape == (1, K)
    wmap = {(min(a, b), max(a, b)): x for (a, b), x in zip(edges, ew)}
    for a, b, x in zip(pi[0], pj[0], pw[0]):
        assert a <= b and wmap[(a, b)] == x                            # divergence.jl:133,203-206
    for a, b in zip(ni[0], nj[0]):
        assert a < b and (a, b) not in und                             # divergence.jl:121-137
    again = draw_samples(edges, ew, n, K, 42, False, True)
    assert all(np.array_equal(x, y) for x, y in zip((pi, pj, pw, ni, nj), again))
    # directed: ordered non-edges; exact mode takes pairs from a second draw (divergence.jl:510)
    dset = {(a, b) for a, b in edges}
    dpi, dpj, dpw, dni, dnj = draw_samples(edges, ew, n, K, 42, True, True)
    assert all((a, b) in dset for a, b in zip(dpi[0], dpj[0]))
    assert all(a != b and (a, b) not in dset for a, b in zip(dni[0], dnj[0]))
    assert np.array_equal(dpw, pw)
    lpi, lpj, lpw, _, _ = draw_samples(edges, ew, n, K, 42, True, False)
    assert not np.array_equal(lpi, dpi) and np.array_equal(lpw, pw)
    dmap = {(a, b): x for (a, b), x in zip(edges, ew)}
    assert all(dmap[(a, b)] == x for a, b, x in zip(lpi[0], lpj[0], lpw[0]))
    # unseeded: one set per alpha
    assert draw_samples(edges, ew, n, 10, -1, False, True)[0].shape == (40, 10)


def test_draw_samples_dense_and_complete_graphs():
    # complete graph: NE is empty, the reference's sample(NE, ...) throws on an empty collection
    tri = np.array([[1, 2], [2, 3], [1, 3]])
    with pytest.raises(ValueError, match="non-empty"):
        draw_samples(tri, np.ones(3), 3, 5, 42, False, True)
    # nearly complete graph: the single non-edge is always the one drawn
    k5 = np.array([[i, j] for i in range(1, 6) for j in range(i + 1, 6) if (i, j) != (2, 4)])
    _, _, _, ni, nj = draw_samples(k5, np.ones(len(k5)), 5, 20, 42, False, True)
    assert np.all(ni == 2) and np.all(nj == 4)


def _bf16_rn(x32):
    """Round-to-nearest-even float32 -> bfloat16 (kept in float32), as __float2bfloat16_rn."""
    b = x32.astype(np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000
    return b.astype(np.uint32).view(np.float32)


def test_diameter_filter_error_bound():
    """The tensor-core diameter filter (cge_diameter.cu) keeps every tile whose approximate maximum
    is within 2E of the global one, E = 1e-3 * max_i |x_i|^2.  Emulate its arithmetic (BF16 hi/lo
    split, three-term Gram entry, FP32 norms) and check that the real error stays an order of
    magnitude below E for centred data of several shapes and scales."""
    rng = np.random.default_rng(0)
    for n, d, scale in ((400, 128, 1.0), (300, 32, 1e3), (300, 7, 1e-3), (200, 64, 37.0)):
        x = rng.normal(size=(n, d)) * scale
        x[:5] *= 3.0                       # a few far points decide the diameter
        x = x - x.mean(0)
        hi = _bf16_rn(x.astype(np.float32))
        lo = _bf16_rn((x - hi.astype(np.float64)).astype(np.float32))
        assert np.abs(x - hi - lo).max() <= 2.0 ** -16 * np.abs(x).max()
        h64, l64 = hi.astype(np.float64), lo.astype(np.float64)
        gram = (h64 @ h64.T + h64 @ l64.T + l64 @ h64.T).astype(np.float32)   # FP32 accumulator
        nf = (x * x).sum(1).astype(np.float32)
        approx = (nf[:, None] + nf[None, :] - 2.0 * gram).astype(np.float64)
        exact = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
        rmax2 = (x * x).sum(1).max()
        err = np.abs(approx - exact).max()
        assert err <= 1e-4 * rmax2, (n, d, scale, err / rmax2)
        # hence the candidate rule cannot lose the true argmax
        E = 1e-3 * rmax2
        i, j = np.unravel_index(np.argmax(exact), exact.shape)
        assert approx[i, j] >= approx.max() - 2 * E


def test_builtin_principal_axis_matches_lapack():
    """cge_b200_sym_top_eigvec (the host-only eigen-solver behind cge_b200_landmarks_select when no
    LAPACK callback is given) against numpy.linalg.eigh on weighted covariance
    matrices: same principal axis to rounding, sign fixed to 'largest component positive'."""
    import ctypes as C

    from cge_jl_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(0)
    for d in (1, 2, 3, 5, 16, 33, 64, 128):
        for trial in range(4):
            s = int(rng.integers(d + 2, 4 * d + 10))
            y = rng.normal(size=(s, d)) * rng.uniform(0.1, 3.0, size=d)
            if trial == 2:
                y[:, 0] *= 1e3
            if trial == 3 and d > 2:  # nearly rank one
                y = y[:, :1] * rng.normal(size=d) + 1e-6 * y
            a = y.T @ y
            v, lam = np.zeros(d), C.c_double()
            assert lib.cge_b200_sym_top_eigvec(a.ctypes.data_as(_lib._pd), d, v.ctypes.data_as(_lib._pd),
                                               C.byref(lam)) == 0
            w, vec = np.linalg.eigh(a)
            ref = vec[:, -1] if vec[int(np.argmax(np.abs(vec[:, -1]))), -1] > 0 else -vec[:, -1]
            gap = (w[-1] - w[-2]) / w[-1] if d > 1 else 1.0
            assert abs(lam.value - w[-1]) <= 1e-13 * abs(w[-1])
            assert np.abs(v - ref).max() * gap <= 1e-13
            assert abs(np.linalg.norm(v) - 1.0) <= 1e-14 and v[int(np.argmax(np.abs(v)))] > 0
