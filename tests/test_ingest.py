"""the native table reader (cge_b200_table_dims / cge_b200_read_table) against
NumPy's parser on the three file kinds the reference reads with readdlm (auxilary.jl:86 edgelist,
:123 communities, :150-155 embedding), and the failure cases the reference relies on.  Host-only:
runs without a GPU."""
import numpy as np
import pytest

from cge_jl_b200.auxilary import parseargs, readdlm
from util import load_fixture


def _write(path, text):
    path.write_text(text)
    return str(path)


def test_reference_style_files_round_trip(tmp_path):
    edges, ew, vw, comm, emb = load_fixture("test115_weighted.npz")
    n = emb.shape[0]
    # edgelist: tab separated, weights in column 3 (example/10k.edgelist style)
    fe = tmp_path / "g.edgelist"
    fe.write_text("".join(f"{a}\t{b}\t{w!r}\n" for (a, b), w in zip(edges.tolist(), ew.tolist())))
    got = readdlm(str(fe))
    assert got.shape == (edges.shape[0], 3)
    assert np.array_equal(got[:, :2], edges) and np.array_equal(got[:, 2], ew)
    assert np.array_equal(got, np.loadtxt(fe, ndmin=2))
    # communities: one integer per line
    fc = tmp_path / "g.ecg"
    fc.write_text("".join(f"{c}\n" for c in comm[:, 0].tolist()))
    assert np.array_equal(readdlm(str(fc), np.int64), comm)
    # embedding: 0-based id in column 1, shortest round-trip decimals, shuffled rows
    order = np.random.default_rng(0).permutation(n)
    fb = tmp_path / "g.embedding"
    fb.write_text("".join(f"{i} " + " ".join(repr(float(x)) for x in emb[i]) + "\n" for i in order))
    got = readdlm(str(fb))
    assert np.array_equal(got[:, 0], order) and np.array_equal(got[:, 1:], emb[order])  # bit exact
    col_major = readdlm(str(fb), order="F")                     # Julia's Matrix{Float64} layout
    assert col_major.flags["F_CONTIGUOUS"] and np.array_equal(col_major, got)
    # the CLI mirror reads all three through the native reader
    out = parseargs(["-g", str(fe), "-c", str(fc), "-e", str(fb), "--seed", "42"])
    assert np.array_equal(out[0], edges) and np.array_equal(out[1], ew)
    assert np.array_equal(out[3], comm) and np.array_equal(out[5], emb)


def test_layout_details(tmp_path):
    text = "\n  1 2\t3  \r\n\n+4.5 -5e-1 .25\n  \t \n7 inf 9"      # blank lines, CRLF, no final newline
    got = readdlm(_write(tmp_path / "a.txt", text))
    assert got.tolist() == [[1.0, 2.0, 3.0], [4.5, -0.5, 0.25], [7.0, np.inf, 9.0]]
    assert readdlm(_write(tmp_path / "b.txt", text), skiprows=3).tolist() == [[4.5, -0.5, 0.25],
                                                                           [7.0, np.inf, 9.0]]
    assert readdlm(_write(tmp_path / "empty.txt", "")).shape == (0, 0)
    assert readdlm(_write(tmp_path / "blank.txt", "\n \n")).shape == (0, 0)
    one = readdlm(_write(tmp_path / "one.txt", "42"))
    assert one.shape == (1, 1) and one[0, 0] == 42.0
    # correctly rounded conversions: the cases where a naive digit accumulation is 1 ulp off
    hard = ["0.1", "2.2250738585072011e-308", "8.41e21", "9007199254740993", "1.7976931348623157e308",
            "4.9e-324", "0.30000000000000004", "123456789012345678901234567890"]
    got = readdlm(_write(tmp_path / "hard.txt", "\n".join(hard)))
    assert got[:, 0].tolist() == [float(s) for s in hard]


def test_failures_the_reference_relies_on(tmp_path):
    # node2vec format: a "rows dims" header line, then rows of 1 + dims cells (auxilary.jl:150-155)
    body = "".join(f"{i} {i + 0.5} {i + 0.25}\n" for i in range(5))
    fn = _write(tmp_path / "n2v.embedding", "5 2\n" + body)
    with pytest.raises(ValueError, match="does not have 2 columns"):
        readdlm(fn)
    assert readdlm(fn, skiprows=1).shape == (5, 3)
    with pytest.raises(ValueError, match="not a number"):
        readdlm(_write(tmp_path / "txt.txt", "1 2\n3 abc\n"))
    with pytest.raises(ValueError, match="not a number"):
        readdlm(_write(tmp_path / "partial.txt", "1 2\n3 4x\n"))
    with pytest.raises(ValueError, match="row 3 does not have 2 columns"):
        readdlm(_write(tmp_path / "short.txt", "1 2\n3 4\n5\n"))
    with pytest.raises(ValueError, match="is not a file"):
        readdlm(str(tmp_path / "missing.txt"))
    with pytest.raises(ValueError, match="InexactError"):
        readdlm(_write(tmp_path / "frac.ecg", "1\n2.5\n"), np.int64)


def test_threads_agree_on_a_larger_table(tmp_path):
    rng = np.random.default_rng(3)
    a = rng.normal(size=(20000, 16)) * 10.0 ** rng.integers(-8, 8, size=(20000, 16))
    fn = tmp_path / "big.txt"
    with open(fn, "w") as f:
        for i, row in enumerate(a):
            f.write(" ".join(repr(float(x)) for x in row) + ("\n\n" if i % 97 == 0 else "\n"))
    ref = readdlm(str(fn), n_threads=1)
    assert np.array_equal(ref, a)
    for th in (2, 3, 8, 0):
        assert np.array_equal(readdlm(str(fn), n_threads=th), ref)


REF_TEST = "/root/reference/test"


@pytest.mark.skipif(not __import__("os").path.isdir(REF_TEST),
                    reason="the reference checkout exists in the build container only")
def test_reference_own_fixture_files_through_the_native_reader():
    """The three command lines of the reference's own test (test/runtests.jl:4-19) on the reference's own
    files -- 0-based edgelist without a final newline, weighted edgelist, 1- and 2-column .ecg, node2vec /
    ordered / unordered embeddings -- parsed by the native reader behind `parseargs`: the assertions of
    runtests.jl:21-41, and the same arrays as the committed fixtures (tests/golden/make_fixtures.py)."""
    lines = [["-g", "test.edgelist", "-c", "test1col.ecg", "-e", "test_n2v.embedding"],
             ["-g", "test.edgelist", "-c", "test2col.ecg", "-e", "test_ordered.embedding"],
             ["-g", "test_weights.edgelist", "-c", "test2col.ecg", "-e", "test_unordered.embedding"]]
    fixtures = ["test115.npz", "test115.npz", "test115_weighted.npz"]
    for argv, fx in zip(lines, fixtures):
        argv = [a if a.startswith("-") else f"{REF_TEST}/{a}" for a in argv] + ["-l", "20", "-f", "1", "-m", "rss"]
        edges, weights, vweights, comm, clusters, embed, verbose, land, forced, method = parseargs(argv)[:10]
        assert edges.dtype == np.int64 and edges.ndim == 2 and edges.min() == 1          # runtests.jl:23-24
        assert weights.dtype == np.float64 and vweights.dtype == np.float64              # :26-27
        assert comm.dtype == np.int64 and comm.shape[1] == 1 and comm.min() == 1         # :29-31
        assert isinstance(clusters, list) and embed.dtype == np.float64 and embed.ndim == 2  # :33-34
        assert verbose is False and land == 20 and forced == 1                           # :36-39
        e, w, vw, c, emb = load_fixture(fx)
        assert np.array_equal(edges, e) and np.array_equal(weights, w) and np.array_equal(vweights, vw)
        assert np.array_equal(comm, c) and np.array_equal(embed, emb)
