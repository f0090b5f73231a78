"""``wGCL`` / ``wGCL_directed`` with the reference's signatures, running on the B200.

Mirror of the Julia entry points (/root/reference/src/divergence.jl:27-31, 282-286): same
positional arguments (1-based ids, as ``parseargs`` / ``landmarks`` produce them), same 7-element
return vector (6 for the directed star-graph exit), same assertion messages, same progress dots
on stderr.  What stays on the host is what SURVEY.md section 8(b) leaves in the Julia wrapper:
building E / NE and drawing the sampled pairs for the local score (divergence.jl:121-137,
184-210, 484-513).  Everything else happens inside ``libcge_b200.so`` through the C ABI.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _lib

N_ALPHA = _lib.N_ALPHA
_ASSERTS = {
    _lib.ERR_ASSERT_COMM: "No. communities not matching no. vertices",
    _lib.ERR_ASSERT_DIST: "Distances vector length is not equal to no. vertices",
}


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64)) if a is not None else None


# ---------------------------------------------------------------------------------------------
# sampling (host side of the boundary)
# ---------------------------------------------------------------------------------------------
def _sample_non_edges(rng, n, codes, K, directed):
    """K pairs drawn uniformly with replacement from NE (divergence.jl:121-137 / 405-421):
    unordered i<j (ordered i!=j when directed) that are not in the edge set.  The reference
    materialises NE and indexes it; rejection sampling draws from the same distribution without
    the O(n^2) array.  ``codes`` = i*(n+1)+j of every edge (unsorted, duplicates allowed); only the
    few thousand candidates are sorted, the edge list is scanned once per round."""
    n_pairs = n * (n - 1) if directed else n * (n - 1) // 2
    if n_pairs - codes.shape[0] * 8 < 0 and n <= 5000:
        # possibly dense: count exactly and enumerate NE like the reference does
        uniq = np.unique(codes)
        n_ne = n_pairs - int(np.count_nonzero(uniq // (n + 1) != uniq % (n + 1)))
        if n_ne <= 0:
            raise ValueError("collection must be non-empty: the graph has no non-edges to sample "
                             "(divergence.jl:137 leaves NE empty)")
        if n_ne * 8 < n_pairs:
            ii, jj = np.meshgrid(np.arange(1, n + 1), np.arange(1, n + 1), indexing="ij")
            keep = (ii != jj) if directed else (ii < jj)
            ii, jj = ii[keep], jj[keep]
            free = ~np.isin(ii * (n + 1) + jj, uniq)
            pick = rng.integers(0, int(free.sum()), size=K)
            return ii[free][pick], jj[free][pick]
    elif n_pairs <= codes.shape[0] and n_pairs - np.unique(codes).shape[0] <= 0:
        raise ValueError("collection must be non-empty: the graph has no non-edges to sample "
                         "(divergence.jl:137 leaves NE empty)")
    out_i = np.empty(K, dtype=np.int64)
    out_j = np.empty(K, dtype=np.int64)
    got = 0
    while got < K:
        need = max(2 * (K - got), 64)
        i = rng.integers(1, n + 1, size=need)
        j = rng.integers(1, n + 1, size=need)
        if not directed:
            i, j = np.minimum(i, j), np.maximum(i, j)
        code = i * (n + 1) + j
        ok = i != j
        if codes.size:
            hit = codes[np.isin(codes, code)]       # the candidates that are edges
            ok &= ~np.isin(code, hit)
        i, j = i[ok][: K - got], j[ok][: K - got]
        out_i[got: got + i.shape[0]] = i
        out_j[got: got + i.shape[0]] = j
        got += i.shape[0]
    return out_i, out_j


# above this many vertices the Julia wrapper (julia/CGEB200.jl) stops materialising NE; the device
# sampler takes over when a Scorer is passed to draw_samples
NE_MATERIALIZE_LIMIT = 30_000


def draw_samples(adj_edges, adj_eweights, adj_n, K, seed, directed, exact, device=None):
    """Draw the positive / negative pairs of the local score where the reference draws them.

    Returns ``(pos_i, pos_j, pos_w, neg_i, neg_j)`` with shape ``(n_sets, K)``, 1-based ids of
    the original graph.  ``seed != -1`` reseeds before every draw, so every alpha sees the same
    sets (n_sets = 1, divergence.jl:184,193,202,209); ``seed == -1`` draws fresh sets per alpha
    (n_sets = 40).  In directed exact mode the positive pairs come from a second, unseeded
    continuation draw while the weights stay those of the first (the overwrite at
    divergence.jl:505-510).  NumPy's generator replaces Julia's RNG stream: the sets are
    identically distributed, not identical.

    ``device``: a :class:`Scorer`; the negative pairs then come from
    ``cge_b200_sample_non_edges`` (edge hash set + rejection draws on the GPU) instead of the NumPy
    rejection sampler -- what a host uses once NE no longer fits (``NE_MATERIALIZE_LIMIT``).
    """
    adj_edges = _i64(adj_edges)
    w = _f64(adj_eweights)
    m = adj_edges.shape[0]
    if directed:
        e_i, e_j = adj_edges[:, 0], adj_edges[:, 1]
    else:
        e_i, e_j = adj_edges.min(axis=1), adj_edges.max(axis=1)  # divergence.jl:133
    codes = e_i * (adj_n + 1) + e_j
    n_sets = 1 if seed != -1 else N_ALPHA
    pi = np.empty((n_sets, K), dtype=np.int64)
    pj = np.empty_like(pi)
    ni = np.empty_like(pi)
    nj = np.empty_like(pi)
    pw = np.empty((n_sets, K))
    free_rng = np.random.default_rng()
    if device is not None:
        dev_seed = int(seed) if seed != -1 else int(free_rng.integers(0, 2**63))
        ni, nj = device.sample_non_edges(adj_edges, adj_n, K, n_sets, dev_seed, directed)
    for s in range(n_sets):
        rng = np.random.default_rng(seed) if seed != -1 else free_rng
        idx = rng.integers(0, m, size=K)
        pw[s] = w[idx]
        if directed and exact:
            idx = rng.integers(0, m, size=K)  # second draw, no reseed (divergence.jl:510)
        pi[s], pj[s] = e_i[idx], e_j[idx]
        if device is None:
            rng = np.random.default_rng(seed) if seed != -1 else free_rng
            ni[s], nj[s] = _sample_non_edges(rng, adj_n, codes, K, directed)
    return pi, pj, pw, ni, nj


# ---------------------------------------------------------------------------------------------
# C ABI call
# ---------------------------------------------------------------------------------------------
class Scorer:
    """Thin RAII wrapper over a ``cge_b200_handle`` (device buffers persist between calls)."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.cge_b200_create(int(device), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"cge_b200_create failed ({rc}): {_lib.last_error()}")
        self._keep = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.cge_b200_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def comm_init(self, id_bytes, rank, n_ranks):
        buf = C.create_string_buffer(bytes(id_bytes), len(id_bytes))
        rc = self._lib.cge_b200_comm_init(self._h, C.cast(buf, C.c_void_p), rank, n_ranks)
        _check(rc)

    def p2p_export(self, max_vertices):
        buf = C.create_string_buffer(self._lib.cge_b200_p2p_handle_size())
        _check(self._lib.cge_b200_p2p_export(self._h, int(max_vertices), C.cast(buf, C.c_void_p)))
        return buf.raw

    def p2p_import(self, handles):
        blob = b"".join(handles)
        buf = C.create_string_buffer(blob, len(blob))
        _check(self._lib.cge_b200_p2p_import(self._h, C.cast(buf, C.c_void_p)))

    def upload(self, problem, keep):
        self._keep = keep
        _check(self._lib.cge_b200_upload(self._h, C.byref(problem)))

    def run(self):
        out = np.zeros(7)
        n_out = C.c_int32(7)
        stats = _lib.Stats()
        _check(self._lib.cge_b200_run(self._h, _pd(out), C.byref(n_out), C.byref(stats)))
        return out[: n_out.value].copy(), stats

    def fp64_peak_tflops(self):
        v = C.c_double()
        _check(self._lib.cge_b200_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def fp64_pipes_tflops(self):
        """``cge_b200_measure_fp64_pipes``: DFMA / DMMA throughput of the device, see the header."""
        out = np.zeros(8)
        _check(self._lib.cge_b200_measure_fp64_pipes(self._h, _pd(out)))
        return dict(zip(("dfma", "dmma_m8n8k4", "dmma_m16n8k16", "mix_m8n8k4_dmma", "mix_m8n8k4_dfma",
                         "mix_m16n8k16_dmma", "mix_m16n8k16_dfma", "dmma_m16n8k8"), out.tolist()))

    def sample_non_edges(self, edges, n, K, n_sets=1, seed=0, directed=False, index_base=1,
                         return_draws=False):
        """``cge_b200_sample_non_edges``: ``(neg_i, neg_j)`` of shape ``(n_sets, K)``, uniform over
        the non-edges of the ``n``-vertex graph with replacement."""
        edges = _i64(edges).reshape(-1, 2)
        src, dst = _i64(edges[:, 0]), _i64(edges[:, 1])
        oi = np.empty((int(n_sets), int(K)), dtype=np.int64)
        oj = np.empty_like(oi)
        draws = C.c_double()
        _check(self._lib.cge_b200_sample_non_edges(
            self._h, int(n), int(edges.shape[0]), _pi(src), _pi(dst), int(index_base),
            int(bool(directed)), int(K), int(n_sets), int(seed) & (2**64 - 1), _pi(oi), _pi(oj),
            C.byref(draws)))
        return (oi, oj, draws.value) if return_draws else (oi, oj)

    def landmarks_aggregate(self, landmark, vweights, comm, embedding, edges, eweights, directed,
                            n_landmarks=None):
        """``cge_b200_landmarks_aggregate``: the aggregation half of
        ``landmarks()`` (landmarks.jl:387-463) on the device, from the 1-based vertex -> landmark
        assignment of ``runsplit``.  Returns ``(dii, embed, cluster, landmark_edges, weights,
        lweight)`` in the reference's conventions (1-based ids, ``cluster`` as a column)."""
        lm = _i64(landmark)
        vw, em, ew = _f64(vweights), np.asarray(embedding, dtype=np.float64), _f64(eweights)
        cm = _i64(np.asarray(comm).reshape(-1))
        edges = _i64(edges).reshape(-1, 2)
        src, dst = _i64(edges[:, 0]), _i64(edges[:, 1])
        n, d = em.shape
        N = int(n_landmarks if n_landmarks is not None else lm.max())
        m = edges.shape[0]
        cap = int(min(m, N * N)) + 1
        embed, lweight, dii = np.zeros((N, d)), np.zeros(N), np.zeros(N)
        cluster = np.zeros(N, dtype=np.int64)
        oa, ob, ow = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64), np.zeros(cap)
        n_e = C.c_int64()
        _check(self._lib.cge_b200_landmarks_aggregate(
            self._h, n, d, N, _pi(lm), 1, _pd(vw), _pi(cm), _pd(em), em.strides[0] // 8,
            em.strides[1] // 8, m, _pi(src), _pi(dst), _pd(ew), int(bool(directed)), _pd(embed),
            _pd(lweight), _pd(dii), _pi(cluster), _pi(oa), _pi(ob), _pd(ow), cap, C.byref(n_e)))
        k = n_e.value
        return (dii, embed, cluster.reshape(-1, 1), np.stack([oa[:k], ob[:k]], axis=1), ow[:k].copy(),
                lweight)

    def unique_rows(self, embedding):
        """``cge_b200_unique_rows``: ``size(unique(embedding, dims=1), 1)`` (landmarks.jl:369) on the
        device."""
        em = np.asarray(embedding, dtype=np.float64)
        n, d = em.shape
        out = C.c_int64()
        _check(self._lib.cge_b200_unique_rows(self._h, n, d, _pd(em), em.strides[0] // 8, em.strides[1] // 8,
                                              C.byref(out)))
        return out.value

    def landmarks_select(self, embedding, vweights, clusters, land, forced, rule, eig="lapack"):
        """``cge_b200_landmarks_select``: ``runsplit`` (landmarks.jl:279-345) with
        the cuts of the split rule on the device.  ``clusters``: list of 1-based vertex-id arrays
        (any order: they are sorted like the reference's ``sort(initial_clusters)``); ``rule``:
        ``"rss"``, ``"size"`` or ``"diameter"``.  ``eig``: ``"lapack"`` hands the d x d eigenproblem of
        every cut to ``numpy.linalg.eigh`` through the library's callback (the routine the host mirror
        calls, sign included); ``"builtin"`` uses the library's own solver (sign fixed: largest-
        magnitude component positive).  Returns ``(group, cuts)``: 0-based landmark id per vertex, and
        the number of cluster cuts made."""
        code = {"rss": 0, "size": 2, "diameter": 3}[rule]
        em, vw = np.asarray(embedding, dtype=np.float64), _f64(vweights)
        n, d = em.shape
        cl = sorted((np.asarray(c, dtype=np.int64) for c in clusters), key=lambda c: c.tolist())
        ptr = np.zeros(len(cl) + 1, dtype=np.int64)
        ptr[1:] = np.cumsum([c.size for c in cl])
        members = _i64(np.concatenate(cl)) if cl else np.zeros(0, dtype=np.int64)
        group = np.zeros(n, dtype=np.int64)
        cuts = C.c_int64()

        def _eigh(c, dd, v, _user):
            try:
                a = np.ctypeslib.as_array(c, shape=(dd, dd))
                np.ctypeslib.as_array(v, shape=(dd,))[:] = np.linalg.eigh(a)[1][:, -1]
                return 0
            except Exception:  # noqa: BLE001 -- must not unwind through the C frames
                return 1

        cb = _lib.EIGVEC_FN(_eigh) if eig == "lapack" else C.cast(None, _lib.EIGVEC_FN)
        _check(self._lib.cge_b200_landmarks_select(
            self._h, n, d, _pd(em), em.strides[0] // 8, em.strides[1] // 8, _pd(vw), len(cl), _pi(ptr),
            _pi(members), 1, int(land), int(forced), code, cb, None, _pi(group), C.byref(cuts)))
        return group, cuts.value

    def selftest_math(self, n_samples, seed=1):
        """(square roots, normalisations) of the recompute epilogue's short branch-free forms that
        land more than 2 ulp from the correctly rounded operations on ``n_samples`` pseudo-random
        operands (``cge_b200_selftest_math``)."""
        a, b = C.c_int64(), C.c_int64()
        _check(self._lib.cge_b200_selftest_math(self._h, int(n_samples), int(seed), C.byref(a),
                                                C.byref(b)))
        return a.value, b.value

    def debug_read(self, what, n):
        size = n * n if what == 0 else n
        buf = np.zeros(size)
        _check(self._lib.cge_b200_debug_read(self._h, what, _pd(buf), size))
        return buf.reshape(n, n) if what == 0 else buf


def score_multi(problem, n_gpus):
    """``cge_b200_score_multi``: one call, ``n_gpus`` GPUs of this box driven by threads of this
    process (what a single-process host such as Julia uses; ``CGE_B200_GPUS`` does the same for
    plain ``cge_b200_score``).  Returns ``(out, stats)``."""
    lib = _lib.load()
    out = np.zeros(7)
    n_out = C.c_int32(7)
    stats = _lib.Stats()
    _check(lib.cge_b200_score_multi(C.byref(problem), int(n_gpus), _pd(out), C.byref(n_out),
                                    C.byref(stats)))
    return out[: n_out.value].copy(), stats


def unique_id():
    lib = _lib.load()
    buf = C.create_string_buffer(lib.cge_b200_comm_id_size())
    _check(lib.cge_b200_comm_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw


def _check(rc):
    if rc == 0:
        return
    if rc in _ASSERTS:
        raise AssertionError(_ASSERTS[rc])
    if rc == _lib.ERR_OOM:
        raise MemoryError(_lib.last_error())
    raise RuntimeError(f"libcge_b200 error {rc}: {_lib.last_error()}")


def make_problem(edges, eweights, comm, embed, distances, vweights, init_vweights, v_to_l,
                 init_embed, split, directed, samples, max_alphas=0, driver=0, regime=0):
    """Fill a ``cge_b200_problem`` from reference-style (1-based) arrays.

    Returns ``(problem, keep)``; ``keep`` holds the arrays the struct points into.
    """
    edges = _i64(edges)
    src, dst = _i64(edges[:, 0]), _i64(edges[:, 1])
    ew = _f64(eweights)
    cm = _i64(np.asarray(comm).reshape(-1))
    em = np.asarray(embed, dtype=np.float64)
    if em.ndim != 2:
        raise ValueError("embed must be a matrix")
    di, vw = _f64(distances), _f64(vweights)
    p = _lib.Problem()
    p.struct_size = C.sizeof(_lib.Problem)
    p.index_base = 1
    p.directed, p.split = int(bool(directed)), int(bool(split))
    p.m, p.edge_src, p.edge_dst, p.eweights = src.shape[0], _pi(src), _pi(dst), _pd(ew)
    p.n_comm, p.comm = cm.shape[0], _pi(cm)
    p.embed, p.embed_rows, p.d = _pd(em), em.shape[0], em.shape[1]
    p.embed_row_stride, p.embed_col_stride = em.strides[0] // 8, em.strides[1] // 8
    p.n_distances, p.distances, p.vweights = di.shape[0], _pd(di), _pd(vw)
    keep = [src, dst, ew, cm, em, di, vw]
    n_full = 0 if v_to_l is None else len(v_to_l)
    if n_full:
        ivw, v2l = _f64(init_vweights), _i64(v_to_l)
        iem = np.asarray(init_embed, dtype=np.float64)
        p.n_full, p.init_vweights, p.v_to_l, p.init_embed = n_full, _pd(ivw), _pi(v2l), _pd(iem)
        p.init_row_stride, p.init_col_stride = iem.strides[0] // 8, iem.strides[1] // 8
        keep += [ivw, v2l, iem]
    if samples is not None:
        pi, pj, pw, ni, nj = samples
        pi, pj, ni, nj = (np.atleast_2d(_i64(x)) for x in (pi, pj, ni, nj))
        pw = np.atleast_2d(_f64(pw))
        p.n_sets, p.n_samples = pi.shape
        p.pos_i, p.pos_j, p.pos_w, p.neg_i, p.neg_j = _pi(pi), _pi(pj), _pd(pw), _pi(ni), _pi(nj)
        keep += [pi, pj, pw, ni, nj]
    p.max_alphas, p.driver, p.regime = int(max_alphas), int(driver), int(regime)
    return p, keep


def _score(directed, edges, eweights, comm, embed, distances, vweights, init_vweights, v_to_l,
           init_edges, init_eweights, init_embed, split, seed, auc_samples, verbose, samples,
           return_stats, max_alphas, driver, scorer, regime=0):
    edges = _i64(edges)
    no_vertices = int(edges.max())                      # divergence.jl:41
    no_edges = edges.shape[0]
    if verbose:
        print(f"auc_samples: {auc_samples}")
    landmarks = v_to_l is not None and len(v_to_l) > 0  # divergence.jl:44
    if verbose:
        print(f"Graph has {no_vertices} vertices and {no_edges} edges")
        if landmarks:
            ie = _i64(init_edges)
            print(f"Original graph has {int(ie.max())} vertices and {ie.shape[0]} edges")
    comm = np.asarray(comm)
    if comm.reshape(-1).shape[0] != no_vertices:        # divergence.jl:50 (before anything else)
        raise AssertionError(_ASSERTS[_lib.ERR_ASSERT_COMM])
    if verbose:
        print(f"Graph has {int(comm.max())} communities")
        print(f"Embedding has {np.asarray(embed).shape[1]} dimensions")
    if len(distances) != no_vertices:                   # divergence.jl:81
        raise AssertionError(_ASSERTS[_lib.ERR_ASSERT_DIST])
    own = scorer is None
    sc = Scorer() if own else scorer
    try:
        if samples is None and auc_samples > 0:
            adj_edges = init_edges if landmarks else edges  # divergence.jl:95-102
            adj_w = init_eweights if landmarks else eweights
            adj_n = len(init_vweights) if landmarks else no_vertices
            # NE (n^2/2 tuples, divergence.jl:121-137) stops fitting: negatives drawn on the device
            samples = draw_samples(adj_edges, adj_w, adj_n, int(auc_samples), int(seed), directed,
                                   exact=not landmarks,
                                   device=sc if adj_n > NE_MATERIALIZE_LIMIT else None)
        problem, keep = make_problem(edges, eweights, comm, embed, distances, vweights,
                                     init_vweights if landmarks else None,
                                     v_to_l if landmarks else None,
                                     init_embed if landmarks else None, split, directed, samples,
                                     max_alphas, driver, regime)
        sc.upload(problem, keep)
        out, stats = sc.run()
    finally:
        if own:
            sc.close()
    sys.stderr.write("." * int(stats.n_alpha_run) + "\n")  # divergence.jl:140,255
    if verbose and directed and out.shape[0] == 6:
        print("Graph is a star in respect to either in or out edges")
    return (out, stats) if return_stats else out


def wGCL(edges, eweights, comm, embed, distances, vweights, init_vweights, v_to_l, init_edges,
         init_eweights, init_embed, split, seed=-1, auc_samples=10000, verbose=False, *,
         samples=None, return_stats=False, max_alphas=0, driver=0, scorer=None, regime=0):
    """Weighted Geometric Chung-Lu fit + global/local divergence (divergence.jl:27-257).

    Positional arguments and the returned ``[best_alpha, best_div, best_div_ext, best_div_int,
    best_alpha_auc, best_auc, best_auc_err]`` are the reference's.  Keyword-only extras:
    ``samples`` (pre-drawn pairs, see :func:`draw_samples`), ``return_stats``, ``max_alphas``,
    ``driver`` and ``scorer`` (a reusable :class:`Scorer`).
    """
    return _score(False, edges, eweights, comm, embed, distances, vweights, init_vweights,
                  v_to_l, init_edges, init_eweights, init_embed, split, seed, auc_samples,
                  verbose, samples, return_stats, max_alphas, driver, scorer, regime)


def wGCL_directed(edges, eweights, comm, embed, distances, vweights, init_vweights, v_to_l,
                  init_edges, init_eweights, init_embed, split, seed=-1, auc_samples=10000,
                  verbose=False, *, samples=None, return_stats=False, max_alphas=0, driver=0,
                  scorer=None, regime=0):
    """Directed variant (divergence.jl:282-561); 6-element ``[-1,0,0,0,0,0]`` for a star graph."""
    return _score(True, edges, eweights, comm, embed, distances, vweights, init_vweights,
                  v_to_l, init_edges, init_eweights, init_embed, split, seed, auc_samples,
                  verbose, samples, return_stats, max_alphas, driver, scorer, regime)
