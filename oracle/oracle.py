"""ctypes loader for oracle/libcge_oracle.so (TEST INFRASTRUCTURE ONLY).

The argument lists mirror the reference signatures ``wGCL`` / ``wGCL_directed``
(/root/reference/src/divergence.jl:27-31, 282-286) except that the sampled edge and non-edge
index arrays are explicit inputs (``samples``) instead of being drawn inside.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcge_oracle.so")
N_ALPHA = 40


class OracleTrace(C.Structure):
    _fields_ = [
        ("n_alpha_run", C.c_int32),
        ("iters", C.c_int32 * N_ALPHA),
        ("div", C.c_double * N_ALPHA),
        ("auc", C.c_double * N_ALPHA),
        ("lo", C.c_double),
        ("hi", C.c_double),
        ("hi_full", C.c_double),
        ("final_diff", C.c_double),
        ("t_phase", C.c_double * 5),  # wGCL: seconds in D build, GD, passes, P, B
    ]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("cge_oracle.c", "cge_oracle_mt.c", "cge_oracle_stream.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libcge_oracle.so"])
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.cge_oracle_idx.restype = C.c_int64
        _lib.cge_oracle_idx.argtypes = [C.c_int64] * 3
        _lib.cge_oracle_dist.restype = C.c_double
        _lib.cge_oracle_js.restype = C.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def idx(n, i, j):
    return int(_load().cge_oracle_idx(n, i, j))


def dist(v1, v2, embed):
    e = _f64(embed)
    lib = _load()
    lib.cge_oracle_dist.argtypes = [C.c_int64, C.c_int64, C.POINTER(C.c_double), C.c_int64]
    return float(lib.cge_oracle_dist(v1, v2, _p(e, C.c_double), e.shape[1]))


def js(vC, vB, vI=None, internal=True):
    vC, vB = _f64(vC), _f64(vB)
    lib = _load()
    lib.cge_oracle_js.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(C.c_uint8), C.c_int, C.c_int64]
    m = None if vI is None or len(vI) == 0 else np.ascontiguousarray(vI, dtype=np.uint8)
    return float(lib.cge_oracle_js(_p(vC, C.c_double), _p(vB, C.c_double), _p(m, C.c_uint8),
                                   int(bool(internal)), vC.shape[0]))


def _common(edges, eweights, comm, embed, distances, vweights, init_vweights, v_to_l,
            init_embed, samples):
    edges = _i64(edges)
    src, dst = _i64(edges[:, 0]), _i64(edges[:, 1])
    ew, cm = _f64(eweights), _i64(np.asarray(comm).reshape(-1))
    em, di, vw = _f64(embed), _f64(distances), _f64(vweights)
    n_full = len(v_to_l) if v_to_l is not None else 0
    if n_full:
        ivw, v2l, iem = _f64(init_vweights), _i64(v_to_l), _f64(init_embed)
    else:
        ivw = v2l = iem = None
    if samples is None:
        K, ns = 0, 1
        pi = pj = ni = nj = np.zeros(1, dtype=np.int64)
        pw = np.zeros(1)
    else:
        pi, pj, pw, ni, nj = samples
        pi, pj, ni, nj = (np.atleast_2d(_i64(x)) for x in (pi, pj, ni, nj))
        pw = np.atleast_2d(_f64(pw))
        ns, K = pi.shape
    keep = (src, dst, ew, cm, em, di, vw, ivw, v2l, iem, pi, pj, pw, ni, nj)
    args = [src.shape[0], _p(src, C.c_int64), _p(dst, C.c_int64), _p(ew, C.c_double),
            _p(cm, C.c_int64), cm.shape[0], _p(em, C.c_double), em.shape[1],
            _p(di, C.c_double), di.shape[0], _p(vw, C.c_double), n_full,
            _p(ivw, C.c_double), _p(v2l, C.c_int64), _p(iem, C.c_double)]
    sargs = [K, ns, _p(pi, C.c_int64), _p(pj, C.c_int64), _p(pw, C.c_double),
             _p(ni, C.c_int64), _p(nj, C.c_int64)]
    return keep, args, sargs


_ARGT = ([C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double),
          C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_double), C.c_int64,
          C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.c_int64,
          C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_int,
          C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
          C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int,
          C.POINTER(C.c_double)])


def wgcl(edges, eweights, comm, embed, distances, vweights, init_vweights=None, v_to_l=None,
         init_embed=None, split=False, samples=None, max_alphas=N_ALPHA):
    """Oracle for ``wGCL`` (divergence.jl:27-257).  Returns (out[7], OracleTrace).

    ``samples`` = (pos_i, pos_j, pos_w, neg_i, neg_j), each (n_sets, K), 1-based ids of the
    original graph; ``None`` computes the global score only.
    """
    lib = _load()
    keep, args, sargs = _common(edges, eweights, comm, embed, distances, vweights,
                                init_vweights, v_to_l, init_embed, samples)
    lib.cge_oracle_wgcl.argtypes = _ARGT + [C.POINTER(OracleTrace)]
    lib.cge_oracle_wgcl.restype = C.c_int
    out = np.zeros(7)
    tr = OracleTrace()
    rc = lib.cge_oracle_wgcl(*args, int(bool(split)), *sargs, int(max_alphas),
                             _p(out, C.c_double), C.byref(tr))
    if rc != 0:
        raise AssertionError({-2: "No. communities not matching no. vertices",
                              -3: "Distances vector length is not equal to no. vertices"}
                             .get(rc, f"oracle error {rc}"))
    del keep
    return out, tr


def wgcl_directed(edges, eweights, comm, embed, distances, vweights, init_vweights=None,
                  v_to_l=None, init_embed=None, split=False, samples=None, max_alphas=N_ALPHA):
    """Oracle for ``wGCL_directed`` (divergence.jl:282-561).  Returns (out[6 or 7], trace)."""
    lib = _load()
    keep, args, sargs = _common(edges, eweights, comm, embed, distances, vweights,
                                init_vweights, v_to_l, init_embed, samples)
    lib.cge_oracle_wgcl_directed.argtypes = _ARGT + [C.POINTER(C.c_int), C.POINTER(OracleTrace)]
    lib.cge_oracle_wgcl_directed.restype = C.c_int
    out = np.zeros(7)
    n_out = C.c_int(7)
    tr = OracleTrace()
    rc = lib.cge_oracle_wgcl_directed(*args, int(bool(split)), *sargs, int(max_alphas),
                                      _p(out, C.c_double), C.byref(n_out), C.byref(tr))
    if rc != 0:
        raise AssertionError({-2: "No. communities not matching no. vertices",
                              -3: "Distances vector length is not equal to no. vertices"}
                             .get(rc, f"oracle error {rc}"))
    del keep
    return out[: n_out.value].copy(), tr


class OracleMtTrace(C.Structure):
    _fields_ = [
        ("n_alpha_run", C.c_int32),
        ("iters", C.c_int32 * N_ALPHA),
        ("div", C.c_double * N_ALPHA),
        ("auc", C.c_double * N_ALPHA),
        ("lo", C.c_double),
        ("hi", C.c_double),
        ("threads", C.c_int32),
    ]


def host_threads():
    return int(_load().cge_oracle_mt_threads())


def wgcl_mt(edges, eweights, comm, embed, vweights, samples=None, max_alphas=N_ALPHA, n_threads=0,
            dist_form=0):
    """The exact-mode undirected algorithm of :func:`wgcl` on ``n_threads`` host cores (0 = all):
    cge_oracle_mt.c, the parallel CPU baseline of SURVEY.md 8(d).  Returns (out[7], trace).

    ``dist_form=1`` is the numerics experiment for the recompute regime's planned row-norm / dot
    form (centred embedding, FMA dot products, difference form under cancellation, extrema from
    the same arithmetic); 0 is the reference's difference form."""
    lib = _load()
    lib.cge_oracle_mt_set_dist_form(int(dist_form))
    edges = _i64(edges)
    src, dst = _i64(edges[:, 0]), _i64(edges[:, 1])
    ew, cm, em, vw = _f64(eweights), _i64(np.asarray(comm).reshape(-1)), _f64(embed), _f64(vweights)
    n = int(max(src.max(), dst.max()))
    if cm.shape[0] != n:
        raise AssertionError("No. communities not matching no. vertices")
    if samples is None:
        K, ns = 0, 1
        pi = pj = ni = nj = np.zeros(1, dtype=np.int64)
        pw = np.zeros(1)
    else:
        pi, pj, pw, ni, nj = samples
        pi, pj, ni, nj = (np.atleast_2d(_i64(x)) for x in (pi, pj, ni, nj))
        pw = np.atleast_2d(_f64(pw))
        ns, K = pi.shape
    i64, f64 = C.POINTER(C.c_int64), C.POINTER(C.c_double)
    lib.cge_oracle_wgcl_mt.argtypes = [C.c_int64, i64, i64, f64, i64, C.c_int64, f64, C.c_int64, f64,
                                       C.c_int64, C.c_int64, i64, i64, f64, i64, i64, C.c_int,
                                       C.c_int, f64, C.POINTER(OracleMtTrace)]
    lib.cge_oracle_wgcl_mt.restype = C.c_int
    out = np.zeros(7)
    tr = OracleMtTrace()
    rc = lib.cge_oracle_wgcl_mt(src.shape[0], _p(src, C.c_int64), _p(dst, C.c_int64),
                                _p(ew, C.c_double), _p(cm, C.c_int64), n, _p(em, C.c_double),
                                em.shape[1], _p(vw, C.c_double), K, ns, _p(pi, C.c_int64),
                                _p(pj, C.c_int64), _p(pw, C.c_double), _p(ni, C.c_int64),
                                _p(nj, C.c_int64), int(max_alphas), int(n_threads),
                                _p(out, C.c_double), C.byref(tr))
    lib.cge_oracle_mt_set_dist_form(0)
    if rc != 0:
        raise RuntimeError(f"parallel oracle error {rc}")
    return out, tr


class OracleStreamTrace(C.Structure):
    _fields_ = [
        ("n_alpha_run", C.c_int32),
        ("iters", C.c_int32 * N_ALPHA),
        ("div", C.c_double * N_ALPHA),
        ("auc", C.c_double * N_ALPHA),
        ("lo", C.c_double),
        ("hi", C.c_double),
        ("threads", C.c_int32),
        ("cached", C.c_int32),
        ("seconds", C.c_double),
    ]


def wgcl_stream(edges, eweights, comm, embed, vweights, samples=None, directed=False,
                max_alphas=N_ALPHA, n_threads=0, mem_budget=0):
    """cge_oracle_stream.c: exact-mode ``wGCL`` / ``wGCL_directed`` (divergence.jl:27-257 / 282-561)
    without any O(n^2) array, on ``n_threads`` cores (0 = all) -- the CPU answer for BASELINE configs
    3 and 4, where the reference's packed arrays fit no host.  ``mem_budget`` bytes may be used to
    keep the per-pair q after the first sweep (same values, faster).  Returns (out[7], trace)."""
    lib = _load()
    edges = _i64(edges)
    src, dst = _i64(edges[:, 0]), _i64(edges[:, 1])
    ew, cm, em, vw = _f64(eweights), _i64(np.asarray(comm).reshape(-1)), _f64(embed), _f64(vweights)
    n = int(max(src.max(), dst.max()))
    if cm.shape[0] != n:
        raise AssertionError("No. communities not matching no. vertices")
    if samples is None:
        K, ns = 0, 1
        pi = pj = ni = nj = np.zeros(1, dtype=np.int64)
        pw = np.zeros(1)
    else:
        pi, pj, pw, ni, nj = samples
        pi, pj, ni, nj = (np.atleast_2d(_i64(x)) for x in (pi, pj, ni, nj))
        pw = np.atleast_2d(_f64(pw))
        ns, K = pi.shape
    i64, f64 = C.POINTER(C.c_int64), C.POINTER(C.c_double)
    lib.cge_oracle_stream.argtypes = [C.c_int, C.c_int64, i64, i64, f64, C.c_int64, i64, f64, C.c_int64,
                                      f64, C.c_int64, C.c_int64, i64, i64, f64, i64, i64, C.c_int,
                                      C.c_int, C.c_int64, f64, C.POINTER(OracleStreamTrace)]
    lib.cge_oracle_stream.restype = C.c_int
    out = np.zeros(7)
    tr = OracleStreamTrace()
    rc = lib.cge_oracle_stream(int(bool(directed)), src.shape[0], _p(src, C.c_int64),
                               _p(dst, C.c_int64), _p(ew, C.c_double), n, _p(cm, C.c_int64),
                               _p(em, C.c_double), em.shape[1], _p(vw, C.c_double), K, ns,
                               _p(pi, C.c_int64), _p(pj, C.c_int64), _p(pw, C.c_double),
                               _p(ni, C.c_int64), _p(nj, C.c_int64), int(max_alphas), int(n_threads),
                               int(mem_budget), _p(out, C.c_double), C.byref(tr))
    if rc != 0:
        raise RuntimeError(f"streaming oracle error {rc}")
    return out, tr
