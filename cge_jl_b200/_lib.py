"""ctypes binding of libcge_b200.so (the C ABI declared in include/cge_b200.h).

Loading fails loudly when the library has not been built: the scorer has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

N_ALPHA = 40
_HERE = os.path.dirname(os.path.abspath(__file__))
# CGE_B200_LIB overrides the location (the variable julia/CGEB200.jl reads as well)
LIB_PATH = os.environ.get("CGE_B200_LIB") or os.path.join(_HERE, "libcge_b200.so")

OK, ERR_ARG, ERR_ASSERT_COMM, ERR_ASSERT_DIST, ERR_OOM, ERR_CUDA, ERR_NCCL, ERR_STATE = (
    0, -1, -2, -3, -4, -5, -6, -7)
DRIVER_AUTO, DRIVER_HOSTLOOP, DRIVER_PERSISTENT, DRIVER_RING = 0, 1, 2, 3
REGIME_AUTO, REGIME_STORED, REGIME_RECOMPUTE, REGIME_RECOMPUTE_DOT, REGIME_RECOMPUTE_DIFF = 0, 1, 2, 3, 4

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int64)


class Problem(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("index_base", C.c_int32),
        ("directed", C.c_int32), ("split", C.c_int32),
        ("m", C.c_int64), ("edge_src", _pi), ("edge_dst", _pi), ("eweights", _pd),
        ("n_comm", C.c_int64), ("comm", _pi),
        ("embed", _pd), ("embed_rows", C.c_int64), ("d", C.c_int64),
        ("embed_row_stride", C.c_int64), ("embed_col_stride", C.c_int64),
        ("n_distances", C.c_int64), ("distances", _pd), ("vweights", _pd),
        ("n_full", C.c_int64), ("init_vweights", _pd), ("v_to_l", _pi), ("init_embed", _pd),
        ("init_row_stride", C.c_int64), ("init_col_stride", C.c_int64),
        ("n_samples", C.c_int64), ("n_sets", C.c_int64),
        ("pos_i", _pi), ("pos_j", _pi), ("pos_w", _pd), ("neg_i", _pi), ("neg_j", _pi),
        ("max_alphas", C.c_int32), ("driver", C.c_int32),
        ("regime", C.c_int32), ("reserved", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("n_alpha_run", C.c_int32),
        ("iters", C.c_int32 * N_ALPHA), ("div", C.c_double * N_ALPHA),
        ("auc", C.c_double * N_ALPHA),
        ("lo", C.c_double), ("hi", C.c_double), ("hi_full", C.c_double),
        ("n", C.c_int64), ("n_pairs", C.c_int64),
        ("fp_sweeps", C.c_int64), ("b_sweeps", C.c_int64),
        ("matrix_bytes", C.c_int64), ("launches", C.c_int64),
        ("n_tiles", C.c_int32), ("grid", C.c_int32), ("driver", C.c_int32),
        ("n_ranks", C.c_int32), ("regime", C.c_int32), ("diam_candidate_tiles", C.c_int32),
        ("ms_upload", C.c_float), ("ms_build", C.c_float), ("ms_solve", C.c_float),
        ("ms_total", C.c_float), ("ms_sweeps", C.c_float), ("ms_bsweeps", C.c_float),
        ("b_fused", C.c_int32), ("ms_fused", C.c_float),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


# cge_b200_eigvec_fn: int (*)(const double *c, int64_t d, double *v, void *user)
EIGVEC_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.c_void_p)

EXPORTS = [
    "cge_b200_version", "cge_b200_device_count", "cge_b200_last_error", "cge_b200_score",
    "cge_b200_score_multi",
    "cge_b200_create", "cge_b200_destroy", "cge_b200_upload", "cge_b200_run",
    "cge_b200_comm_id_size", "cge_b200_comm_unique_id", "cge_b200_comm_init",
    "cge_b200_shard_plan", "cge_b200_debug_read", "cge_b200_p2p_handle_size",
    "cge_b200_p2p_export", "cge_b200_p2p_import", "cge_b200_measure_fp64_peak",
    "cge_b200_selftest_math", "cge_b200_sample_non_edges", "cge_b200_table_dims",
    "cge_b200_read_table", "cge_b200_measure_fp64_pipes", "cge_b200_landmarks_aggregate",
    "cge_b200_landmarks_select", "cge_b200_sym_top_eigvec", "cge_b200_unique_rows",
]

_lib = None


def load():
    """Return the loaded library; raises ImportError if libcge_b200.so is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C cge_jl_b200/csrc` or "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.cge_b200_version.argtypes = [C.POINTER(C.c_int)] * 3
    lib.cge_b200_version.restype = None
    lib.cge_b200_device_count.restype = C.c_int
    lib.cge_b200_last_error.restype = C.c_char_p
    lib.cge_b200_score.argtypes = [C.POINTER(Problem), _pd, C.POINTER(C.c_int32), C.POINTER(Stats)]
    lib.cge_b200_score_multi.argtypes = [C.POINTER(Problem), C.c_int, _pd, C.POINTER(C.c_int32),
                                         C.POINTER(Stats)]
    lib.cge_b200_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.cge_b200_destroy.argtypes = [vp]
    lib.cge_b200_destroy.restype = None
    lib.cge_b200_upload.argtypes = [vp, C.POINTER(Problem)]
    lib.cge_b200_run.argtypes = [vp, _pd, C.POINTER(C.c_int32), C.POINTER(Stats)]
    lib.cge_b200_comm_id_size.restype = C.c_int
    lib.cge_b200_comm_unique_id.argtypes = [vp]
    lib.cge_b200_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.cge_b200_shard_plan.argtypes = [C.c_int64, C.c_int, C.c_int, _pi, _pi, _pi]
    lib.cge_b200_p2p_handle_size.restype = C.c_int
    lib.cge_b200_p2p_export.argtypes = [vp, C.c_int64, vp]
    lib.cge_b200_p2p_import.argtypes = [vp, vp]
    lib.cge_b200_measure_fp64_peak.argtypes = [vp, _pd]
    lib.cge_b200_measure_fp64_pipes.argtypes = [vp, _pd]
    lib.cge_b200_sample_non_edges.argtypes = [vp, C.c_int64, C.c_int64, _pi, _pi, C.c_int32, C.c_int32,
                                              C.c_int64, C.c_int64, C.c_uint64, _pi, _pi, _pd]
    lib.cge_b200_table_dims.argtypes = [C.c_char_p, C.c_int64, C.c_int32, _pi, _pi]
    lib.cge_b200_read_table.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_int64, _pd]
    lib.cge_b200_selftest_math.argtypes = [vp, C.c_int64, C.c_uint64, _pi, _pi]
    lib.cge_b200_landmarks_aggregate.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, _pi, C.c_int32, _pd, _pi,
                                                 _pd, C.c_int64, C.c_int64, C.c_int64, _pi, _pi, _pd,
                                                 C.c_int32, _pd, _pd, _pd, _pi, _pi, _pi, _pd, C.c_int64, _pi]
    lib.cge_b200_landmarks_select.argtypes = [vp, C.c_int64, C.c_int64, _pd, C.c_int64, C.c_int64, _pd,
                                              C.c_int64, _pi, _pi, C.c_int32, C.c_int64, C.c_int64,
                                              C.c_int32, EIGVEC_FN, vp, _pi, _pi]
    lib.cge_b200_sym_top_eigvec.argtypes = [_pd, C.c_int64, _pd, _pd]
    lib.cge_b200_unique_rows.argtypes = [vp, C.c_int64, C.c_int64, _pd, C.c_int64, C.c_int64, _pi]
    lib.cge_b200_debug_read.argtypes = [vp, C.c_int, _pd, C.c_int64]
    _lib = lib
    return lib


def last_error():
    return load().cge_b200_last_error().decode("utf-8", "replace")


def shard_plan(n, rank, n_ranks):
    """(n_tiles, tile_begin, tile_end) owned by ``rank`` -- host-only, no GPU needed."""
    nt, b, e = C.c_int64(), C.c_int64(), C.c_int64()
    rc = load().cge_b200_shard_plan(n, rank, n_ranks, C.byref(nt), C.byref(b), C.byref(e))
    if rc != 0:
        raise ValueError(last_error())
    return nt.value, b.value, e.value
