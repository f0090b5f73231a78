"""Host-side argument and file handling: the Python mirror of CGE.jl's ``parseargs``.

In a deployment with Julia this step stays in Julia (north_star: parsing, Louvain and
landmark selection are not on the B200 path).  Julia is absent from this image, so the
host side above the C-ABI is mirrored here with the same flags, the same return tuple
and the same error behaviour (usage text + exit status 1), following
``/root/reference/src/auxilary.jl:61-247`` step by step.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import landmarks as _lm

_USAGE = (
    "\n\nUsage:\n"
    "\tpython -m cge_jl_b200 -g edgelist -e embedding [-c communities] [--seed seed] "
    "[--samples-local samples] [-v] [-d] [--split-global] [-l [landmarks]] [-f [forced]] "
    "[--force-exact] [-m method]\n"
    "\nParameters:\n"
    "edgelist: rows should contain two whitespace separated vertices ids (edge) and optional "
    "weights in third column\n"
    "embedding: rows should contain whitespace separated embeddings of vertices\n"
    "communities: rows should contain cluster identifiers of vertices with optional vertices ids "
    "in the first column\n"
    "if no file is given communities are calculated with Louvain algorithm\n"
    "seed: RNG seed for local measure sampling\n"
    "samples: no. samples to draw for local score calculation\n"
    "-v: flag for debugging messages\n"
    "-d: flag for usage of directed framework\n"
    "--split-global: flag for using splitted global score; kept for backward compatibility\n"
    "landmarks: required number of landmarks; 4*sqrt(no.vertices) by default\n"
    "forced: required number of forced splits of a cluster; 4 by default\n"
    "method: one of rss, rss2, size, diameter\n"
)


def readdlm(path, dtype=np.float64, skiprows=0, order="C", n_threads=0):
    """Whitespace-delimited numeric table, the typed ``readdlm(fn, Float64 | Int)`` of
    ``auxilary.jl:86,123,150-155``, parsed by the native multi-threaded reader of libcge_b200.so
    (``cge_b200_table_dims`` + ``cge_b200_read_table``).  Raises ``ValueError``
    on ragged rows and on cells that are not numbers, like the reference's call does; ``order="F"``
    returns Julia's column-major layout."""
    import ctypes as C

    from . import _lib

    lib = _lib.load()
    rows, cols = C.c_int64(), C.c_int64()
    fn = os.fsencode(path)
    if lib.cge_b200_table_dims(fn, int(skiprows), int(n_threads), C.byref(rows), C.byref(cols)) != 0:
        raise ValueError(_lib.last_error())
    arr = np.empty((rows.value, cols.value), dtype=np.float64, order=order)
    rs, cs = (s // 8 for s in arr.strides) if arr.size else (cols.value, 1)
    if lib.cge_b200_read_table(fn, int(skiprows), int(n_threads), rows.value, cols.value, rs, cs,
                               arr.ctypes.data_as(C.POINTER(C.c_double))) != 0:
        raise ValueError(_lib.last_error())
    if np.issubdtype(np.dtype(dtype), np.integer):
        if not np.all(arr == np.round(arr)):
            raise ValueError(f"InexactError: {path} holds non-integer values")
        return arr.astype(dtype)
    return arr


_readdlm = readdlm


def _find(argv, flag):
    try:
        return argv.index(flag)
    except ValueError:
        return None


def _require(cond, msg):
    """The reference's ``@assert`` (auxilary.jl:81-158): raises AssertionError also under ``python -O``."""
    if not cond:
        raise AssertionError(msg)


def _parse(argv):
    methods = {
        "rss": _lm.split_cluster_rss,
        "rss2": _lm.split_cluster_rss2,
        "size": _lm.split_cluster_size,
        "diameter": _lm.split_cluster_diameter,
    }
    # flags (auxilary.jl:72-76)
    verbose = "-v" in argv
    directed = "-d" in argv
    split = "--split-global" in argv

    # edgelist (auxilary.jl:80-112)
    i = _find(argv, "-g")
    _require(i is not None, "Edgelist file is required")
    fn_edges = argv[i + 1]
    _require(os.path.isfile(fn_edges), f"{fn_edges} is not a file")
    raw = _readdlm(fn_edges, np.float64)
    rows, no_cols = raw.shape
    if verbose:
        print(f"{no_cols} columns and {rows} rows in edgelist file.")
    _require(no_cols in (2, 3), "Expected 2 or 3 columns in edgelist file")
    v_min = raw[:, :2].min()
    _require(v_min in (0.0, 1.0), "Vertices should be either 0-based or 1-based")
    if v_min == 0.0:
        raw[:, :2] += 1.0
    no_vertices = int(raw[:, :2].max())
    if verbose:
        print(f"Graph contains {no_vertices} vertices")
    eweights = np.ones(rows) if no_cols == 2 else raw[:, 2].copy()
    ids = raw[:, :2]
    if not np.all(ids == np.round(ids)):
        raise ValueError("InexactError: vertex ids must be integers")
    edges = ids.astype(np.int64)
    vweight = np.zeros(no_vertices)
    np.add.at(vweight, edges[:, 0] - 1, eweights)
    np.add.at(vweight, edges[:, 1] - 1, eweights)
    if verbose:
        print("Done preparing edgelist and vertices weights")

    # communities (auxilary.jl:115-141)
    i = _find(argv, "-c")
    if i is not None:
        fn_comm = argv[i + 1]
    else:
        from .clustering import louvain_clust

        if no_cols == 2:
            louvain_clust(v_min, fn_edges)
        else:
            louvain_clust(fn_edges, edges, eweights)
        fn_comm = fn_edges + ".ecg"
    comm = _readdlm(fn_comm, np.int64)
    comm_rows, c_cols = comm.shape
    _require(comm_rows == no_vertices,
             f"No. communities ({comm_rows}) differ from no. nodes ({no_vertices})")
    _require(c_cols in (1, 2),
             f"Expected 1 or 2 columns in communities file, but encountered {c_cols}.")
    if c_cols == 2:
        comm = comm[np.argsort(comm[:, 0], kind="stable"), 1].reshape(-1, 1)
    c_min = comm.min()
    _require(c_min in (0, 1),
             f"Communities should be either 0-based or 1-based, but are {c_min} based.")
    if c_min == 0:
        comm = comm + 1
    comm = np.ascontiguousarray(comm.reshape(-1, 1))
    if verbose:
        print("Done preparing communities.")

    # embedding (auxilary.jl:144-168)
    i = _find(argv, "-e")
    _require(i is not None, "Embedding file is required")
    fn_embed = argv[i + 1]
    _require(os.path.isfile(fn_embed), f"{fn_embed} is not a file")
    try:
        embedding = _readdlm(fn_embed, np.float64)
    except ValueError:
        if verbose:
            print("Embedding in node2vec format. Loading without first line.")
        embedding = _readdlm(fn_embed, np.float64, skiprows=1)
    _require(no_vertices == embedding.shape[0],
             "No. rows in embedding and no. vertices in a graph differ.")
    first = embedding[:, 0]
    if np.all(first == np.round(first)):
        if verbose:
            print("Sorting embedding by first column")
        order = first.astype(np.int64)
        embedding = embedding[np.argsort(order, kind="stable"), 1:]
    embedding = np.ascontiguousarray(embedding)
    if verbose:
        print("Done preparing embedding.")

    # landmarks (auxilary.jl:172-208)
    clusters = []
    landmarks = -1
    i = _find(argv, "-l")
    if i is not None:
        try:
            landmarks = int(argv[i + 1])
        except (IndexError, ValueError):
            landmarks = int(round(4 * np.sqrt(no_vertices)))
            print(f"[ Info: Using {landmarks} landmarks", file=sys.stderr)
    i = _find(argv, "-f")
    if i is not None:
        forced = int(argv[i + 1])
        landmarks = 1 if landmarks == -1 else landmarks
    else:
        forced = 4
    if no_vertices >= 10000 and "--force-exact" not in argv and landmarks == -1:
        landmarks = max(int(round(4 * np.sqrt(no_vertices))), 4 * int(comm.max()))
        print(
            "[ Info: Number of vertices is equal or higher than 10 000. Automatically switching "
            f"to approximate algortihm with {landmarks} landmarks. If you want to force exact "
            "algorithm use --force-exact flag.",
            file=sys.stderr,
        )
    if landmarks != -1:
        by_comm = {}
        for v, c in enumerate(comm[:, 0], start=1):
            by_comm.setdefault(int(c), []).append(v)
        clusters = [np.asarray(v, dtype=np.int64) for v in by_comm.values()]

    i = _find(argv, "--seed")
    seed = int(argv[i + 1]) if i is not None else -1
    i = _find(argv, "--samples-local")
    samples = int(argv[i + 1]) if i is not None else 10000
    i = _find(argv, "-m")
    method_str = argv[i + 1].strip().lower() if i is not None else "rss"
    method = methods[method_str]
    return (edges, eweights, vweight, comm, clusters, embedding, verbose, landmarks, forced,
            method, directed, split, seed, samples)


def parseargs(argv=None):
    """Parse CGE_CLI flags and input files (mirror of ``auxilary.jl:63-247``).

    Returns the reference's 14-tuple ``(edges, eweights, vweight, comm, clusters, embedding,
    verbose, landmarks, forced, method, directed, split, seed, samples)`` with 1-based
    ``edges``/``comm``/``clusters`` exactly as the Julia function does.  On any error the
    message and the usage text are printed and the process exits with status 1.
    """
    argv = list(sys.argv[1:] if argv is None else argv)
    try:
        return _parse(argv)
    except Exception as e:  # auxilary.jl:221-246
        print(f"{type(e).__name__}: {e}", file=sys.stderr)
        print(_USAGE)
        raise SystemExit(1)
