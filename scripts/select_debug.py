"""Debug aid for cge_b200_landmarks_select: where does the device assignment leave the host mirror's?"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cge_jl_b200 import divergence as dv  # noqa: E402
from util import clusters_of, load_fixture, planted_partition  # noqa: E402

lm = importlib.import_module("cge_jl_b200.landmarks")
sc = dv.Scorer(0)
cases = [("test115", load_fixture("test115.npz"), 20, 1), ("test115w", load_fixture("test115_weighted.npz"), 20, 1),
         ("test115", load_fixture("test115.npz"), 13, 1), ("test115", load_fixture("test115.npz"), 40, 4),
         ("pp300", planted_partition(300, 4, 5, seed=3, weighted=True), 30, 2)]
for name, (edges, ew, vw, comm, emb), land, forced in cases:
    cl = clusters_of(comm)
    for rule in ("rss", "size", "diameter"):
        for eig in ("lapack", "builtin"):
            g, cuts = sc.landmarks_select(emb, vw, cl, land, forced, rule, eig=eig)
            lm.CANONICAL_SIGN = eig == "builtin"
            ref = lm.runsplit(emb, vw, cl, land, forced, lm.RULES[rule] if hasattr(lm, "RULES") else
                              {"rss": lm.split_cluster_rss, "size": lm.split_cluster_size,
                               "diameter": lm.split_cluster_diameter}[rule])
            lm.CANONICAL_SIGN = False
            bad = np.nonzero(g != ref)[0]
            # same partition up to relabeling?
            pairs = set(zip(g.tolist(), ref.tolist()))
            relabel = len(pairs) == len(set(g.tolist())) == len(set(ref.tolist()))
            print(name, land, forced, rule, eig, "cuts", cuts, "differ", bad.size, "same partition" if relabel else "PARTITION DIFFERS",
                  "sizes dev", np.bincount(g)[:12].tolist(), "ref", np.bincount(ref)[:12].tolist(), flush=True)
