"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck): every kernel family once.

  compute-sanitizer --tool memcheck python scripts/sanitize_case.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cge_jl_b200 import divergence as dv  # noqa: E402
from cge_jl_b200.landmarks import landmarks, split_cluster_rss  # noqa: E402
from util import clusters_of, empty_landmark_args, load_fixture, planted_partition  # noqa: E402

EMPTY = empty_landmark_args()
sc = dv.Scorer(0)
for directed in (False, True):
    n = 300
    edges, ew, vw, comm, emb = planted_partition(n, 4, 20, seed=11, directed=directed, weighted=True)
    f = dv.wGCL_directed if directed else dv.wGCL
    for driver, regime in ((1, 1), (2, 1), (3, 1), (1, 2), (2, 2)):
        out = f(edges, ew, comm, emb, np.zeros(n), vw, *EMPTY, False, 42, 300, False, scorer=sc,
                driver=driver, regime=regime, max_alphas=3)
        print(directed, driver, regime, out[:2])
edges, ew, vw, comm, emb = load_fixture("test115.npz")
lm = landmarks(edges, ew, vw, clusters_of(comm), comm, emb, False, 20, 1, split_cluster_rss, False)
out = dv.wGCL(lm[3], lm[4], lm[2], lm[1], lm[0], lm[5], vw, lm[6], edges, ew, emb, False, 42, 200,
              False, scorer=sc, max_alphas=3)
print("landmarks", out[:2])
sc.close()
print("done")
