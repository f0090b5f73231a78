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
from cge_jl_b200.landmarks import (landmarks, split_cluster_diameter, split_cluster_rss,  # noqa: E402
                                   split_cluster_size)
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
# per-alpha instantiations forced on a small problem: the deferred B sweep (k_bfp + k_fixed_point without its
# first tile phase)
os.environ["CGE_B200_RT_EXPONENT"] = "0"
os.environ["CGE_B200_FUSE_B"] = "1"
edges, ew, vw, comm, emb = planted_partition(700, 5, 12, seed=3, weighted=True)
out, st = dv.wGCL(edges, ew, comm, emb, np.zeros(700), vw, *EMPTY, False, 42, 300, False, scorer=sc, driver=2,
                  regime=1, max_alphas=4, return_stats=True)
print("fused B", out[:2], "b_fused", st.b_fused)
del os.environ["CGE_B200_RT_EXPONENT"], os.environ["CGE_B200_FUSE_B"]
# landmark mode: selection (cge_b200_landmarks_select, cge_b200_unique_rows) and aggregation on the device, the
# tensor-core diameter filter forced on a small graph
os.environ["CGE_B200_DIAM_MIN"] = "100"
edges, ew, vw, comm, emb = planted_partition(1500, 6, 40, seed=5, weighted=True)
for rule in (split_cluster_rss, split_cluster_size, split_cluster_diameter):
    lm = landmarks(edges, ew, vw, clusters_of(comm), comm, emb, False, 60, 2, rule, False, device=sc)
    print("device landmarks", rule.__name__, lm[1].shape)
out = dv.wGCL(lm[3], lm[4], lm[2], lm[1], lm[0], lm[5], vw, lm[6], edges, ew, emb, False, 42, 200,
              False, scorer=sc, max_alphas=3)
print("landmarks + diameter filter", out[:2])
edges, ew, vw, comm, emb = load_fixture("test115.npz")
lm = landmarks(edges, ew, vw, clusters_of(comm), comm, emb, False, 20, 1, split_cluster_rss, False)
out = dv.wGCL(lm[3], lm[4], lm[2], lm[1], lm[0], lm[5], vw, lm[6], edges, ew, emb, False, 42, 200,
              False, scorer=sc, max_alphas=3)
print("landmarks", out[:2])
sc.close()
print("done")
