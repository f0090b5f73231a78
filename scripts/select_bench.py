"""landmark selection (runsplit, landmarks.jl:279-345, rss rule) on the device against
the NumPy mirror: n vertices (default 1M), d = 128, 64 communities, -l 4000 -f 4 -- the landmark half of
BASELINE config 5.  Appends to gpurun_out/select_bench.jsonl.

  python scripts/select_bench.py [n] [land] [--no-host]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cge_jl_b200 import divergence as dv  # noqa: E402
import importlib  # noqa: E402

lm_mod = importlib.import_module("cge_jl_b200.landmarks")
from cge_jl_b200.synth import planted_partition  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if args else 1000000
    land = int(args[1]) if len(args) > 1 else 4000
    edges, ew, vw, comm, emb = planted_partition(n, k=64, d=128, seed=1005)
    by = {}
    for v, c in enumerate(comm[:, 0], start=1):
        by.setdefault(int(c), []).append(v)
    clusters = [np.asarray(v, dtype=np.int64) for v in by.values()]
    sc = dv.Scorer(0)
    sc.landmarks_select(emb[:2000], vw[:2000], [np.arange(1, 2001)], 8, 1, "rss")  # warm-up (module load)
    line = {"n": n, "d": 128, "land": land, "forced": 4, "rule": "rss"}
    for eig in ("builtin", "lapack"):
        t0 = time.perf_counter()
        group, cuts = sc.landmarks_select(emb, vw, clusters, land, 4, "rss", eig=eig)
        line[f"s_device_{eig}"] = time.perf_counter() - t0
        line["cuts"] = int(cuts)
        line[f"landmarks_{eig}"] = int(group.max()) + 1
        if eig == "lapack":
            dev_lapack = group
    if "--no-host" not in sys.argv:
        t0 = time.perf_counter()
        ref = lm_mod.runsplit(emb, vw, clusters, land, 4, lm_mod.split_cluster_rss)
        line["s_numpy_mirror"] = time.perf_counter() - t0
        line["labels_equal_lapack_callback_vs_mirror"] = bool(np.array_equal(ref, dev_lapack))
        line["vertices_differing"] = int((ref != dev_lapack).sum())
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "select_bench.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
