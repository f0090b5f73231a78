"""Times SURVEY.md 8(f) F1 -- the device sampler of non-edges (cge_b200_sample_non_edges) --
beside the host rejection sampler it replaces, on a 1M-vertex graph with 8M random edges
(the reference's own NE construction, divergence.jl:121-137, cannot run at this size:
5*10^11 tuples).  Appends one JSON line to gpurun_out/sampler_bench.jsonl.

  python scripts/sampler_bench.py [--n 1000000 --m 8000000 --k 10000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cge_jl_b200 import divergence as dv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=8_000_000)
    ap.add_argument("--k", type=int, default=10_000)
    args = ap.parse_args()
    rng = np.random.default_rng(11)
    edges = rng.integers(1, args.n + 1, size=(args.m, 2))
    edges = edges[edges[:, 0] != edges[:, 1]]
    sc = dv.Scorer(0)
    sc.sample_non_edges(edges[:1000], args.n, 16)  # context + module load outside the timings
    res = {"n": args.n, "m": int(edges.shape[0]), "k": args.k}
    for sets in (1, 40):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            ni, nj, draws = sc.sample_non_edges(edges, args.n, args.k, n_sets=sets, seed=42,
                                                return_draws=True)
            best = min(best, time.perf_counter() - t0)
        res[f"device_s_sets{sets}"] = best
        res[f"draws_per_sample_sets{sets}"] = draws
    codes = np.minimum(edges[:, 0], edges[:, 1]) * (args.n + 1) + np.maximum(edges[:, 0], edges[:, 1])
    t0 = time.perf_counter()
    hi, hj = dv._sample_non_edges(np.random.default_rng(42), args.n, codes, args.k, False)
    res["host_numpy_s_sets1"] = time.perf_counter() - t0
    # both samplers return non-edges only
    for a, b in ((ni[0], nj[0]), (hi, hj)):
        assert not np.isin(a * (args.n + 1) + b, codes).any() and np.all(a < b)
    res["h2d_bytes"] = int(edges.shape[0]) * 16
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sampler_bench.jsonl"), "a") as f:
        f.write(json.dumps(res) + "\n")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
