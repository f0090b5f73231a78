"""landmarks() aggregation (landmarks.jl:387-463) on the device against the NumPy
restatement, 1M vertices, d = 128, 8M edges, 4000 landmarks.  Appends to gpurun_out/landmarks_bench.jsonl."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cge_jl_b200 import divergence as dv  # noqa: E402
from cge_jl_b200.landmarks import aggregate_host  # noqa: E402
from cge_jl_b200.synth import planted_partition  # noqa: E402


def main():
    n, N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 4000
    edges, ew, vw, comm, emb = planted_partition(n, k=64, d=128, seed=1005)
    rng = np.random.default_rng(7)
    lm = (comm[:, 0] - 1) * (N // 64) + rng.integers(0, N // 64, size=n) + 1  # landmarks inside communities
    N = int(lm.max())
    sc = dv.Scorer(0)
    sc.landmarks_aggregate(lm[:1000], vw[:1000], comm[:1000], emb[:1000], edges[:10] * 0 + 1, ew[:10], False, N)
    t0 = time.perf_counter()
    dev = sc.landmarks_aggregate(lm, vw, comm, emb, edges, ew, False, N)
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    host = aggregate_host(lm, edges, ew, vw, comm, emb, False)
    t_host = time.perf_counter() - t0
    same = all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(dev, host))
    if not same:
        for nm, a, b in zip(("dii", "embed", "cluster", "edges", "weights", "lweight"), dev, host):
            print(nm, a.shape, b.shape, a.shape == b.shape and np.array_equal(a, b), file=sys.stderr)
    line = {"n": n, "d": 128, "m": int(edges.shape[0]), "landmarks": N, "landmark_edges": int(dev[3].shape[0]),
            "s_device_incl_h2d_d2h": t_dev, "s_numpy": t_host, "bit_identical": bool(same)}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "landmarks_bench.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
