"""Freezes the CPU answers for BASELINE configs 3 and 4 -- sizes where the line-by-line oracle (and
the reference itself) cannot allocate its packed arrays -- with the streaming oracle
(oracle/cge_oracle_stream.c, held to the line-by-line oracle by tests/test_oracle_stream.py).

  python tests/golden/make_size_goldens.py --config 4 [--alphas 1] [--threads 6]   # ~80 min on 6 cores
  python tests/golden/make_size_goldens.py --config 3 [--alphas 40]                # ~40 min, q cached (10 GB)

Writes tests/golden/config4_oracle_alpha1.json / config3_oracle.json: per-alpha pass counts, global
and local scores for the first `alphas` grid points, the distance maximum and the run time.  The
inputs are the generators of cge_jl_b200/synth.py with the seeds of scripts/run_config.py and
bench.py; the sampled pairs come from draw_samples(seed 42), which is deterministic, so the GPU
tests (tests/test_gpu_size.py) and bench.py regenerate identical inputs.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from cge_jl_b200.divergence import draw_samples  # noqa: E402
from cge_jl_b200.synth import abcd_like, planted_partition  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def inputs(cfg):
    if cfg == 3:
        e, w, vw, c, emb = planted_partition(50000, k=32, d=64, seed=1003, directed=True, weighted=True)
        return e, w, vw, c, emb, True
    e, w, vw, c, emb = abcd_like(200000, k=64, d=128, seed=1004)
    return e, w, vw, c, emb, False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4])
    ap.add_argument("--alphas", type=int, default=0)
    ap.add_argument("--threads", type=int, default=6)
    ap.add_argument("--mem-gb", type=float, default=14.0)
    args = ap.parse_args()
    alphas = args.alphas or (40 if args.config == 3 else 1)
    e, w, vw, c, emb, directed = inputs(args.config)
    n = vw.shape[0]
    samples = draw_samples(e, w, n, 10000, 42, directed, True)
    t0 = time.time()
    out, tr = oracle.wgcl_stream(e, w, c, emb, vw, samples=samples, directed=directed,
                                 max_alphas=alphas, n_threads=args.threads,
                                 mem_budget=int(args.mem_gb * 2**30))
    na = int(tr.n_alpha_run)
    res = {"config": args.config, "n": int(n), "d": int(emb.shape[1]), "directed": directed,
           "alphas": na, "max_alphas": alphas, "iters": [int(x) for x in list(tr.iters)[:na]],
           "div": [float(x) for x in list(tr.div)[:na]], "auc": [float(x) for x in list(tr.auc)[:na]],
           "out": [float(x) for x in out], "hi": float(tr.hi), "threads": int(tr.threads),
           "q_cached": bool(tr.cached), "seconds": time.time() - t0,
           "made_by": "tests/golden/make_size_goldens.py (oracle/cge_oracle_stream.c)"}
    name = "config3_oracle.json" if args.config == 3 else f"config4_oracle_alpha{alphas}.json"
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
