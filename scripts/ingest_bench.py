"""Times SURVEY.md 8(f) F3 -- the native table reader -- beside numpy.loadtxt on an embedding
file of the reference's layout (id column + d shortest-round-trip decimals per row).  Host-only.

  python scripts/ingest_bench.py [--rows 100000 --d 128]      # -> profiles/r01_ingest_bench.jsonl
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cge_jl_b200.auxilary import readdlm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--d", type=int, default=128)
    args = ap.parse_args()
    emb = np.random.default_rng(0).normal(scale=0.7, size=(args.rows, args.d))
    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, "bench.embedding")
        with open(fn, "w") as f:
            for i, row in enumerate(emb):
                f.write(f"{i} " + " ".join(repr(float(x)) for x in row) + "\n")
        mb = os.path.getsize(fn) / 1e6
        res = {"rows": args.rows, "d": args.d, "file_mb": round(mb, 1),
               "host_threads": os.cpu_count()}
        for th in (1, 0):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                got = readdlm(fn, n_threads=th)
                best = min(best, time.perf_counter() - t0)
            res[f"native_s_threads{th or 'all'}"] = round(best, 4)
            res[f"native_mb_per_s_threads{th or 'all'}"] = round(mb / best)
        t0 = time.perf_counter()
        ref = np.loadtxt(fn, ndmin=2)
        res["numpy_loadtxt_s"] = round(time.perf_counter() - t0, 3)
        res["numpy_loadtxt_mb_per_s"] = round(mb / res["numpy_loadtxt_s"])
        res["bit_identical"] = bool(np.array_equal(got, ref) and np.array_equal(got[:, 1:], emb))
    with open(os.path.join(ROOT, "profiles", "r01_ingest_bench.jsonl"), "a") as f:
        f.write(json.dumps(res) + "\n")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
