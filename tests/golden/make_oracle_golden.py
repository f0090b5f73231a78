"""Freezes oracle results that are too slow to recompute inside the test-suite.

Run in the build container:  python tests/golden/make_oracle_golden.py
Writes tests/golden/oracle_example10k_exact.npz : the oracle's wGCL result for the reference's
10k example in exact mode (--force-exact --seed 42, BASELINE.json configs[1]) together with the
sampled pairs that were used, so the GPU test feeds the identical pairs through the C ABI.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from cge_jl_b200.divergence import draw_samples  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    z = np.load(os.path.join(OUT, "example10k.npz"))
    edges, ew, vw, comm, emb = z["edges"], z["eweights"], z["vweights"], z["comm"], z["embedding"]
    n = vw.shape[0]
    samples = draw_samples(edges, ew, n, 10000, 42, directed=False, exact=True)
    t = time.time()
    out, tr = oracle.wgcl(edges, ew, comm, emb, np.zeros(n), vw, samples=samples)
    dt = time.time() - t
    print(out.tolist(), list(tr.iters), dt)
    np.savez_compressed(
        os.path.join(OUT, "oracle_example10k_exact.npz"), out=out,
        iters=np.array(list(tr.iters)), div=np.array(list(tr.div)), auc=np.array(list(tr.auc)),
        lo=tr.lo, hi=tr.hi, n_alpha_run=tr.n_alpha_run, seconds=dt,
        pos_i=samples[0].astype(np.int32), pos_j=samples[1].astype(np.int32), pos_w=samples[2],
        neg_i=samples[3].astype(np.int32), neg_j=samples[4].astype(np.int32))


if __name__ == "__main__":
    main()
