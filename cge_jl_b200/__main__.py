"""``python -m cge_jl_b200 -g G -e E [-c C] [flags]`` -- the CGE_CLI.jl driver
(/root/reference/example/CGE_CLI.jl:1-25) on top of the B200 scorer; same flags, same output."""
import os

import numpy as np

# One-shot CLI process: CUDA's lazy module loading would pay for every kernel instantiation on its
# first use inside the scoring call (0.39 s for the 10k example against 0.08 s when the library's
# modules are loaded up front; measured with scripts/cold_start.py).  Must be set before the CUDA
# runtime initialises; long-lived hosts that share the process with other CUDA libraries keep the
# default.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

from . import landmarks, parseargs
from .divergence import Scorer, wGCL, wGCL_directed


def main(argv=None):
    (edges, weights, vweights, comm, clusters, embed, verbose, land, forced, method, directed,
     split, seed, samples) = parseargs(argv)
    distances = np.zeros(vweights.shape[0])
    init_edges = np.zeros((0, 0), dtype=np.int64)
    init_vweights, init_eweights = np.zeros(0), np.zeros(0)
    init_embed = np.zeros((0, 0))
    v_to_l = np.zeros(0, dtype=np.int64)
    sc = Scorer()  # one device handle for the landmark aggregation and the scoring call
    try:
        if land != -1:
            init_edges, init_vweights = edges.copy(), vweights.copy()
            init_eweights, init_embed = weights.copy(), embed.copy()
            # selection on the host (north_star), aggregation on the device
            distances, embed, comm, edges, weights, vweights, v_to_l = landmarks(
                edges, weights, vweights, clusters, comm, embed, verbose, land, forced, method,
                directed, device=sc)
        f = wGCL_directed if directed else wGCL
        results = f(edges, weights, comm, embed, distances, vweights, init_vweights, v_to_l,
                    init_edges, init_eweights, init_embed, split, seed, samples, verbose, scorer=sc)
    finally:
        sc.close()
    print([float(x) for x in results])
    return results


if __name__ == "__main__":
    main()
