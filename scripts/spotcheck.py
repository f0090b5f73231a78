"""Sub-block spot check for runs no oracle can reach (SURVEY.md section 7 "(iv) sub-block spot checks at
1M"): the degree sums S_i of a handful of vertices recomputed from scratch in NumPy,

    S_i = T_i * sum_j T_j * (1 - ||x_i - x_j|| / hi)^alpha        (divergence.jl:142-159, lo = 0),

against the S the device reports for the last pass (cge_b200_debug_read).  T of that pass is recovered
from the updated T the device holds: T_new = T + eps*T*(w/S - 1)  (divergence.jl:160-165).  O(rows * n * d)
on the host: ~1 s per vertex at 10^6 vertices, d = 128.  Used by scripts/run_config.py --config 5 --exact
and by tests/test_gpu_scale.py on a problem small enough for the test suite."""
import numpy as np


def degree_sum_spot_check(emb, t_new, s_dev, w, hi, alpha, rows, eps=0.25):
    """Largest relative difference |S_i(numpy) - S_i(device)| / S_i over `rows` (0-based vertices)."""
    emb = np.asarray(emb, dtype=np.float64)
    t_old = t_new / (1.0 + eps * (w / s_dev - 1.0))
    worst = 0.0
    for i in rows:
        d = np.sqrt(((emb - emb[i]) ** 2).sum(axis=1))
        d[i] = 0.0
        x = np.maximum(1.0 - d / hi, 0.0)
        s = t_old[i] * float(np.dot(t_old, x ** alpha))
        worst = max(worst, abs(s - s_dev[i]) / abs(s_dev[i]))
    return worst
