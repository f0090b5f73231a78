"""Size-level parity (SURVEY.md section 7 "(ii)"): BASELINE configs 3 and 4 against
the frozen answers of the streaming CPU oracle (tests/golden/make_size_goldens.py, hours of CPU).
The reference's packed D / GD / P arrays fit no host at these sizes; the streaming oracle keeps its
per-pair arithmetic and is itself held to the line-by-line oracle at small sizes
(tests/test_oracle_stream.py).  Bars as everywhere: identical pass counts per alpha, global and local
score per alpha within 1e-9 relative."""
import json
import os

import numpy as np
import pytest

from cge_jl_b200 import divergence as dv
from util import GOLDEN, RTOL

pytestmark = pytest.mark.gpu


def _golden(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated yet (tests/golden/make_size_goldens.py)")
    return json.load(open(path))


def _compare(st, g):
    na = g["alphas"]
    assert int(st.n_alpha_run) == na
    assert [int(x) for x in list(st.iters)[:na]] == g["iters"], "fixed-point pass counts differ"
    np.testing.assert_allclose(np.array(list(st.div))[:na], np.array(g["div"]), rtol=RTOL, atol=0)
    np.testing.assert_allclose(np.array(list(st.auc))[:na], np.array(g["auc"]), rtol=RTOL, atol=1e-15)
    assert abs(st.hi - g["hi"]) <= 1e-15 * g["hi"] and st.lo == 0.0


@pytest.mark.parametrize("regime", [1, 2], ids=["stored", "recompute"])
def test_config3_directed_50k_against_streaming_oracle(regime):
    """50 000 vertices, directed, weighted, d = 64 (1.25e9 unordered pairs): the first alphas of the
    frozen oracle run -- all of them in the stored regime, the first two when recomputing."""
    from cge_jl_b200.synth import planted_partition
    g = _golden("config3_oracle.json")
    n = 50000
    e, w, vw, c, emb = planted_partition(n, k=32, d=64, seed=1003, directed=True, weighted=True)
    samples = dv.draw_samples(e, w, n, 10000, 42, True, True)
    na = g["alphas"] if regime == 1 else min(2, g["alphas"])
    sc = dv.Scorer(0)
    try:
        p, keep = dv.make_problem(e, w, c, emb, np.zeros(n), vw, None, None, None, False, True,
                                  samples, na, 0, regime)
        sc.upload(p, keep)
        out, st = sc.run()
    finally:
        sc.close()
    assert st.regime == regime
    gg = dict(g, alphas=na, iters=g["iters"][:na], div=g["div"][:na], auc=g["auc"][:na])
    _compare(st, gg)
    if na == g["alphas"] == g["max_alphas"] or na == 40:
        np.testing.assert_allclose(out, np.array(g["out"]), rtol=RTOL, atol=1e-15)


def test_config4_abcd_200k_against_streaming_oracle():
    """200 000 vertices, d = 128, 64 communities (2.0e10 pairs, 160 GB of stored tiles on one B200):
    the first alpha of the frozen oracle run (47 passes)."""
    import torch
    from cge_jl_b200.synth import abcd_like
    g = _golden("config4_oracle_alpha1.json")
    if torch.cuda.mem_get_info(0)[0] < 170 * 2**30:
        pytest.skip("needs 170 GB of free HBM")
    n = 200000
    e, w, vw, c, emb = abcd_like(n, k=64, d=128, seed=1004)
    samples = dv.draw_samples(e, w, n, 10000, 42, False, True)
    sc = dv.Scorer(0)
    try:
        p, keep = dv.make_problem(e, w, c, emb, np.zeros(n), vw, None, None, None, False, False,
                                  samples, g["alphas"], 0, 1)
        sc.upload(p, keep)
        out, st = sc.run()
    finally:
        sc.close()
    assert st.regime == 1
    _compare(st, g)
