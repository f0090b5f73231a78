#!/bin/bash
# one-shot validation of the opt-in row-norm/dot form of the recompute regime (CGE_B200_RC_FORM=dot)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
export CUDA_MODULE_LOADING=EAGER
CGE_B200_RC_FORM=dot timeout 45 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q \
  -k "recompute or regimes_agree or abcd" > gpurun_out/dot_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/dot_pytest.txt
timeout 30 python - > gpurun_out/dot_timing.txt 2>&1 <<'P'
import os, time, numpy as np
from cge_jl_b200 import divergence as dv
z = np.load("tests/golden/example10k.npz")
edges, ew, vw, comm, emb = (z[k] for k in ("edges", "eweights", "vweights", "comm", "embedding"))
n = 10000
E = (np.zeros(0), np.zeros(0, dtype=np.int64), np.zeros((0, 0), dtype=np.int64), np.zeros(0), np.zeros((0, 0)))
samples = dv.draw_samples(edges, ew, n, 10000, 42, False, True)
sc = dv.Scorer(0)
for form in ("diff", "dot", "diff", "dot"):
    os.environ["CGE_B200_RC_FORM"] = form
    out, st = dv.wGCL(edges, ew, comm, emb, np.zeros(n), vw, *E, False, 42, 10000, False, samples=samples,
                      return_stats=True, scorer=sc, regime=2)
    print(form, "ms_sweeps", round(st.ms_sweeps, 1), "passes", st.fp_sweeps, "us/pass", round(1e3 * st.ms_sweeps / st.fp_sweeps, 1),
          [float(x) for x in out], flush=True)
P
tail -n 3 gpurun_out/dot_pytest.txt; cat gpurun_out/dot_timing.txt
