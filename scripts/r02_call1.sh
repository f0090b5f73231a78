#!/bin/bash
# round 2, GPU call 1: FP64 pipe microbenchmarks, recompute-regime baseline at d=128, GPU suite with the dot form
mkdir -p gpurun_out
python - > gpurun_out/r02_pipes.json 2> gpurun_out/r02_pipes.err <<'P'
import json
from cge_jl_b200 import divergence as dv
sc = dv.Scorer(0)
r = sc.fp64_pipes_tflops(); r["dfma_peak_old"] = sc.fp64_peak_tflops()
r2 = sc.fp64_pipes_tflops()
print(json.dumps({"first": r, "second": r2}))
P
cat gpurun_out/r02_pipes.json
rm -f gpurun_out/config_runs.jsonl
for reg in 2 3; do
  timeout 600 python scripts/run_config.py --synthetic 20000,128,64,0 --regime $reg --max-alphas 2 > gpurun_out/r02_rc_d128_reg$reg.txt 2>&1
  tail -1 gpurun_out/r02_rc_d128_reg$reg.txt | cut -c1-600
done
CGE_B200_RC_FORM=dot timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_dot.txt 2>&1
tail -3 gpurun_out/r02_pytest_dot.txt
