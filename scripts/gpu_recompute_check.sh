#!/bin/bash
# recompute-kernel rework: tests, recompute bench, config 3 + d=128 synthetic in recompute, default bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
export CUDA_MODULE_LOADING=EAGER
python - <<'P' > gpurun_out/rc_selftest.txt 2>&1
from cge_jl_b200 import divergence as dv
s = dv.Scorer(0)
for n, seed in ((1 << 24, 1), (1 << 28, 7), (1 << 30, 99)):
    print(n, seed, s.selftest_math(n, seed), flush=True)
P
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/rc_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/rc_pytest.txt
timeout 300 python bench.py --regime 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/rc_bench_recompute.json 2> gpurun_out/rc_bench_recompute.err
n0=$(wc -l < profiles/r01_config_runs.jsonl)
timeout 300 python scripts/run_config.py --config 3 --regime 2 > gpurun_out/rc_cfg3.txt 2>&1
timeout 300 python scripts/run_config.py --synthetic 20000,128,64,0 --regime 2 > gpurun_out/rc_d128.txt 2>&1
tail -n +$((n0+1)) profiles/r01_config_runs.jsonl > gpurun_out/rc_cfg_runs.jsonl
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/rc_bench_default.json 2> gpurun_out/rc_bench_default.err
tail -3 gpurun_out/rc_pytest.txt; cat gpurun_out/rc_selftest.txt; cut -c1-400 gpurun_out/rc_bench_recompute.json
