#!/bin/bash
# round 2, GPU call 20: selection at 1M with the faster eigen-solver, unique-rows test, config 5 landmark half,
# the default bench line with the fixed-point launches and the fused passes reported separately
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_select.py tests/test_gpu_landmarks.py -m gpu -q > gpurun_out/r02_c20_select_tests.txt 2>&1
tail -3 gpurun_out/r02_c20_select_tests.txt
rm -f gpurun_out/select_bench.jsonl gpurun_out/config_runs.jsonl
timeout 900 python scripts/select_bench.py 1000000 4000 > gpurun_out/r02_c20_select_bench.txt 2>&1
tail -1 gpurun_out/r02_c20_select_bench.txt
timeout 900 python scripts/run_config.py --config 5 > gpurun_out/r02_c20_cfg5_landmarks.txt 2>&1
tail -2 gpurun_out/r02_c20_cfg5_landmarks.txt | cut -c1-300
CGE_B200_EIG=builtin timeout 900 python scripts/run_config.py --config 5 > gpurun_out/r02_c20_cfg5_landmarks_builtin.txt 2>&1
tail -2 gpurun_out/r02_c20_cfg5_landmarks_builtin.txt | cut -c1-300
cp gpurun_out/config_runs.jsonl gpurun_out/r02_c20_config_runs.jsonl
timeout 1500 python bench.py > gpurun_out/r02_c20_bench_n1.json 2> gpurun_out/r02_c20_bench_n1.err
python -c "
import json
l=json.loads(open('gpurun_out/r02_c20_bench_n1.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['frac'], l['roofline']['avg_pass_us'], l['roofline']['fused_pass'], l['clocks'])"
