#!/bin/bash
# round 2, GPU call 16: F4 tests, double-buffered TMEM diameter filter (tests, timing, ncu), selection at 1M,
# config 5 landmark half with landmarks() on the GPU, ncu of the fused B + first-pass kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_select.py tests/test_gpu_landmarks.py -m gpu -q > gpurun_out/r02_c16_select_tests.txt 2>&1
tail -6 gpurun_out/r02_c16_select_tests.txt
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -k "diameter or landmark or cli" > gpurun_out/r02_c16_diam_tests.txt 2>&1
tail -4 gpurun_out/r02_c16_diam_tests.txt
rm -f gpurun_out/config_runs.jsonl gpurun_out/select_bench.jsonl
timeout 600 python scripts/run_config.py --synthetic 60000,128,64,0 --landmarks 300 > gpurun_out/r02_c16_diam60k.txt 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_diameter_filter -c 1 -f -o gpurun_out/prof_r02_diameter \
  python scripts/run_config.py --synthetic 60000,128,64,0 --landmarks 300 > gpurun_out/r02_c16_ncu_diam.log 2>&1
tail -2 gpurun_out/r02_c16_diam60k.txt | cut -c1-600
timeout 900 python scripts/select_bench.py 1000000 4000 > gpurun_out/r02_c16_select_bench.txt 2>&1
tail -2 gpurun_out/r02_c16_select_bench.txt
timeout 900 python scripts/run_config.py --config 5 > gpurun_out/r02_c16_cfg5_landmarks.txt 2>&1
tail -2 gpurun_out/r02_c16_cfg5_landmarks.txt | cut -c1-900
timeout 600 python scripts/run_config.py --synthetic 60000,64,32,0 --max-alphas 3 > gpurun_out/r02_c16_bfp60k.txt 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bfp -c 1 -f -o gpurun_out/prof_r02_bfp \
  python scripts/run_config.py --synthetic 60000,64,32,0 --max-alphas 3 > gpurun_out/r02_c16_ncu_bfp.log 2>&1
tail -1 gpurun_out/r02_c16_bfp60k.txt | cut -c1-400
cp gpurun_out/config_runs.jsonl gpurun_out/r02_c16_config_runs.jsonl
ls -la gpurun_out/*.ncu-rep
