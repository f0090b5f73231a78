#!/bin/bash
# round 2, GPU call 18: warp-specialised diameter filter (producer warp + 4 epilogue warps, two TMEM
# accumulators): tests, timing, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -k "diameter or landmark" > gpurun_out/r02_c18_diam_tests.txt 2>&1
tail -4 gpurun_out/r02_c18_diam_tests.txt
timeout 600 python -m pytest tests/test_gpu_select.py -m gpu -q > gpurun_out/r02_c18_select_tests.txt 2>&1
tail -3 gpurun_out/r02_c18_select_tests.txt
rm -f gpurun_out/config_runs.jsonl
timeout 600 python scripts/run_config.py --synthetic 60000,128,64,0 --landmarks 300 > gpurun_out/r02_c18_diam60k.txt 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_diameter_filter -c 1 -f -o gpurun_out/prof_r02_diameter_ws \
  python scripts/run_config.py --synthetic 60000,128,64,0 --landmarks 300 > gpurun_out/r02_c18_ncu_diam.log 2>&1
tail -1 gpurun_out/r02_c18_diam60k.txt | cut -c1-300
timeout 900 python scripts/run_config.py --config 5 > gpurun_out/r02_c18_cfg5_landmarks.txt 2>&1
tail -2 gpurun_out/r02_c18_cfg5_landmarks.txt | cut -c1-400
cp gpurun_out/config_runs.jsonl gpurun_out/r02_c18_config_runs.jsonl
