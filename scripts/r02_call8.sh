#!/bin/bash
# round 2, GPU call 8: recompute regime with the short epilogue forms: tests, timing, ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q -k "recompute or store_what_fits or super_tiles or dot_form or regimes_agree or abcd or random_small" > gpurun_out/r02_c8_tests.txt 2>&1
tail -5 gpurun_out/r02_c8_tests.txt
rm -f gpurun_out/config_runs.jsonl
CGE_B200_PHASES=1 timeout 600 python scripts/run_config.py --synthetic 20000,128,64,0 --regime 2 --max-alphas 4 > gpurun_out/r02_c8_d128.txt 2>&1
grep "us per pass" gpurun_out/r02_c8_d128.txt; tail -1 gpurun_out/r02_c8_d128.txt | cut -c1-400
timeout 600 python scripts/run_config.py --config 2 --regime 2 > gpurun_out/r02_c8_cfg2_rc.txt 2>&1
tail -1 gpurun_out/r02_c8_cfg2_rc.txt | cut -c1-300
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_fixed_point_rc -c 1 \
  -o gpurun_out/prof_r02_rc_d128_mma -f python scripts/run_config.py --synthetic 20000,128,64,0 --regime 2 --max-alphas 1 > gpurun_out/r02_c8_ncu.log 2>&1
tail -2 gpurun_out/r02_c8_ncu.log | cut -c1-200
