#!/bin/bash
# round 2, GPU call 7: recompute regime with the DMMA Gram step; directed stored pass with the TMA ring driver
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q -k "recompute or store_what_fits or super_tiles or dot_form or regimes_agree or abcd or random_small" > gpurun_out/r02_c7_tests.txt 2>&1
tail -5 gpurun_out/r02_c7_tests.txt
rm -f gpurun_out/config_runs.jsonl
CGE_B200_PHASES=1 timeout 600 python scripts/run_config.py --synthetic 20000,128,64,0 --regime 2 --max-alphas 2 > gpurun_out/r02_c7_d128.txt 2>&1
grep "us per pass" gpurun_out/r02_c7_d128.txt; tail -1 gpurun_out/r02_c7_d128.txt | cut -c1-900
timeout 600 python scripts/run_config.py --synthetic 20000,128,64,1 --regime 2 --max-alphas 2 > gpurun_out/r02_c7_d128_dir.txt 2>&1
tail -1 gpurun_out/r02_c7_d128_dir.txt | cut -c1-700
timeout 600 python scripts/run_config.py --config 2 --regime 2 > gpurun_out/r02_c7_cfg2_rc.txt 2>&1
tail -1 gpurun_out/r02_c7_cfg2_rc.txt | cut -c1-700
for drv in 2 3; do
  timeout 600 python scripts/run_config.py --config 3 --driver $drv > gpurun_out/r02_c7_cfg3_drv$drv.txt 2>&1
  tail -1 gpurun_out/r02_c7_cfg3_drv$drv.txt | cut -c1-900
done
