#!/bin/bash
# round 2, GPU call 2: the rewritten recompute regime (super-tiles, TMA ring, dot form default)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "recompute or sample_sets" > gpurun_out/r02_c2_parity.txt 2>&1
tail -5 gpurun_out/r02_c2_parity.txt
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "regimes_agree or dot_form or abcd" > gpurun_out/r02_c2_scale.txt 2>&1
tail -5 gpurun_out/r02_c2_scale.txt
rm -f gpurun_out/config_runs.jsonl
for reg in 2 4; do
  CGE_B200_PHASES=1 timeout 600 python scripts/run_config.py --synthetic 20000,128,64,0 --regime $reg --max-alphas 2 > gpurun_out/r02_c2_d128_reg$reg.txt 2>&1
  tail -3 gpurun_out/r02_c2_d128_reg$reg.txt | cut -c1-700
done
CGE_B200_RC_SB=4 timeout 600 python scripts/run_config.py --synthetic 20000,128,64,0 --regime 2 --max-alphas 2 > gpurun_out/r02_c2_d128_sb4.txt 2>&1
tail -1 gpurun_out/r02_c2_d128_sb4.txt | cut -c1-700
timeout 600 python scripts/run_config.py --config 2 --regime 2 > gpurun_out/r02_c2_cfg2_rc.txt 2>&1
tail -1 gpurun_out/r02_c2_cfg2_rc.txt | cut -c1-700
