#!/bin/bash
# round 2, GPU call 3: ncu capture of the recompute fixed-point kernel at d=128 + the full GPU suite
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_fixed_point_rc -c 1 \
  -o gpurun_out/prof_r02_rc_d128 -f python scripts/run_config.py --synthetic 20000,128,64,0 --regime 2 --max-alphas 1 > gpurun_out/r02_c3_ncu.log 2>&1
tail -2 gpurun_out/r02_c3_ncu.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c3_pytest.txt 2>&1
tail -5 gpurun_out/r02_c3_pytest.txt
