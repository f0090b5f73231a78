#!/bin/bash
# round 2, GPU call 13: A/B of the staged recompute epilogue (libcge_b200_epi2.so) against the default
mkdir -p gpurun_out
rm -f gpurun_out/config_runs.jsonl
for lib in libcge_b200.so libcge_b200_epi2.so; do
  for rep in 1 2; do
    CGE_B200_LIB=$PWD/cge_jl_b200/$lib timeout 600 python scripts/run_config.py --synthetic 20000,128,64,0 --regime 2 --max-alphas 4 > gpurun_out/r02_c13_$lib.$rep.txt 2>&1
    tail -1 gpurun_out/r02_c13_$lib.$rep.txt | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$lib', l['avg_pass_ms'], l['fp64_frac'], l['iters'], l['result'][:2])"
  done
  CGE_B200_LIB=$PWD/cge_jl_b200/$lib timeout 600 python scripts/run_config.py --config 2 --regime 2 > gpurun_out/r02_c13_cfg2_$lib.txt 2>&1
  tail -1 gpurun_out/r02_c13_cfg2_$lib.txt | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$lib cfg2', l['avg_pass_ms'], l['fp64_frac'], l['result'][:2])"
done
CGE_B200_LIB=$PWD/cge_jl_b200/libcge_b200_epi2.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q -k "recompute or store_what_fits or super_tiles or dot_form or regimes_agree or spot" > gpurun_out/r02_c13_tests.txt 2>&1
tail -3 gpurun_out/r02_c13_tests.txt
