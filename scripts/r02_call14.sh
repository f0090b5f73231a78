#!/bin/bash
# round 2, GPU call 14: A/B of (a) row buffers carried across tiles (CGE_XTILE, libcge_b200.so vs
# libcge_b200_x0.so) and (b) the B sweep fused into the next alpha's first pass (CGE_B200_FUSE_B)
mkdir -p gpurun_out
rm -f gpurun_out/config_runs.jsonl
show() { tail -1 $1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$2', 's_run %.4f' % l['s_run'], 'fp_ms %.2f' % l['ms_fp_kernels'], 'b_ms %.2f' % l['ms_b_kernels'], 'pass_ms %.4f' % l['avg_pass_ms'], 'fused', l['b_fused'], 'passes', l['fp_passes'], l['result'][:2], l['result'][4:6])" || tail -5 $1; }
for lib in libcge_b200_x0.so libcge_b200.so; do
  for fuse in 0 1; do
    tag=$lib.f$fuse
    CGE_B200_FUSE_B=$fuse CGE_B200_LIB=$PWD/cge_jl_b200/$lib timeout 300 python scripts/run_config.py --config 2 > gpurun_out/r02_c14_cfg2_$tag.txt 2>&1
    show gpurun_out/r02_c14_cfg2_$tag.txt "cfg2 $tag"
    CGE_B200_FUSE_B=$fuse CGE_B200_LIB=$PWD/cge_jl_b200/$lib timeout 300 python scripts/run_config.py --config 4 --max-alphas 6 > gpurun_out/r02_c14_cfg4_$tag.txt 2>&1
    show gpurun_out/r02_c14_cfg4_$tag.txt "cfg4a6 $tag"
  done
  CGE_B200_LIB=$PWD/cge_jl_b200/$lib timeout 300 python scripts/run_config.py --config 3 > gpurun_out/r02_c14_cfg3_$lib.txt 2>&1
  show gpurun_out/r02_c14_cfg3_$lib.txt "cfg3 $lib"
done
timeout 600 python -m pytest tests/test_gpu_select.py tests/test_gpu_landmarks.py -m gpu -q > gpurun_out/r02_c14_select_tests.txt 2>&1
tail -15 gpurun_out/r02_c14_select_tests.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/r02_c14_tests.txt 2>&1
tail -3 gpurun_out/r02_c14_tests.txt
cp gpurun_out/config_runs.jsonl gpurun_out/r02_c14_config_runs.jsonl
