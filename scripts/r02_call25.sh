#!/bin/bash
# round 2, GPU call 25: A/B of the per-batch row reduction (CGE_ROWRED4, libcge_b200_rr.so): bit-identical
# partial slots, 4 instead of 16 live row sums per warp in the pass and in the fused pass
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "
import sys,json
l=json.loads(sys.stdin.read()); c=l['config']; b=c['ms_breakdown_last_step']; r=l['roofline']
print('$2', 'ms_per_step %.3f' % l['ms_per_step'], 'fp %.2f b %.2f' % (b['fp_kernels'], b['b_kernels']), 'frac %.4f pass_us %.2f' % (r['frac'], r['avg_pass_us']), 'fused_us', r['fused_pass'] and round(r['fused_pass']['avg_launch_us'],1), {k: v.get('ok') for k, v in l.items() if k.startswith('parity')}, c['result'][:2])" || tail -5 $1; }
for lib in libcge_b200.so libcge_b200_rr.so; do
  CGE_B200_LIB=$PWD/cge_jl_b200/$lib timeout 400 python bench.py --workload 4 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r02_c25_w4_$lib.txt 2>&1
  show gpurun_out/r02_c25_w4_$lib.txt "w4 $lib"
  CGE_B200_FUSE_B=1 CGE_B200_LIB=$PWD/cge_jl_b200/$lib timeout 300 python bench.py --workload 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c25_w2_$lib.txt 2>&1
  show gpurun_out/r02_c25_w2_$lib.txt "w2 $lib"
done
CGE_B200_LIB=$PWD/cge_jl_b200/libcge_b200_rr.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "deferred or exact_10k or reproducible or one_pass" > gpurun_out/r02_c25_tests.txt 2>&1
tail -3 gpurun_out/r02_c25_tests.txt
