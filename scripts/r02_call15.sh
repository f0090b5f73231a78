#!/bin/bash
# round 2, GPU call 15: (a) where the device landmark selection leaves the host mirror (debug), (b) warm A/B
# of the B sweep fused into the next alpha's first pass (CGE_B200_FUSE_B) with 4 / 8 / 16 rows requested
# together in the fused pass (libcge_b200_ub4.so / libcge_b200.so / libcge_b200_ub16.so)
mkdir -p gpurun_out
timeout 300 python scripts/select_debug.py > gpurun_out/r02_c15_select_debug.txt 2>&1
tail -40 gpurun_out/r02_c15_select_debug.txt
show() { tail -1 $1 | python -c "
import sys,json
l=json.loads(sys.stdin.read()); c=l['config']; b=c['ms_breakdown_last_step']
print('$2', 'ms_per_step %.3f' % l['ms_per_step'], 'e2e %.3f' % l['e2e']['ms_per_step'], 'fp %.2f b %.2f build %.2f' % (b['fp_kernels'], b['b_kernels'], b['build']), 'fused', c['b_passes_fused_with_a_fixed_point_pass'], 'frac %.4f' % l['roofline']['frac'], 'bgbs', l['roofline']['b_sweep_gbs'], {k: v.get('ok') for k, v in l.items() if k.startswith('parity')}, c['result'][:2])" || tail -5 $1; }
for lib in libcge_b200.so libcge_b200_ub4.so libcge_b200_ub16.so; do
  for fuse in 0 1; do
    [ $fuse = 0 ] && [ $lib != libcge_b200.so ] && continue
    tag=$lib.f$fuse
    CGE_B200_FUSE_B=$fuse CGE_B200_LIB=$PWD/cge_jl_b200/$lib timeout 300 python bench.py --workload 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c15_w2_$tag.txt 2>&1
    show gpurun_out/r02_c15_w2_$tag.txt "w2 $tag"
  done
done
for lib in libcge_b200.so libcge_b200_ub4.so libcge_b200_ub16.so; do
  for fuse in 0 1; do
    [ $fuse = 0 ] && [ $lib != libcge_b200.so ] && continue
    tag=$lib.f$fuse
    CGE_B200_FUSE_B=$fuse CGE_B200_LIB=$PWD/cge_jl_b200/$lib timeout 400 python bench.py --workload 4 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r02_c15_w4_$tag.txt 2>&1
    show gpurun_out/r02_c15_w4_$tag.txt "w4 $tag"
  done
done
