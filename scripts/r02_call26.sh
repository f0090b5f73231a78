#!/bin/bash
# round 2, GPU call 26: (a) directed pass with compiler-scheduled loads (CGE_D_PLAIN, libcge_b200_dp.so) vs the
# explicit register double buffering, config 3; (b) fused pass with the branch-free fast path (new default) vs
# the batched one (libcge_b200_ubold.so), config 4
mkdir -p gpurun_out
rm -f gpurun_out/config_runs.jsonl
showc() { tail -1 $1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$2', 's_run %.4f' % l['s_run'], 'fp_ms %.2f' % l['ms_fp_kernels'], 'b_ms %.2f' % l['ms_b_kernels'], 'pass_ms %.4f' % l['avg_pass_ms'], 'passes', l['fp_passes'], l['result'][:2])" || tail -5 $1; }
for lib in libcge_b200_ubold.so libcge_b200_dp.so libcge_b200_ubold.so libcge_b200_dp.so; do
  CGE_B200_LIB=$PWD/cge_jl_b200/$lib timeout 300 python scripts/run_config.py --config 3 > gpurun_out/r02_c26_cfg3_$lib.txt 2>&1
  showc gpurun_out/r02_c26_cfg3_$lib.txt "cfg3 $lib"
done
show() { tail -1 $1 | python -c "
import sys,json
l=json.loads(sys.stdin.read()); c=l['config']; b=c['ms_breakdown_last_step']; r=l['roofline']
print('$2', 'ms_per_step %.3f' % l['ms_per_step'], 'fp %.2f b %.2f' % (b['fp_kernels'], b['b_kernels']), 'frac %.4f pass_us %.2f' % (r['frac'], r['avg_pass_us']), 'fused_us', r['fused_pass'] and round(r['fused_pass']['avg_launch_us'],1), {k: v.get('ok') for k, v in l.items() if k.startswith('parity')}, c['result'][:2])" || tail -5 $1; }
for lib in libcge_b200_ubold.so libcge_b200.so; do
  CGE_B200_LIB=$PWD/cge_jl_b200/$lib timeout 400 python bench.py --workload 4 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r02_c26_w4_$lib.txt 2>&1
  show gpurun_out/r02_c26_w4_$lib.txt "w4 $lib"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "deferred or exact_10k or reproducible" > gpurun_out/r02_c26_tests.txt 2>&1
tail -3 gpurun_out/r02_c26_tests.txt
