#!/bin/bash
# round 2, GPU call 17 (2 GPUs): multi-GPU tests and the config-4 bench line with the deferred B sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r02_c17_mg_tests.txt 2>&1
tail -5 gpurun_out/r02_c17_mg_tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_c17_bench_n2.txt 2>&1
tail -1 gpurun_out/r02_c17_bench_n2.txt | python -c "
import sys,json
l=json.loads(sys.stdin.read()); c=l['config']
print('n2 ms_per_step %.1f e2e %.1f frac %.4f fused %d' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['frac'], c['b_passes_fused_with_a_fixed_point_pass']), {k: v for k, v in l.items() if k.startswith('parity')})" || tail -20 gpurun_out/r02_c17_bench_n2.txt
