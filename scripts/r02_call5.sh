#!/bin/bash
# round 2, GPU call 5 (2 GPUs): new single-GPU tests, the multi-GPU suite, bench at N=2, hybrid timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "store_what_fits or super_tiles or sample_sets" > gpurun_out/r02_c5_parity.txt 2>&1
tail -3 gpurun_out/r02_c5_parity.txt
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r02_c5_multi.txt 2>&1
tail -4 gpurun_out/r02_c5_multi.txt
rm -f gpurun_out/config_runs.jsonl
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for mb in 0 3500; do
  CGE_B200_STORE_MB=$mb CGE_B200_PHASES=1 timeout 600 $TR scripts/run_config.py --synthetic 60000,128,64,0 --regime 2 --max-alphas 1 --spot 4 > gpurun_out/r02_c5_hybrid_$mb.txt 2>&1
  grep -E "store what fits|us per pass" gpurun_out/r02_c5_hybrid_$mb.txt | head -4
  tail -1 gpurun_out/r02_c5_hybrid_$mb.txt | cut -c1-900
done
timeout 900 $TR bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_cfg4_n2.json 2> gpurun_out/bench_cfg4_n2.err
tail -c 1200 gpurun_out/bench_cfg4_n2.json; tail -3 gpurun_out/bench_cfg4_n2.err
