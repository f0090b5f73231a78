#!/bin/bash
# round 2, GPU call 6 (8 GPUs): 4- and 8-rank parity tests, bench at N=8 on config 4, config 5 exact (1M vertices)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "4gpu or 8gpu" > gpurun_out/r02_c6_multi.txt 2>&1
tail -4 gpurun_out/r02_c6_multi.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 900 $TR bench.py --gpus 8 --steps 2 --warmup 2 > gpurun_out/bench_cfg4_n8.json 2> gpurun_out/bench_cfg4_n8.err
tail -c 1500 gpurun_out/bench_cfg4_n8.json; tail -3 gpurun_out/bench_cfg4_n8.err
rm -f gpurun_out/config_runs.jsonl
CGE_B200_PHASES=1 timeout 1500 $TR scripts/run_config.py --config 5 --exact --max-alphas 2 --spot 8 > gpurun_out/r02_c6_config5.txt 2>&1
grep -E "store what fits|us per pass|recompute regime" gpurun_out/r02_c6_config5.txt | head -20
tail -1 gpurun_out/r02_c6_config5.txt | cut -c1-1800
nvidia-smi --query-gpu=index,memory.used,memory.total --format=csv,noheader | head -8
