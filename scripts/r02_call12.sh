#!/bin/bash
# round 2, GPU call 12 (8 GPUs): config 5 exact with the final recompute kernel (DMMA Gram step, short epilogue)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
rm -f gpurun_out/config_runs.jsonl
CGE_B200_PHASES=1 timeout 1200 $TR scripts/run_config.py --config 5 --exact --max-alphas 2 --spot 8 > gpurun_out/r02_c12_config5.txt 2>&1
grep -E "store what fits|us per pass" gpurun_out/r02_c12_config5.txt | head -6
tail -1 gpurun_out/r02_c12_config5.txt | cut -c1-1800
