#!/bin/bash
# round 2, GPU call 11: whole GPU suite (with the config-4 size test), landmark aggregation timing at 1M
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_c11_pytest.txt 2>&1
tail -8 gpurun_out/r02_c11_pytest.txt
rm -f gpurun_out/landmarks_bench.jsonl
timeout 900 python scripts/landmarks_bench.py > gpurun_out/r02_c11_lm.txt 2>&1
tail -3 gpurun_out/r02_c11_lm.txt
