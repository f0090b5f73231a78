#!/bin/bash
# round 2, GPU call 22: smoke(), selection tests and timing after the pinned-buffer change
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_c22_smoke.txt 2>&1; tail -1 gpurun_out/r02_c22_smoke.txt
timeout 600 python -m pytest tests/test_gpu_select.py tests/test_gpu_landmarks.py -m gpu -q > gpurun_out/r02_c22_select_tests.txt 2>&1
tail -3 gpurun_out/r02_c22_select_tests.txt
rm -f gpurun_out/select_bench.jsonl
timeout 900 python scripts/select_bench.py 1000000 4000 > gpurun_out/r02_c22_select_bench.txt 2>&1
tail -1 gpurun_out/r02_c22_select_bench.txt
