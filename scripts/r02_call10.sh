#!/bin/bash
# round 2, GPU call 10: F2 landmarks aggregation on the device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_landmarks.py -m gpu -x -q > gpurun_out/r02_c10_tests.txt 2>&1
tail -15 gpurun_out/r02_c10_tests.txt
timeout 900 python scripts/landmarks_bench.py > gpurun_out/r02_c10_bench.txt 2>&1
tail -2 gpurun_out/r02_c10_bench.txt
