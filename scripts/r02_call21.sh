#!/bin/bash
# round 2, GPU call 21: fused-vs-unfused B test, compute-sanitizer attempt on the new kernels, cold first call
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "deferred or reproducible or exact_10k" > gpurun_out/r02_c21_tests.txt 2>&1
tail -3 gpurun_out/r02_c21_tests.txt
timeout 300 python scripts/sanitize_case.py > gpurun_out/r02_c21_sanitize_plain.txt 2>&1
tail -4 gpurun_out/r02_c21_sanitize_plain.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/r02_c21_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/r02_c21_memcheck.txt
for i in 1 2; do timeout 300 python scripts/run_config.py --config 2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cold cfg2 s_run', d['s_run'], 'fused', d['b_fused'])"; done
