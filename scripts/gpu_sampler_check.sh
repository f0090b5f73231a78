#!/bin/bash
# F1 device sampler: its tests, the full GPU suite, the sampler timing, smoke()
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
export CUDA_MODULE_LOADING=EAGER
timeout 600 python -m pytest tests/test_gpu_sampler.py -x -q > gpurun_out/f1_pytest_sampler.txt 2>&1; echo "rc=$?" >> gpurun_out/f1_pytest_sampler.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f1_pytest_all.txt 2>&1; echo "rc=$?" >> gpurun_out/f1_pytest_all.txt
timeout 300 python scripts/sampler_bench.py > gpurun_out/f1_sampler_bench.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f1_smoke.txt 2>&1
tail -n 4 gpurun_out/f1_pytest_sampler.txt; tail -n 3 gpurun_out/f1_pytest_all.txt; tail -n 2 gpurun_out/f1_sampler_bench.txt; tail -n 1 gpurun_out/f1_smoke.txt
