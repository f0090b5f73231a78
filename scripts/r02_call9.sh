#!/bin/bash
# round 2, GPU call 9: the whole GPU suite on the current build
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_c9_pytest.txt 2>&1
tail -8 gpurun_out/r02_c9_pytest.txt
