#!/bin/bash
# round 2, GPU call 19: whole GPU suite on the final build, the default bench line, ncu launch list of a step
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_c19_pytest.txt 2>&1
tail -6 gpurun_out/r02_c19_pytest.txt
timeout 1500 python bench.py > gpurun_out/r02_c19_bench_n1.json 2> gpurun_out/r02_c19_bench_n1.err
tail -c 1500 gpurun_out/r02_c19_bench_n1.json | head -c 1500; echo
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c19_launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > gpurun_out/r02_c19_ncu.log 2>&1
tail -2 gpurun_out/r02_c19_ncu.log | cut -c1-300
